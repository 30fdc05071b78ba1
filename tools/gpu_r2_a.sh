#!/bin/bash
# round 2, GPU call A: pipe rates, first execution of the batched-affine path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 300 ./tools/bin/ubench_pipes > gpurun_out/ubench_pipes_r2.jsonl 2> gpurun_out/ubench_pipes_r2.err
CTT_B200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "binary_gcd or batched_affine" > gpurun_out/r2a_affine_tests.log 2>&1
echo "affine tests rc=$?" >> gpurun_out/r2a_affine_tests.log
timeout 600 python tools/bench_affine.py --levels 0,1,2,3,4,5 --reps 4 > gpurun_out/bench_affine_r2.jsonl 2> gpurun_out/bench_affine_r2.err
timeout 300 python tools/bench_affine.py --logn 16 --levels 0,1,2,3,4 --reps 4 >> gpurun_out/bench_affine_r2.jsonl 2>> gpurun_out/bench_affine_r2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_affine_level -c 3 -o gpurun_out/ncu_affine_r2a python tools/bench_affine.py --levels 3 --reps 1 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_affine_tests.log; cat gpurun_out/bench_affine_r2.jsonl | cut -c1-400; cat gpurun_out/ubench_pipes_r2.jsonl
