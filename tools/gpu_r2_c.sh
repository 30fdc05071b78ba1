#!/bin/bash
# round 2, GPU call C (re-run after the container was re-created): pipe rates, safegcd + batched-affine tests, level sweep, sanitizer, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c_smi.txt 2>&1
timeout 300 ./tools/bin/ubench_pipes > gpurun_out/ubench_pipes_r2.jsonl 2> gpurun_out/ubench_pipes_r2.err
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "safegcd or batched_affine" > gpurun_out/r2c_affine_tests.log 2>&1
echo "affine tests rc=$?" >> gpurun_out/r2c_affine_tests.log
tail -15 gpurun_out/r2c_affine_tests.log
timeout 600 python tools/bench_affine.py --levels 0,1,2,3,4,5 --reps 4 > gpurun_out/bench_affine_r2c.jsonl 2> gpurun_out/bench_affine_r2c.err
timeout 300 python tools/bench_affine.py --logn 18 --levels 0,2,3,4 --reps 4 >> gpurun_out/bench_affine_r2c.jsonl 2>> gpurun_out/bench_affine_r2c.err
timeout 300 python tools/bench_affine.py --logn 22 --levels 0,3,4 --reps 2 >> gpurun_out/bench_affine_r2c.jsonl 2>> gpurun_out/bench_affine_r2c.err
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 0,2,3 --reps 3 >> gpurun_out/bench_affine_r2c.jsonl 2>> gpurun_out/bench_affine_r2c.err
timeout 300 python tools/bench_affine.py --curve pallas_ec --logn 20 --levels 0,2,3,4 --reps 3 >> gpurun_out/bench_affine_r2c.jsonl 2>> gpurun_out/bench_affine_r2c.err
cut -c1-420 gpurun_out/bench_affine_r2c.jsonl; tail -3 gpurun_out/bench_affine_r2c.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2 > gpurun_out/r2c_sanitize_affine.log 2>&1; tail -4 gpurun_out/r2c_sanitize_affine.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_affine_pairs -c 2 -o gpurun_out/ncu_affine_r2c python tools/bench_affine.py --levels 3 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
cat gpurun_out/ubench_pipes_r2.jsonl | cut -c1-260
