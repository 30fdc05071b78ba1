#!/bin/bash
# round 2, GPU call B: safegcd inversion + new batched-affine pipeline (msm_affine.cuh): tests, sweep, sanitizer, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "safegcd or batched_affine" > gpurun_out/r2b_affine_tests.log 2>&1
echo "affine tests rc=$?" >> gpurun_out/r2b_affine_tests.log
tail -15 gpurun_out/r2b_affine_tests.log
timeout 600 python tools/bench_affine.py --levels 0,1,2,3,4,5 --reps 4 > gpurun_out/bench_affine_r2b.jsonl 2> gpurun_out/bench_affine_r2b.err
timeout 300 python tools/bench_affine.py --logn 18 --levels 0,2,3,4 --reps 4 >> gpurun_out/bench_affine_r2b.jsonl 2>> gpurun_out/bench_affine_r2b.err
timeout 300 python tools/bench_affine.py --logn 22 --levels 0,3,4 --reps 2 >> gpurun_out/bench_affine_r2b.jsonl 2>> gpurun_out/bench_affine_r2b.err
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 0,2,3 --reps 3 >> gpurun_out/bench_affine_r2b.jsonl 2>> gpurun_out/bench_affine_r2b.err
timeout 300 python tools/bench_affine.py --curve pallas_ec --logn 20 --levels 0,2,3,4 --reps 3 >> gpurun_out/bench_affine_r2b.jsonl 2>> gpurun_out/bench_affine_r2b.err
cut -c1-420 gpurun_out/bench_affine_r2b.jsonl; tail -3 gpurun_out/bench_affine_r2b.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2 > gpurun_out/r2b_sanitize_affine.log 2>&1; tail -4 gpurun_out/r2b_sanitize_affine.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_affine_pairs -c 2 -o gpurun_out/ncu_affine_r2b python tools/bench_affine.py --levels 3 --reps 1 > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_ncu.log
