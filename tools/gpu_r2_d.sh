#!/bin/bash
# round 2, GPU call D: pipe-rate occupancy sweep with clock/power samples; batched-affine kernel with unrolled multipliers
# (+ two occupancy variants); the whole -m gpu suite on the new host plumbing; bench line; ncu of the affine kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 100 > gpurun_out/r2d_smi_ubench.csv 2>&1 &
SMI=$!
timeout 300 ./tools/bin/ubench_pipes > gpurun_out/ubench_pipes_r2d.jsonl 2> gpurun_out/ubench_pipes_r2d.err
kill $SMI
timeout 400 python tools/bench_affine.py --levels 0,1,2,3,4,5 --reps 4 > gpurun_out/bench_affine_r2d.jsonl 2> gpurun_out/bench_affine_r2d.err
for v in v1 v2; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/bench_affine_r2d.jsonl
  CTT_B200_LIB=$PWD/constantine_b200/lib/libctt_b200_msm_$v.so timeout 300 python tools/bench_affine.py --levels 1,3,4 --reps 4 >> gpurun_out/bench_affine_r2d.jsonl 2>> gpurun_out/bench_affine_r2d.err
done
echo '{"variant": "main"}' >> gpurun_out/bench_affine_r2d.jsonl
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 0,3,4 --reps 3 >> gpurun_out/bench_affine_r2d.jsonl 2>> gpurun_out/bench_affine_r2d.err
timeout 300 python tools/bench_affine.py --curve pallas_ec --logn 20 --levels 0,3,4 --reps 3 >> gpurun_out/bench_affine_r2d.jsonl 2>> gpurun_out/bench_affine_r2d.err
timeout 300 python tools/bench_affine.py --logn 16 --levels 0,2,3 --reps 4 >> gpurun_out/bench_affine_r2d.jsonl 2>> gpurun_out/bench_affine_r2d.err
cut -c1-330 gpurun_out/bench_affine_r2d.jsonl; tail -3 gpurun_out/bench_affine_r2d.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.log
tail -8 gpurun_out/r2d_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; cut -c1-1500 gpurun_out/bench_r2d.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_affine_pairs -c 1 -o gpurun_out/ncu_affine_r2d python tools/bench_affine.py --levels 3 --reps 1 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
grep -E "sweep|chains of 4|IMAD \(mad|DFMA|side by side" gpurun_out/ubench_pipes_r2d.jsonl | cut -c1-300 | tail -30
