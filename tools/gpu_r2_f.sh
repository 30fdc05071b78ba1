#!/bin/bash
# round 2, GPU call F: two-pipe multiplier microbenchmark + MSM-level variants; out-of-line multiplier calls in the reduce/fix-up
# kernels; contiguous bit planes + radix-16 combine; KZG commitment entry; full GPU suite; bench line; launch list
mkdir -p gpurun_out
UBENCH_TWO_PIPE_ONLY=1 timeout 300 ./tools/bin/ubench > gpurun_out/ubench_two_pipe_r2f.jsonl 2> gpurun_out/ubench_two_pipe_r2f.err
cut -c1-220 gpurun_out/ubench_two_pipe_r2f.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log
tail -6 gpurun_out/r2f_pytest_gpu.log
: > gpurun_out/bench_affine_r2f.jsonl
for v in main m5 m1; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/bench_affine_r2f.jsonl
  L=$PWD/constantine_b200/lib/libctt_b200_msm.so; [ $v != main ] && L=$PWD/constantine_b200/lib/libctt_b200_msm_$v.so
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --levels 0,3 --reps 4 >> gpurun_out/bench_affine_r2f.jsonl 2>> gpurun_out/bench_affine_r2f.err
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --curve pallas_ec --levels 0 --reps 4 >> gpurun_out/bench_affine_r2f.jsonl 2>> gpurun_out/bench_affine_r2f.err
done
echo '{"variant": "main"}' >> gpurun_out/bench_affine_r2f.jsonl
timeout 300 python tools/bench_affine.py --logn 16 --levels 0,2 --reps 4 >> gpurun_out/bench_affine_r2f.jsonl 2>> gpurun_out/bench_affine_r2f.err
timeout 300 python tools/bench_affine.py --logn 18 --levels 0,3 --reps 4 >> gpurun_out/bench_affine_r2f.jsonl 2>> gpurun_out/bench_affine_r2f.err
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 4 --reps 3 >> gpurun_out/bench_affine_r2f.jsonl 2>> gpurun_out/bench_affine_r2f.err
cut -c1-330 gpurun_out/bench_affine_r2f.jsonl; tail -3 gpurun_out/bench_affine_r2f.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; cut -c1-900 gpurun_out/bench_r2f.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2f.csv > gpurun_out/launches_r2f_summary.txt 2>&1; head -30 gpurun_out/launches_r2f_summary.txt
