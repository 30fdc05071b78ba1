#!/bin/bash
# round 2, GPU call E: bit-plane bucket reduction, chunked input pipeline, prefetching affine kernel at 4 blocks/SM -- full GPU suite,
# A/B of the reduce modes and input chunk counts, bench line, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.log
tail -8 gpurun_out/r2e_pytest_gpu.log
: > gpurun_out/bench_affine_r2e.jsonl
for mode in 0 1; do
  echo "{\"reduce_mode\": $mode}" >> gpurun_out/bench_affine_r2e.jsonl
  CTT_B200_REDUCE_MODE=$mode timeout 300 python tools/bench_affine.py --levels 0,3 --reps 4 >> gpurun_out/bench_affine_r2e.jsonl 2>> gpurun_out/bench_affine_r2e.err
  CTT_B200_REDUCE_MODE=$mode timeout 300 python tools/bench_affine.py --logn 16 --levels 0,2 --reps 4 >> gpurun_out/bench_affine_r2e.jsonl 2>> gpurun_out/bench_affine_r2e.err
  CTT_B200_REDUCE_MODE=$mode timeout 300 python tools/bench_affine.py --logn 18 --levels 0,3 --reps 4 >> gpurun_out/bench_affine_r2e.jsonl 2>> gpurun_out/bench_affine_r2e.err
  CTT_B200_REDUCE_MODE=$mode timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 4 --reps 3 >> gpurun_out/bench_affine_r2e.jsonl 2>> gpurun_out/bench_affine_r2e.err
  CTT_B200_REDUCE_MODE=$mode timeout 300 python tools/bench_affine.py --curve pallas_ec --logn 20 --levels 0 --reps 3 >> gpurun_out/bench_affine_r2e.jsonl 2>> gpurun_out/bench_affine_r2e.err
done
cut -c1-330 gpurun_out/bench_affine_r2e.jsonl; tail -3 gpurun_out/bench_affine_r2e.err
for ch in 1 2 3 4 6; do
  CTT_B200_INPUT_CHUNKS=$ch timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_chunks$ch.json 2> gpurun_out/bench_r2e_chunks$ch.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2e_chunks$ch.json"))
    print("chunks $ch: resident %.3f ms  e2e pinned %.3f ms  pageable %.3f ms  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
except Exception as e:
    print("chunks $ch: failed", e)
PY
done
tail -3 gpurun_out/bench_r2e_chunks4.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; cut -c1-2500 gpurun_out/bench_r2e.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2e.csv > gpurun_out/launches_r2e_summary.txt 2>&1; head -40 gpurun_out/launches_r2e_summary.txt
