#!/bin/bash
# round 2, GPU call Q (1 GPU): slice length K of k_accumulate on the latency-bound shapes (window shards of a multi-GPU run, N = 2^16, 2^18)
mkdir -p gpurun_out
: > gpurun_out/slice_sweep_r2q.jsonl
for k in 0 16 32 64; do
  timeout 200 python tools/bench_affine.py --levels -1 --win 14:16 --slice $k --reps 5 >> gpurun_out/slice_sweep_r2q.jsonl 2>> gpurun_out/slice_sweep_r2q.err
  timeout 200 python tools/bench_affine.py --levels -1 --win 0:4 --slice $k --reps 5 >> gpurun_out/slice_sweep_r2q.jsonl 2>> gpurun_out/slice_sweep_r2q.err
  timeout 200 python tools/bench_affine.py --logn 16 --levels -1 --slice $k --reps 5 >> gpurun_out/slice_sweep_r2q.jsonl 2>> gpurun_out/slice_sweep_r2q.err
  timeout 200 python tools/bench_affine.py --logn 18 --levels -1 --slice $k --reps 5 >> gpurun_out/slice_sweep_r2q.jsonl 2>> gpurun_out/slice_sweep_r2q.err
  timeout 200 python tools/bench_affine.py --logn 18 --levels 0 --cs 15 --slice $k --reps 5 >> gpurun_out/slice_sweep_r2q.jsonl 2>> gpurun_out/slice_sweep_r2q.err
done
python - <<'PY'
import json
for l in open("gpurun_out/slice_sweep_r2q.jsonl"):
    d=json.loads(l)
    print(d["curve"], d["logn"], "windows", d["num_windows"], "c", d["c"], "K", d["slice_len"], "ok", d["ok"], "total %.3f acc %.3f fix %.3f red %.3f tail %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"]))
PY
tail -2 gpurun_out/slice_sweep_r2q.err
