#!/bin/bash
# round 2, GPU call V (1 GPU): block-wide sums in the bit-plane reduction (chains 13 -> 9 additions) -- suite, shapes, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2v.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2v.log
O=gpurun_out/reduce_lanes_r2v.jsonl; : > $O
run() { timeout 120 python tools/bench_affine.py --levels -1 --reps 5 "$@" >> $O 2>> gpurun_out/reduce_lanes_r2v.err; }
run
run --win 14:16
run --win 0:4
run --logn 16
run --logn 18
run --curve bls12_381_g2 --logn 18
run --curve bls12_381_g2 --logn 18 --win 0:2
run --curve bn254_snarks_g1 --win 0:2
run --curve pallas_ec --logn 22 --win 0:2
python - <<'PY'
import json
for l in open("gpurun_out/reduce_lanes_r2v.jsonl"):
    d=json.loads(l)
    print(d["curve"], d["logn"], "windows", d["num_windows"], "c", d["c"], "AL", d["affine_levels"], "K", d["slice_len"], "ok", d["ok"], "total %.3f acc %.3f (aff %.3f) fix %.3f red %.3f tail %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_affine"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"]))
PY
tail -2 gpurun_out/reduce_lanes_r2v.err
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r2v.json
