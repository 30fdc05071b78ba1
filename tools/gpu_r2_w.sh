#!/bin/bash
# round 2, GPU call W (1 GPU): final state -- suite, the 8-word shard shapes after the lane cap, bench line, smoke
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2w.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2w.log
O=gpurun_out/reduce_lanes_r2w.jsonl; : > $O
run() { timeout 60 python tools/bench_affine.py --levels -1 --reps 5 "$@" >> $O 2>> gpurun_out/reduce_lanes_r2w.err; }
run --curve bn254_snarks_g1 --win 0:2
run --curve pallas_ec --logn 22 --win 0:2
python - <<'PY'
import json
for l in open("gpurun_out/reduce_lanes_r2w.jsonl"):
    d=json.loads(l)
    print(d["curve"], d["logn"], "windows", d["num_windows"], "c", d["c"], "AL", d["affine_levels"], "K", d["slice_len"], "ok", d["ok"], "total %.3f acc %.3f (aff %.3f) fix %.3f red %.3f tail %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_affine"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"]))
PY
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r2w.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
