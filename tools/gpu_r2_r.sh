#!/bin/bash
# round 2, GPU call R (1 GPU): final validation of the shipped defaults -- full GPU suite, the bench line, smoke, the CPU arm,
# and the window-shard shape with the automatic slice length
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_final.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2_final.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r2_final.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
: > gpurun_out/slice_auto_r2r.jsonl
timeout 100 python tools/bench_affine.py --levels -1 --win 14:16 --reps 5 >> gpurun_out/slice_auto_r2r.jsonl 2>> gpurun_out/slice_auto_r2r.err
timeout 100 python tools/bench_affine.py --levels -1 --win 0:2 --reps 5 >> gpurun_out/slice_auto_r2r.jsonl 2>> gpurun_out/slice_auto_r2r.err
timeout 100 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --win 0:2 --reps 5 >> gpurun_out/slice_auto_r2r.jsonl 2>> gpurun_out/slice_auto_r2r.err
timeout 100 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --win 0:2 --slice 16 --reps 5 >> gpurun_out/slice_auto_r2r.jsonl 2>> gpurun_out/slice_auto_r2r.err
python - <<'PY'
import json
for l in open("gpurun_out/slice_auto_r2r.jsonl"):
    d=json.loads(l)
    print(d["curve"], d["logn"], "windows", d["num_windows"], "c", d["c"], "K", d["slice_len"], "total %.3f acc %.3f fix %.3f red %.3f tail %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"]))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err
echo "ref rc=$?"; cut -c1-500 gpurun_out/bench_r2_final_reference.json
