#!/bin/bash
# round 2, GPU call G (1 GPU): window-digits path + pageable staging + multi-device C caller in the GPU suite; window-size sweeps with the
# new reduction; per-config table; bench line with the staged pageable leg
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_gpu.log
tail -6 gpurun_out/r2g_pytest_gpu.log
: > gpurun_out/sweep_c_r2g.jsonl
timeout 300 python tools/bench_affine.py --logn 16 --levels 0 --cs 9,10,11,12,13,14 --reps 4 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --logn 16 --levels 2 --cs 11,12,13 --reps 4 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --logn 18 --levels 0 --cs 12,13,14,15 --reps 4 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --logn 18 --levels 3 --cs 12,13,14,15 --reps 4 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --logn 20 --levels 3 --cs 15,16,17 --reps 3 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --logn 14 --levels 0 --cs 8,9,10,11,12 --reps 4 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 4 --cs 12,13,14 --reps 3 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
timeout 300 python tools/bench_affine.py --curve pallas_ec --logn 22 --levels 0,3 --cs 16,17,18 --reps 2 >> gpurun_out/sweep_c_r2g.jsonl 2>> gpurun_out/sweep_c_r2g.err
python - <<'PY'
import json
for l in open("gpurun_out/sweep_c_r2g.jsonl"):
    d=json.loads(l); print(d["curve"], d["logn"], "AL", d["affine_levels"], "c", d["c"], "ok", d["ok"], "total %.3f acc %.3f fix %.3f red %.3f tail %.3f sort %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"], d["ms_sort"]))
PY
tail -3 gpurun_out/sweep_c_r2g.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2g.json"))
print("resident %.3f ms  e2e pinned %.3f  pageable %.3f  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
print(d["phases_ms_serial_launch_order"]); print(d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench_r2g.err
