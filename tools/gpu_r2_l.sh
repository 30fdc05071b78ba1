#!/bin/bash
# round 2, GPU call L (1 GPU): points in pieces, level 0 of the affine sums by arrival (ctt_b200_set_point_chunks): full GPU suite + e2e sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_gpu.log
tail -12 gpurun_out/r2l_pytest_gpu.log
for pc in 1 2 4 8; do
  CTT_B200_POINT_CHUNKS=$pc timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2l_pieces$pc.json 2> gpurun_out/bench_r2l_pieces$pc.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2l_pieces$pc.json"))
    print("point pieces $pc: resident %.3f ms  e2e pinned %.3f ms  pageable %.3f ms  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
except Exception as e:
    print("pieces $pc: failed", e)
PY
done
tail -3 gpurun_out/bench_r2l_pieces4.err
timeout 300 python tools/bench_multi_device.py --reps 6 > gpurun_out/multi_device_r2l.jsonl 2> gpurun_out/multi_device_r2l.err; cat gpurun_out/multi_device_r2l.jsonl
