#!/bin/bash
# round 2, GPU call N (1 GPU): scalars ahead of the point pieces on ONE copy stream -- e2e sweep over the piece count, twice
mkdir -p gpurun_out
for rep in a b; do
for pc in 1 2 3 4; do
  CTT_B200_POINT_CHUNKS=$pc timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2n_pieces${pc}${rep}.json 2> gpurun_out/bench_r2n_pieces${pc}${rep}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2n_pieces${pc}${rep}.json"))
    print("run $rep point pieces $pc: resident %.3f ms  e2e pinned %.3f ms  pageable %.3f ms  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
except Exception as e:
    print("pieces $pc: failed", e)
PY
done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "input_chunks or device_list or sizes_vs_oracle" > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
