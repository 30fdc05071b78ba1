#!/bin/bash
# round 2, GPU call M (1 GPU): final validation -- full GPU suite, sanitizer over the batched-affine + point-piece paths, bench line, ncu
# launch list and one `--set full` capture of the accumulation phase (3 pair-kernel launches + k_accumulate) and of the reduce kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.log
tail -5 gpurun_out/r2m_pytest_gpu.log
CTT_B200_POINT_CHUNKS=3 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2 > gpurun_out/r2m_sanitize.log 2>&1; tail -4 gpurun_out/r2m_sanitize.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2m.json"))
print("resident %.3f ms  e2e pinned %.3f  pageable %.3f  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
print(d["phases_ms_serial_launch_order"], d["config"]["window_c"]); print(d["roofline"]["frac"], d["roofline"]["frac_executed"])
PY
CTT_B200_POINT_CHUNKS=3 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2m_pieces3.json 2> /dev/null
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2m_pieces3.json")); print("pieces 3: e2e pinned %.3f pageable %.3f" % (d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"]))
except Exception as e: print("pieces3 failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r2m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2m.csv > gpurun_out/launches_r2m_summary.txt 2>&1; head -32 gpurun_out/launches_r2m_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_affine_pairs|k_accumulate" -c 4 -o gpurun_out/ncu_accumulate_phase_r2m python tools/bench_affine.py --levels 3 --reps 1 > gpurun_out/r2m_ncu1.log 2>&1; tail -2 gpurun_out/r2m_ncu1.log
timeout 600 ncu --set full --clock-control none -k regex:"k_rowcol_sums|k_plane_sums|k_plane_combine" -c 3 -o gpurun_out/ncu_reduce_r2m python tools/bench_affine.py --levels 3 --reps 1 > gpurun_out/r2m_ncu2.log 2>&1; tail -2 gpurun_out/r2m_ncu2.log
