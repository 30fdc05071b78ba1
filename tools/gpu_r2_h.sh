#!/bin/bash
# round 2, GPU call H (1 GPU): Fp2 kernel variants, per-config table with the retuned window / affine-level choice, bench line
mkdir -p gpurun_out
: > gpurun_out/g2_variants_r2h.jsonl
for v in main g2a g2b; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/g2_variants_r2h.jsonl
  L=$PWD/constantine_b200/lib/libctt_b200_msm.so; [ $v != main ] && L=$PWD/constantine_b200/lib/libctt_b200_msm_$v.so
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels 0,4 --reps 3 >> gpurun_out/g2_variants_r2h.jsonl 2>> gpurun_out/g2_variants_r2h.err
done
cut -c1-330 gpurun_out/g2_variants_r2h.jsonl; tail -2 gpurun_out/g2_variants_r2h.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs_r2h.jsonl 2> gpurun_out/configs_r2h.err; cut -c1-300 gpurun_out/configs_r2h.jsonl; tail -2 gpurun_out/configs_r2h.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closed_form or sizes_vs_oracle or forced_window" > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2h.json"))
print("resident %.3f ms  e2e pinned %.3f  pageable %.3f  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
print(d["phases_ms_serial_launch_order"], d["config"]["window_c"])
PY
