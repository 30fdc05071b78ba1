#!/bin/bash
# round 2, GPU call K (1 GPU): paired-product point operations (vs variant without), 5 resident blocks in the pair kernel, Fp2 level
# threshold, staged points after the plan; full GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_gpu.log
tail -5 gpurun_out/r2k_pytest_gpu.log
: > gpurun_out/variants_r2k.jsonl
for v in main nm2 mb5; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/variants_r2k.jsonl
  L=$PWD/constantine_b200/lib/libctt_b200_msm.so; [ $v != main ] && L=$PWD/constantine_b200/lib/libctt_b200_msm_$v.so
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --levels -1 --reps 4 >> gpurun_out/variants_r2k.jsonl 2>> gpurun_out/variants_r2k.err
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --logn 16 --levels -1 --reps 4 >> gpurun_out/variants_r2k.jsonl 2>> gpurun_out/variants_r2k.err
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --logn 18 --levels -1 --reps 4 >> gpurun_out/variants_r2k.jsonl 2>> gpurun_out/variants_r2k.err
  CTT_B200_LIB=$L timeout 300 python tools/bench_affine.py --logn 20 --levels 0 --cs 16 --reps 3 >> gpurun_out/variants_r2k.jsonl 2>> gpurun_out/variants_r2k.err
done
# a window shard as an 8-GPU rank sees it: 2 of 16 windows (forced c = 16 at N = 2^20 is the default plan)
python - <<'PY'
import json
for l in open("gpurun_out/variants_r2k.jsonl"):
    d=json.loads(l)
    if "variant" in d: print("==", d["variant"]); continue
    print(d["curve"], d["logn"], "AL", d["affine_levels"], "c", d["c"], "ok", d["ok"], "total %.3f acc %.3f fix %.3f red %.3f tail %.3f sort %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"], d["ms_sort"]))
PY
tail -2 gpurun_out/variants_r2k.err
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --reps 3 > gpurun_out/g2_r2k.jsonl 2>> gpurun_out/variants_r2k.err; cut -c1-330 gpurun_out/g2_r2k.jsonl
timeout 300 python tools/bench_affine.py --curve bls12_381_g2 --logn 16 --levels -1 --reps 3 >> gpurun_out/g2_r2k.jsonl 2>> gpurun_out/variants_r2k.err; tail -1 gpurun_out/g2_r2k.jsonl | cut -c1-330
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2k.json"))
print("resident %.3f ms  e2e pinned %.3f  pageable %.3f  ok=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
print(d["phases_ms_serial_launch_order"], d["config"]["window_c"]); print(d["cpu_baseline"])
PY
