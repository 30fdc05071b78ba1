#!/bin/bash
# round 2, GPU call P (1 GPU): reduce-phase variants on the shapes that are latency-bound -- a 2-window shard (the 8-GPU rank), N = 2^16, 2^18 --
# main (out-of-line multiplier calls, paired products) vs rolled inline multiplier vs 4x more lanes per sum
mkdir -p gpurun_out
: > gpurun_out/reduce_variants_r2p.jsonl
for v in main rol rt; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/reduce_variants_r2p.jsonl
  L=$PWD/constantine_b200/lib/libctt_b200_msm.so; [ $v != main ] && L=$PWD/constantine_b200/lib/libctt_b200_msm_$v.so
  CTT_B200_LIB=$L timeout 200 python tools/bench_affine.py --levels -1 --win 14:16 --reps 5 >> gpurun_out/reduce_variants_r2p.jsonl 2>> gpurun_out/reduce_variants_r2p.err
  CTT_B200_LIB=$L timeout 200 python tools/bench_affine.py --levels -1 --win 0:4 --reps 5 >> gpurun_out/reduce_variants_r2p.jsonl 2>> gpurun_out/reduce_variants_r2p.err
  CTT_B200_LIB=$L timeout 200 python tools/bench_affine.py --logn 16 --levels -1 --reps 5 >> gpurun_out/reduce_variants_r2p.jsonl 2>> gpurun_out/reduce_variants_r2p.err
  CTT_B200_LIB=$L timeout 200 python tools/bench_affine.py --logn 18 --levels -1 --reps 5 >> gpurun_out/reduce_variants_r2p.jsonl 2>> gpurun_out/reduce_variants_r2p.err
  CTT_B200_LIB=$L timeout 200 python tools/bench_affine.py --levels -1 --reps 4 >> gpurun_out/reduce_variants_r2p.jsonl 2>> gpurun_out/reduce_variants_r2p.err
  CTT_B200_LIB=$L timeout 200 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --win 0:3 --reps 3 >> gpurun_out/reduce_variants_r2p.jsonl 2>> gpurun_out/reduce_variants_r2p.err
done
python - <<'PY'
import json
for l in open("gpurun_out/reduce_variants_r2p.jsonl"):
    d=json.loads(l)
    if "variant" in d: print("==", d["variant"]); continue
    print(d["curve"], d["logn"], "windows", d["num_windows"], "AL", d["affine_levels"], "c", d["c"], "ok", d["ok"], "total %.3f acc %.3f fix %.3f red %.3f tail %.3f sort %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"], d["ms_sort"]))
PY
tail -2 gpurun_out/reduce_variants_r2p.err
