#!/bin/bash
# round 2, GPU call S (1 GPU): where a 2-window shard spends its time (launch list), slice lengths on G2 / Pallas shards
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_shard_r2s.csv \
  python tools/bench_affine.py --levels -1 --win 14:16 --reps 2 > gpurun_out/launches_shard_r2s.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/launches_shard_r2s.csv") if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1000.0 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1000.0)
    k = r[ki][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
for k, (n, t) in agg.items(): print("%8.1f us/launch %4d  %s" % (t / n, n, k))
PY
: > gpurun_out/slice_shards_r2s.jsonl
for k in 16 32 64; do
  timeout 100 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --win 0:2 --slice $k --reps 5 >> gpurun_out/slice_shards_r2s.jsonl 2>> gpurun_out/slice_shards_r2s.err
  timeout 100 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --win 0:5 --slice $k --reps 5 >> gpurun_out/slice_shards_r2s.jsonl 2>> gpurun_out/slice_shards_r2s.err
  timeout 100 python tools/bench_affine.py --curve pallas_ec --logn 22 --levels -1 --win 0:2 --slice $k --reps 5 >> gpurun_out/slice_shards_r2s.jsonl 2>> gpurun_out/slice_shards_r2s.err
  timeout 100 python tools/bench_affine.py --curve bn254_snarks_g1 --logn 20 --levels -1 --win 0:2 --slice $k --reps 5 >> gpurun_out/slice_shards_r2s.jsonl 2>> gpurun_out/slice_shards_r2s.err
done
timeout 100 python tools/bench_affine.py --levels -1 --win 14:16 --reps 5 >> gpurun_out/slice_shards_r2s.jsonl 2>> gpurun_out/slice_shards_r2s.err
timeout 100 python tools/bench_affine.py --curve bls12_381_g2 --logn 18 --levels -1 --win 0:2 --reps 5 >> gpurun_out/slice_shards_r2s.jsonl 2>> gpurun_out/slice_shards_r2s.err
python - <<'PY'
import json
for l in open("gpurun_out/slice_shards_r2s.jsonl"):
    d=json.loads(l)
    print(d["curve"], d["logn"], "windows", d["num_windows"], "c", d["c"], "entries", d["entries"], "K", d["slice_len"], "total %.3f acc %.3f fix %.3f red %.3f tail %.3f" % (d["ms_total"], d["ms_accumulate"], d["ms_fixup"], d["ms_reduce"], d["ms_d2h_tail"]))
PY
tail -2 gpurun_out/slice_shards_r2s.err
