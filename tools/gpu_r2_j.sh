#!/bin/bash
# round 2, GPU call J (8 GPUs): scaling points of BASELINE configs 3, 4, 5 (row g of the round-1 verdict) + the in-process device list
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt
run() {  # curve logn ngpu
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $((29700 + $3)) bench.py --gpus $3 --steps 8 --warmup 3 --curve $1 --logn $2 --no-cpu-baseline > gpurun_out/bench_r2_${1}_${2}_${3}gpu.json 2> gpurun_out/bench_r2_${1}_${2}_${3}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2_${1}_${2}_${3}gpu.json"))
    print("$1 2^$2 x$3: resident %.3f ms (%.1f MSM/s)  e2e pinned %.3f  ok=%s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["closed_form_check"]))
except Exception as e:
    print("$1 2^$2 x$3 failed", e)
PY
}
run bls12_381_g1 20 8
run bls12_381_g1 20 4
run pallas_ec 22 8
run bls12_381_g2 18 8
run bls12_381_g2 18 4
run bls12_381_g2 18 2
timeout 240 python tools/bench_multi_device.py --reps 6 > gpurun_out/multi_device_r2j.jsonl 2> gpurun_out/multi_device_r2j.err; cat gpurun_out/multi_device_r2j.jsonl; tail -2 gpurun_out/multi_device_r2j.err
