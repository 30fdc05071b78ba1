#!/bin/bash
# round 2, GPU call O (4 GPUs): configs 4 and 5 at 2 / 4 GPUs with the final build (Fp2 affine levels on for the window shards), config 3 at 4
mkdir -p gpurun_out
run() {  # curve logn ngpu
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $((29800 + $3)) bench.py --gpus $3 --steps 8 --warmup 3 --curve $1 --logn $2 --no-cpu-baseline > gpurun_out/bench_r2final_${1}_${2}_${3}gpu.json 2> gpurun_out/bench_r2final_${1}_${2}_${3}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2final_${1}_${2}_${3}gpu.json"))
    print("$1 2^$2 x$3: resident %.3f ms (%.1f MSM/s)  e2e pinned %.3f  ok=%s  phases %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["closed_form_check"], d["phases_ms_serial_launch_order"]))
except Exception as e:
    print("$1 2^$2 x$3 failed", e)
PY
}
run bls12_381_g2 18 2
run bls12_381_g2 18 4
run bls12_381_g1 20 4
run bls12_381_g1 20 2
run pallas_ec 22 4
