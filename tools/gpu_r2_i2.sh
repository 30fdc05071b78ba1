#!/bin/bash
# round 2, GPU call I2 (2 GPUs): torchrun bench at N = 2 (window digits + one all_gather) and the in-process device list again
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2_2gpu.json"))
    print("2 GPUs: resident %.3f ms (%.1f MSM/s)  e2e pinned %.3f  pageable %.3f  ok=%s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
    print(d["phases_ms_serial_launch_order"])
except Exception as e:
    print("2-GPU bench failed", e)
PY
grep -n "Error" -B2 -A6 gpurun_out/bench_r2_2gpu.err | head -30
timeout 600 python tools/bench_multi_device.py --reps 6 > gpurun_out/multi_device_r2i.jsonl 2> gpurun_out/multi_device_r2i.err; cat gpurun_out/multi_device_r2i.jsonl; tail -2 gpurun_out/multi_device_r2i.err
