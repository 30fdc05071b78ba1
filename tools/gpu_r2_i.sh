#!/bin/bash
# round 2, GPU call I (2 GPUs): the multi-device tests on two real devices, the torchrun bench at N = 2, the in-process device list
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2i_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "device_list or caller_thread or multi_gpu or window_digits" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -4 gpurun_out/r2i_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2_2gpu.json"))
    print("2 GPUs: resident %.3f ms (%.1f MSM/s)  e2e pinned %.3f  pageable %.3f  ok=%s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["closed_form_check"]))
except Exception as e:
    print("2-GPU bench failed", e)
PY
tail -3 gpurun_out/bench_r2_2gpu.err
timeout 600 python tools/bench_multi_device.py --reps 6 > gpurun_out/multi_device_r2i.jsonl 2> gpurun_out/multi_device_r2i.err; cat gpurun_out/multi_device_r2i.jsonl; tail -2 gpurun_out/multi_device_r2i.err
