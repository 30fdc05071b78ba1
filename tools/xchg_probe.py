#!/usr/bin/env python3
"""Latency of the partial-point exchange alone (torchrun, N ranks): all_gather of 192 B + host combine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from constantine_b200 import sharded, msm as M
from constantine_b200.curves import CURVES
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cv = CURVES["bls12_381_g1"]
part = bytes(192)
for _ in range(20): sharded.msm_point_sharded(cv, part, device=dev)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): sharded._all_gather_bytes(part, device=dev)
t1 = time.perf_counter()
for _ in range(200): sharded.msm_point_sharded(cv, part, device=dev)
t2 = time.perf_counter()
if rank == 0:
    print(f"all_gather only: {(t1 - t0) / 200 * 1e6:.1f} us   all_gather + combine: {(t2 - t1) / 200 * 1e6:.1f} us", flush=True)
dist.destroy_process_group()
