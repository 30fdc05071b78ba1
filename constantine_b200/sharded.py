"""One MSM split over several GPUs (one process per GPU, torch.distributed).

The reference has no multi-device path; its closest construct is "MSM-level parallelism"
(reference constantine/math/elliptic/ec_multi_scalar_mul_parallel.nim:386-431, msmAffine_vartime_parallel_split):
cut the N (coef, point) pairs into balanced chunks, run one sub-MSM per chunk, add the partial results. Here a chunk
is a rank: every rank runs the single-GPU engine on its shard and the <= world_size partial points (raw XYZZ, a few
hundred bytes each) are exchanged with ONE all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests) and summed on
the host of every rank. An elliptic-curve addition is not an NCCL reduction op, hence gather-then-add.

Window sharding (north star) is the alternative when every rank already holds all N pairs: rank g owns a contiguous
range of windows and returns sum_{w in range} 2^(c w) S_w; the combination step is identical.

A bank of independent MSMs (msm.msm_batch / PrecomputedMSMBank) shards by member: rank g evaluates members
balanced_chunk(batch, world, g) in one engine pass; the only exchange is the all_gather of the result structs.
"""
import ctypes
from typing import Callable, Optional, Tuple

from . import msm as _msm
from .curves import CURVES, CurveParams


def balanced_chunk(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of chunk `rank` when n items are split into `world` chunks whose sizes differ by at most one
    (same contract as the reference's balancedChunksPrioNumber, constantine/threadpool/partitioners.nim:44)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def window_range(num_windows: int, world: int, rank: int) -> Tuple[int, int]:
    return balanced_chunk(num_windows, world, rank)


_xchg_cache = {}


def _all_gather_bytes(payload: bytes, group=None, device=None) -> list:
    """all_gather of one small fixed-size byte string per rank. Buffers (pinned host staging + device tensors for NCCL)
    are created once per (size, device) and reused: the exchange is latency-bound, allocations would dominate it."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = len(payload)
    key = (n, str(device), world, id(group))
    buf = _xchg_cache.get(key)
    if buf is None:
        if device is None:
            buf = (torch.empty(n, dtype=torch.uint8), torch.empty(world * n, dtype=torch.uint8), None, None)
        else:
            buf = (torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(world * n, dtype=torch.uint8).pin_memory(),
                   torch.empty(n, dtype=torch.uint8, device=device), torch.empty(world * n, dtype=torch.uint8, device=device))
        _xchg_cache[key] = buf
    h_in, h_out, d_in, d_out = buf
    h_in.numpy()[:] = memoryview(payload)
    if device is None:
        dist.all_gather_into_tensor(h_out, h_in, group=group) if hasattr(dist, "all_gather_into_tensor") and dist.get_backend(group) != "gloo" \
            else dist.all_gather(list(h_out.view(world, n).unbind(0)), h_in, group=group)
    else:
        d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in, group=group)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
    raw = h_out.numpy().tobytes()
    return [raw[i * n:(i + 1) * n] for i in range(world)]


def combine_partials(curve, partials: list, out=_msm.OUT_JAC) -> bytes:
    cv = curve if isinstance(curve, CurveParams) else CURVES[curve]
    return _msm.sum_partials(cv, b"".join(partials), len(partials), out=out)


def msm_point_sharded(curve, local_partial_xyzz: bytes, group=None, device=None, out=_msm.OUT_JAC) -> bytes:
    """Given this rank's raw XYZZ partial (MSM over its shard of the points), return the full MSM on every rank."""
    parts = _all_gather_bytes(local_partial_xyzz, group=group, device=device)
    return combine_partials(curve, parts, out=out)


def msm_sharded_device(curve, d_coefs: int, d_points: int, n_local: int, group=None, device=None, out=_msm.OUT_JAC,
                       fr_mont=False, windows: Optional[Tuple[int, int]] = None, force_c: int = 0,
                       local_msm: Optional[Callable] = None) -> bytes:
    """Full sharded MSM: local engine call on device-resident shard + all_gather + host combine.

    `windows` = (begin, end) switches to window sharding (d_coefs/d_points then hold ALL pairs).
    `local_msm` lets the CPU tests substitute the local engine call (no GPU there)."""
    cv = curve if isinstance(curve, CurveParams) else CURVES[curve]
    wb, we = windows if windows is not None else (0, -1)
    if local_msm is None:
        part = _msm.msm_device_ptrs(cv, d_coefs, d_points, n_local, out=_msm.OUT_XYZZ, fr_mont=fr_mont, force_c=force_c,
                                    win_begin=wb, win_end=we)
    else:
        part = local_msm(cv, d_coefs, d_points, n_local)
    return msm_point_sharded(cv, part, group=group, device=device, out=out)


_digit_cache = {}


def msm_window_sharded_device(curve, d_coefs: int, d_points: int, n: int, c: int, num_windows: int, group=None, device=None,
                              out=_msm.OUT_JAC, fr_mont=False, local_digits: Optional[Callable] = None) -> bytes:
    """North-star window sharding with ONE exchange and ONE host pass: every rank holds all n pairs, runs the engine for its
    window range and leaves the radix-16 digits of its window sums on the device (no synchronisation, no per-rank host tail);
    one all_gather (NCCL on the engine's stream) collects the <= world x ceil(W / world) x ceil((c-1)/4) digit points, one
    device-to-host copy brings them over, and one Horner pass over the 255 bit positions finishes the MSM on every rank.
    `local_digits(curve, wb, we) -> bytes` lets the CPU tests substitute the engine (gloo, host tensors)."""
    import torch
    import torch.distributed as dist
    cv = curve if isinstance(curve, CurveParams) else CURVES[curve]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    groups = _msm.digits_per_window(c)
    xyzz = 4 * cv.coord_bytes
    per_rank_windows = -(-num_windows // world)
    blk = per_rank_windows * groups * xyzz                     # fixed-size payload; unused slots stay zero = infinity
    wb, we = window_range(num_windows, world, rank)
    key = (blk, world, str(device), id(group))
    bufs = _digit_cache.get(key)
    if bufs is None:
        if device is None:
            bufs = (torch.zeros(blk, dtype=torch.uint8), torch.zeros(world * blk, dtype=torch.uint8), None)
        else:
            bufs = (torch.zeros(blk, dtype=torch.uint8, device=device), torch.zeros(world * blk, dtype=torch.uint8, device=device),
                    torch.zeros(world * blk, dtype=torch.uint8).pin_memory())
        _digit_cache[key] = bufs
    mine, everyone, h_all = bufs
    if local_digits is None:
        # the engine must launch on the stream the collective is issued on: stream order is the only synchronisation
        from . import _lib
        _lib.load().ctt_b200_set_stream(ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _msm.msm_device_digits(cv, mine.data_ptr(), d_coefs, d_points, n, fr_mont=fr_mont, force_c=c, win_begin=wb, win_end=we)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        h_all.copy_(everyone, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        raw = h_all.numpy()
    else:
        payload = local_digits(cv, wb, we)
        mine.zero_()
        mine.numpy()[:len(payload)] = memoryview(payload)
        dist.all_gather(list(everyone.view(world, blk).unbind(0)), mine, group=group)
        raw = everyone.numpy()
    # ranks own consecutive window ranges: concatenating the used part of every block gives the global window-major order
    parts = []
    for g in range(world):
        gb, ge = window_range(num_windows, world, g)
        parts.append(raw[g * blk:g * blk + (ge - gb) * groups * xyzz].tobytes())
    return _msm.combine_window_digits(cv, b"".join(parts), c, num_windows, out=out)


def msm_batch_sharded(curve, coefs, points, batch: int, length: int, group=None, device=None, out=_msm.OUT_JAC,
                      coef_kind="big", shared_points=False, local_batch: Optional[Callable] = None) -> list:
    """`batch` independent MSMs of `length` terms split by member over the ranks of `group`; every rank returns all
    `batch` results. coefs / points are the full host buffers (each rank reads only its members' slices).
    `local_batch(curve, coefs, points, members, length)` lets the CPU tests substitute the engine call."""
    import torch.distributed as dist
    cv = curve if isinstance(curve, CurveParams) else CURVES[curve]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = balanced_chunk(batch, world, rank)
    cmv, pmv = memoryview(coefs).cast("B"), memoryview(points).cast("B")
    c_loc = cmv[lo * length * 32:hi * length * 32]
    p_loc = pmv if shared_points else pmv[lo * length * cv.aff_bytes:hi * length * cv.aff_bytes]
    if local_batch is None:
        res = _msm.msm_batch(cv, c_loc, p_loc, hi - lo, length, out=out, coef_kind=coef_kind, shared_points=shared_points)
    else:
        res = local_batch(cv, c_loc, p_loc, hi - lo, length)
    size = cv.coord_bytes * (4 if out == _msm.OUT_XYZZ else 3)
    per = -(-batch // world)                        # members per rank, rounded up: fixed-size payload for the all_gather
    payload = b"".join(res) + bytes(size * (per - (hi - lo)))
    parts = _all_gather_bytes(payload, group=group, device=device)
    out_list = []
    for g in range(world):
        glo, ghi = balanced_chunk(batch, world, g)
        out_list += [parts[g][i * size:(i + 1) * size] for i in range(ghi - glo)]
    return out_list
