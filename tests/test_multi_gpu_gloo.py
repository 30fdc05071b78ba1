"""CPU, world_size 2 over gloo: the multi-GPU sharding logic (shard -> local partial -> all_gather -> host combine).
The local engine call is replaced by the oracle (no GPU here); everything else is the product code path."""
import os
import socket
import sys

import pytest

from helpers import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import random
    import torch.distributed as dist
    from helpers import CURVES, affine_to_xyzz_bytes, pack, point_pool, pyref
    from constantine_b200 import sharded, msm as M
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cv = CURVES["bls12_381_g1"]
        rnd = random.Random(123)  # same instance on every rank
        _, pool = point_pool(cv)
        n = 301
        pts = [pool[rnd.randrange(len(pool))] for _ in range(n)]
        ks = [rnd.getrandbits(255) for _ in range(n)]
        want = pyref.msm_naive_fast(ks, pts, cv)
        if mode == "points":
            lo, hi = sharded.balanced_chunk(n, world, rank)
            cb, pb = pack(cv, ks[lo:hi], pts[lo:hi])

            def local(cvv, d_coefs, d_points, n_local):  # stands in for ctt_b200_msm_device(..., OUT_XYZZ)
                aff = pyref.jac_bytes_to_affine(oracle.msm(cvv, cb, pb, n_local), cvv)
                return affine_to_xyzz_bytes(aff, cvv)

            got = sharded.msm_sharded_device(cv, 0, 0, hi - lo, local_msm=local)
        elif mode == "bank":   # a bank of 5 MSMs x 9 terms split by member (3 + 2); every rank ends with all 5 results
            batch, m = 5, 9
            bks = [rnd.getrandbits(255) for _ in range(batch * m)]
            bpts = [pool[rnd.randrange(len(pool))] for _ in range(batch * m)]
            cb, pb = pack(cv, bks, bpts)

            def local_batch(cvv, c_loc, p_loc, members, length):  # stands in for ctt_b200_msm_batch_host
                return [oracle.msm(cvv, bytes(c_loc[i * length * 32:(i + 1) * length * 32]),
                                   bytes(p_loc[i * length * cvv.aff_bytes:(i + 1) * length * cvv.aff_bytes]), length)
                        for i in range(members)]

            res = sharded.msm_batch_sharded(cv, cb, pb, batch, m, local_batch=local_batch)
            wants = [pyref.msm_naive_fast(bks[i * m:(i + 1) * m], bpts[i * m:(i + 1) * m], cv) for i in range(batch)]
            q.put((rank, [pyref.jac_bytes_to_affine(r, cv) for r in res] == wants))
            return
        elif mode == "window_digits":   # north-star window sharding: digits of the window sums, one exchange, one host pass
            c = 8
            W = 255 // c + 1
            groups = M.digits_per_window(c)

            def local_digits(cvv, wb, we):   # stands in for ctt_b200_msm_device_digits: D_0 = S_w, the other digits empty
                out = b""
                for w in range(wb, we):
                    sw = None
                    for k, P in zip(ks, pts):
                        val, neg = oracle.signed_digit(k, 255, c, w)
                        if val:
                            term = pyref.ec_mul_fast(val, P, cvv)
                            sw = pyref.ec_add(sw, pyref.ec_neg(term, cvv) if neg else term, cvv)
                    out += affine_to_xyzz_bytes(sw, cvv) + bytes(4 * cvv.coord_bytes) * (groups - 1)
                return out

            got = sharded.msm_window_sharded_device(cv, 0, 0, n, c, W, local_digits=local_digits)
        else:  # window sharding: every rank holds all pairs, owns a window range, returns sum 2^(cw) S_w over its range
            c = 8
            W = 255 // c + 1
            wb, we = sharded.window_range(W, world, rank)
            mask_ks = []
            for k in ks:  # scalar restricted to the rank's signed digits: sum_{w in range} d_w 2^(cw)
                tot = 0
                for w in range(wb, we):
                    val, neg = oracle.signed_digit(k, 255, c, w)
                    tot += (-val if neg else val) << (w * c)
                mask_ks.append(tot)
            part = None
            for k, P in zip(mask_ks, pts):
                term = pyref.ec_mul_fast(abs(k), P, cv)
                part = pyref.ec_add(part, pyref.ec_neg(term, cv) if k < 0 else term, cv)
            got = sharded.msm_point_sharded(cv, affine_to_xyzz_bytes(part, cv))
        q.put((rank, pyref.jac_bytes_to_affine(got, cv) == want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["points", "windows", "window_digits", "bank"])
def test_two_rank_sharded_msm(mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
