#!/usr/bin/env python3
"""Build the committed golden fixtures from the reference's own test vectors (run in the dev container, where
/root/reference exists; the fixtures -- not this script -- are what the tests read).

Sources (SURVEY.md section 8c "what pins results"):
  1. reference tests/protocol_ethereum_evm_precompiles/eip-2537/multiexp_G1_bls.json, multiexp_G2_bls.json
     -- byte-pinned BLS12-381 G1/G2 MSM results (EIP-2537 wire format: per pair 128/256 B point || 32 B scalar,
        big-endian, Fp padded to 64 B, Fp2 as c0 || c1; semantics reference constantine/ethereum_evm_precompiles.nim:894-975,
        scalars are reduced mod r before the MSM :948-961).
  2. reference tests/math_elliptic_curves/vectors/tv_<curve>_scalar_mul_<G1|G2>_<bits>bit.json (Sage generated,
     sage/testgen_scalar_mul.sage) -- 40 x ([k]P = Q) per file; the sum of the Q's (computed here with the exact
     big-int tier) turns each file into a 40-term MSM known answer.
Output: tests/golden/msm_kat.json  (one compact file in this repo's own schema).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from constantine_b200.curves import CURVES  # noqa: E402
from oracle import pyref  # noqa: E402

REF = "/root/reference/tests"


def hx(v):
    return "%x" % v


def enc_point(P):
    return None if P is None else [[hx(c) for c in P[0]], [hx(c) for c in P[1]]]


def eip2537(path, curve):
    cases = []
    d = curve.ext_degree
    psize = 128 * d
    r = curve.fr.modulus
    for t in json.load(open(path)):
        raw = bytes.fromhex(t["Input"])
        exp = bytes.fromhex(t["Expected"])
        step = psize + 32
        assert len(raw) % step == 0 and len(exp) == psize
        pts, ks = [], []

        def dec(b):
            co = [int.from_bytes(b[64 * i:64 * (i + 1)], "big") for i in range(2 * d)]
            x, y = tuple(co[:d]), tuple(co[d:])
            return None if all(v == 0 for v in co) else (x, y)

        for i in range(len(raw) // step):
            chunk = raw[i * step:(i + 1) * step]
            pts.append(dec(chunk[:psize]))
            ks.append(int.from_bytes(chunk[psize:], "big") % r)
        want = dec(exp)
        assert all(pyref.on_curve(P, curve) for P in pts) and pyref.on_curve(want, curve)
        assert pyref.msm_naive_fast(ks, pts, curve) == want, t["Name"]  # the exact tier agrees with the reference's vector
        cases.append({"name": t["Name"], "curve": curve.name, "source": os.path.relpath(path, "/root/reference"),
                      "scalars": [hx(k) for k in ks], "points": [enc_point(P) for P in pts], "expected": enc_point(want),
                      "raw_input": t["Input"], "raw_expected": t["Expected"]})
    return cases


def sage_file(path, curve):
    j = json.load(open(path))
    d = curve.ext_degree
    assert int(j["modulus"], 16) == curve.fp.modulus and int(j["order"], 16) == curve.fr.modulus

    def co(v):
        return (int(v, 16),) if d == 1 else (int(v["c0"], 16), int(v["c1"], 16))

    pts, ks, qs = [], [], []
    for v in j["vectors"]:
        P = (co(v["P"]["x"]), co(v["P"]["y"]))
        Q = (co(v["Q"]["x"]), co(v["Q"]["y"]))
        k = int(v["scalar"], 16)
        assert pyref.on_curve(P, curve) and pyref.on_curve(Q, curve)
        assert pyref.ec_mul_fast(k, P, curve) == Q
        pts.append(P), ks.append(k), qs.append(Q)
    total = None
    for Q in qs:
        total = pyref.ec_add(total, Q, curve)
    return {"name": os.path.basename(path)[:-5], "curve": curve.name, "source": os.path.relpath(path, "/root/reference"),
            "scalars": [hx(k) for k in ks], "points": [enc_point(P) for P in pts],
            "products": [enc_point(Q) for Q in qs], "expected": enc_point(total)}


def main():
    out = {"eip2537": [], "eip2537_fail": [], "sage_scalar_mul": []}
    out["eip2537"] += eip2537(f"{REF}/protocol_ethereum_evm_precompiles/eip-2537/multiexp_G1_bls.json", CURVES["bls12_381_g1"])
    out["eip2537"] += eip2537(f"{REF}/protocol_ethereum_evm_precompiles/eip-2537/multiexp_G2_bls.json", CURVES["bls12_381_g2"])
    out["eip2537_fail"] = []
    for g in ("G1", "G2"):
        path = f"{REF}/protocol_ethereum_evm_precompiles/eip-2537/fail-multiexp_{g}_bls.json"
        for t in json.load(open(path)):
            out["eip2537_fail"].append({"name": t["Name"], "group": g, "raw_input": t["Input"], "expected_error": t["ExpectedError"],
                                        "source": os.path.relpath(path, "/root/reference")})
    files = [("BLS12_381", "G1", "bls12_381_g1", (32, 64, 128, 255)), ("BLS12_381", "G2", "bls12_381_g2", (32, 64, 128, 255)),
             ("BN254_Snarks", "G1", "bn254_snarks_g1", (32, 64, 128, 254)), ("BN254_Snarks", "G2", "bn254_snarks_g2", (32, 64, 128, 254)),
             ("Pallas", "G1", "pallas_ec", (255,)), ("Vesta", "G1", "vesta_ec", (255,))]
    for cname, g, ours, bitsets in files:
        for b in bitsets:
            out["sage_scalar_mul"].append(sage_file(f"{REF}/math_elliptic_curves/vectors/tv_{cname}_scalar_mul_{g}_{b}bit.json", CURVES[ours]))
    path = os.path.join(HERE, "msm_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["eip2537"]), "EIP-2537 cases,", len(out["sage_scalar_mul"]), "Sage files")


if __name__ == "__main__":
    main()
