#!/usr/bin/env python3
"""Extract the reference's blob_to_kzg_commitment known answers as a 4096-term BLS12-381 G1 MSM fixture.

Sources (dev container only):
  reference constantine/commitments_setups/trusted_setup_ethereum_kzg4844_reference.dat   (c-kzg text format: "4096\\n65\\n",
      then 4096 compressed G1 points = the Lagrange-form SRS; loaded and bit-reversal-permuted by
      constantine/commitments_setups/ethereum_kzg_srs.nim:299-384)
  reference tests/protocol_ethereum_eip4844_deneb_kzg/blob_to_kzg_commitment/kzg-mainnet/*valid*/data.yaml
      (blob = 4096 x 32-byte big-endian Fr elements, output = 48-byte compressed commitment)
commitment = sum_i blob[i] * srs_lagrange_brp[i]   (reference constantine/commitments/kzg.nim:186, parallel: kzg_parallel.nim:42)
Output: tests/golden/kzg_commit_kat.npz  (SRS kept compressed: 4096 x 48 B; all seven valid blobs, densest first; their expected commitments).
"""
import glob
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from constantine_b200.curves import CURVES  # noqa: E402
from oracle import pyref  # noqa: E402

REF = "/root/reference"
cv = CURVES["bls12_381_g1"]


def brp(n, bits):
    return int(bin(n)[2:].zfill(bits)[::-1], 2)


def main():
    lines = open(f"{REF}/constantine/commitments_setups/trusted_setup_ethereum_kzg4844_reference.dat").read().split()
    n1, n2 = int(lines[0]), int(lines[1])
    assert (n1, n2) == (4096, 65)
    srs = [bytes.fromhex(h) for h in lines[2:2 + n1]]
    srs_brp = [srs[brp(i, 12)] for i in range(n1)]          # bit-reversal permutation, as the reference does after loading
    pts = [pyref.bls12_381_g1_decompress(b, cv) for b in srs_brp]
    assert all(pyref.on_curve(P, cv) for P in pts)
    blobs, outs, names = [], [], []
    for d in sorted(glob.glob(f"{REF}/tests/protocol_ethereum_eip4844_deneb_kzg/blob_to_kzg_commitment/kzg-mainnet/*valid_blob*")):
        if "invalid" in d:
            continue
        y = open(os.path.join(d, "data.yaml")).read()
        blob = bytes.fromhex(re.search(r"blob: '0x([0-9a-f]+)'", y).group(1))
        out = bytes.fromhex(re.search(r"output: '0x([0-9a-f]+)'", y).group(1))
        assert len(blob) == 4096 * 32 and len(out) == 48
        ks = [int.from_bytes(blob[32 * i:32 * i + 32], "big") for i in range(4096)]
        nz = sum(1 for k in ks if k)
        # keep the vector only if the exact tier reproduces the reference's answer (it must)
        got = pyref.bls12_381_g1_compress(pyref.msm_naive_fast(ks, pts, cv), cv)
        assert got == out, d
        print(os.path.basename(d), "non-zero scalars:", nz, "ok")
        blobs.append(np.frombuffer(blob, dtype=np.uint8)); outs.append(np.frombuffer(out, dtype=np.uint8)); names.append(os.path.basename(d))
    # all of them, densest first (the sparse ones compress to almost nothing); record how many the exact tier verified
    order = sorted(range(len(blobs)), key=lambda i: -int(np.count_nonzero(blobs[i])))
    np.savez_compressed(os.path.join(HERE, "kzg_commit_kat.npz"),
                        srs_lagrange_brp_compressed=np.frombuffer(b"".join(srs_brp), dtype=np.uint8).reshape(4096, 48),
                        blobs=np.stack([blobs[i] for i in order]), commitments=np.stack([outs[i] for i in order]),
                        names=np.array([names[i] for i in order]), verified_cases=np.array(len(blobs)))
    print("wrote kzg_commit_kat.npz", os.path.getsize(os.path.join(HERE, "kzg_commit_kat.npz")), "bytes; verified", len(blobs), "reference cases")


if __name__ == "__main__":
    main()
