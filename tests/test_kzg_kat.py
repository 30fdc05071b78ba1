"""4096-term BLS12-381 G1 MSM known answers: the reference's blob_to_kzg_commitment vectors
(reference tests/protocol_ethereum_eip4844_deneb_kzg/blob_to_kzg_commitment/kzg-mainnet/, SRS
constantine/commitments_setups/trusted_setup_ethereum_kzg4844_reference.dat; the MSM is the whole computation:
constantine/commitments/kzg.nim:186 / kzg_parallel.nim:42).  Fixture: tests/golden/kzg_commit_kat.npz."""
import os

import numpy as np
import pytest

from helpers import CURVES, ROOT, pyref


@pytest.fixture(scope="module")
def kzg():
    z = np.load(os.path.join(ROOT, "tests", "golden", "kzg_commit_kat.npz"))
    cv = CURVES["bls12_381_g1"]
    pts = [pyref.bls12_381_g1_decompress(bytes(row), cv) for row in z["srs_lagrange_brp_compressed"]]
    pb = b"".join(pyref.aff_to_bytes(P, cv) for P in pts)
    cases = []
    for blob, out in zip(z["blobs"], z["commitments"]):
        ks = [int.from_bytes(bytes(blob[32 * i:32 * i + 32]), "big") for i in range(4096)]
        cases.append((ks, bytes(out)))
    assert int(z["verified_cases"]) == 7
    return cv, pb, cases


def test_oracle_reproduces_kzg_commitments(kzg, oracle_lib):
    cv, pb, cases = kzg
    for ks, want in cases:
        cb = b"".join(k.to_bytes(32, "little") for k in ks)
        got = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, 4096), cv)
        assert pyref.bls12_381_g1_compress(got, cv) == want


@pytest.mark.gpu
def test_cuda_path_reproduces_kzg_commitments(kzg):
    from constantine_b200 import msm as M
    cv, pb, cases = kzg
    tp = M.Threadpool.new(1)
    for ks, want in cases:
        cb = b"".join(k.to_bytes(32, "little") for k in ks)
        got = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, 4096, out="jac", coef_kind="big"), cv)
        assert pyref.bls12_381_g1_compress(got, cv) == want
        cbm = b"".join(cv.fr.to_mont(k).to_bytes(32, "little") for k in ks)   # Fr Montgomery residues, as kzg_parallel.nim passes them
        got = pyref.prj_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cbm, pb, 4096, out="prj", coef_kind="fr"), cv)
        assert pyref.bls12_381_g1_compress(got, cv) == want
    tp.shutdown()


@pytest.mark.gpu
def test_blob_to_kzg_commitment_entry(kzg):
    """The reference's blob_to_kzg_commitment vectors byte for byte through the commitment entry on the resident SRS
    (ctt_b200_eth_kzg_*): blob bytes in, 48-byte compressed commitment out; context from the compressed trusted-setup points
    and from affine Montgomery structs, with and without the precomputed window table; a blob element >= r is refused with
    the reference's status (reference tests: blob_to_kzg_commitment/kzg-mainnet/*invalid_blob*)."""
    from constantine_b200 import msm as M
    cv, pb, cases = kzg
    z = np.load(os.path.join(ROOT, "tests", "golden", "kzg_commit_kat.npz"))
    for ctx in (M.EthKzgContext(z["srs_lagrange_brp_compressed"].tobytes(), compressed=True), M.EthKzgContext(pb, compressed=False)):
        for table in (False, True):
            if table:
                assert ctx.precompute(0) >= 2
            for blob, want in zip(z["blobs"], z["commitments"]):
                assert ctx.blob_to_kzg_commitment(blob.tobytes()) == bytes(want)
        bad = bytearray(z["blobs"][0].tobytes())
        bad[32 * 5:32 * 6] = cv.fr.modulus.to_bytes(32, "big")
        with pytest.raises(ValueError) as e:
            ctx.blob_to_kzg_commitment(bytes(bad))
        assert e.value.args[0] == M.EthKzgContext.ScalarLargerThanCurveOrder
        zero = bytes(4096 * 32)                              # the zero polynomial commits to the point at infinity
        assert ctx.blob_to_kzg_commitment(zero) == bytes([0xC0]) + bytes(47)
        ctx.delete()
