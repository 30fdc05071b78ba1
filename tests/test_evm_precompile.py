"""EIP-2537 BLS12_G1MSM / BLS12_G2MSM through the native wire-format entries ctt_eth_evm_bls12381_g{1,2}msm
(reference constantine/ethereum_evm_precompiles.nim:894-1060; vectors tests/protocol_ethereum_evm_precompiles/eip-2537/
multiexp_G{1,2}_bls.json and fail-multiexp_G{1,2}_bls.json, replayed like the reference's own runner
tests/t_ethereum_evm_precompiles.nim:60-100: byte-exact output on success, any non-success status on the fail vectors)."""
import pytest

EXPECTED_STATUS = {
    "invalid input length": "cttEVM_InvalidInputSize",
    "invalid fp.Element encoding": "cttEVM_IntLargerThanModulus",
    "invalid field element top bytes": "cttEVM_IntLargerThanModulus",
    "invalid point: not on curve": "cttEVM_PointNotOnCurve",
    "g1 point is not on correct subgroup": "cttEVM_PointNotInSubgroup",
    "g2 point is not on correct subgroup": "cttEVM_PointNotInSubgroup",
}


def test_fail_vectors_are_rejected_on_the_host(kat):
    """input validation (sizes, field encodings, curve and subgroup membership) is host code: no GPU needed"""
    from constantine_b200 import msm as M
    assert len(kat["eip2537_fail"]) == 14
    for case in kat["eip2537_fail"]:
        fn = M.eth_evm_bls12381_g1msm if case["group"] == "G1" else M.eth_evm_bls12381_g2msm
        status, _ = fn(bytes.fromhex(case["raw_input"]))
        assert status != "cttEVM_Success", case["name"]
        assert status == EXPECTED_STATUS[case["expected_error"]], (case["name"], status)


def test_output_size_is_checked(kat):
    from constantine_b200 import msm as M
    case = kat["eip2537"][0]
    status, _ = M.eth_evm_bls12381_g1msm(bytes.fromhex(case["raw_input"]), out_len=64)
    assert status == "cttEVM_InvalidOutputSize"


@pytest.mark.gpu
def test_success_vectors_byte_exact(kat):
    from constantine_b200 import msm as M
    for case in kat["eip2537"]:
        fn = M.eth_evm_bls12381_g1msm if case["curve"] == "bls12_381_g1" else M.eth_evm_bls12381_g2msm
        status, out = fn(bytes.fromhex(case["raw_input"]))
        assert status == "cttEVM_Success", case["name"]
        assert out == bytes.fromhex(case["raw_expected"]), case["name"]
