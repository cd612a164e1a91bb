"""The .bed body (plink_interface.cpp:10-35, 67-103, 255-284).

CPU: the numpy restatement in oracle/ and the swap-flag generator of the host mirror are pinned to the bytes
the unmodified reference wrote (tests/golden/bed_toy_chr22_K10_res2.npz, made by make_golden_bed.py).
GPU: the device transpose reproduces those bytes and the restatement on ragged shapes."""
import numpy as np
import pytest

import kgw_oracle as ko
import knockoffgwas_b200 as kg
from kgw_testutil import GOLDEN


def _fixture():
    f = np.load(GOLDEN / "bed_toy_chr22_K10_res2.npz")
    return f, int(f["p"])


def test_oracle_bed_matches_reference_file():
    f, p = _fixture()
    assert np.array_equal(ko.bed_body(f["H"], f["Hk"], p, f["swap"]), f["bed_body"])


def test_swap_flags_match_reference_bim():
    """mt19937_64(2019) + uniform_int_distribution<int>(0,1) as libstdc++ evaluates it (plink_interface.cpp:267-274)."""
    f, p = _fixture()
    assert np.array_equal(kg.swap_flags(p), f["swap"])


def test_oracle_bed_layout_small():
    """Hand-checked codes of writeBED_variant (:17-33): 5 samples -> 2 bytes, high bit for heterozygotes."""
    bits = np.array([[1, 0], [1, 0], [1, 1], [0, 1], [0, 0], [0, 0], [0, 1], [1, 1], [1, 0], [1, 1]], dtype=np.uint8)  # 10 haps x 2 SNPs
    H = ko.pack_bits(bits)
    body = ko.bed_body(None, H, 2)
    # SNP 0: samples (1,1)->11 (1,0)->10 (0,0)->00 (0,1)->10 | (1,1)->11  => bytes 0b10001011, 0b00000011
    assert body[0].tolist() == [0b10001011, 0b00000011]
    # SNP 1: (0,0)->00 (1,1)->11 (0,0)->00 (1,1)->11 | (0,1)->10
    assert body[1].tolist() == [0b11001100, 0b00000010]


@pytest.mark.gpu
def test_device_bed_matches_reference_file():
    f, p = _fixture()
    body = kg.bed_encode(f["H"], f["Hk"], p)          # swap flags drawn by the host mirror
    assert np.array_equal(body, f["bed_body"])
    tilde = kg.bed_encode(None, f["Hk"], p)
    assert np.array_equal(tilde, ko.bed_body(None, f["Hk"], p))


@pytest.mark.gpu
@pytest.mark.parametrize("N,p", [(2, 1), (10, 2), (254, 127), (256, 128), (258, 129), (770, 333), (4098, 1000),
                                 (2064, 40), (3300, 21), (8210, 70), (20000, 50), (2296, 65)])   # columns of several 128-byte tiles at every alignment
def test_device_bed_ragged_shapes(N, p):
    rng = np.random.default_rng(N * 1000 + p)
    H = ko.pack_bits((rng.random((N, p)) < 0.4).astype(np.uint8))
    Hk = ko.pack_bits((rng.random((N, p)) < 0.4).astype(np.uint8))
    sw = (rng.random(p) < 0.5).astype(np.uint8)
    assert np.array_equal(kg.bed_encode(H, Hk, p, swap=sw), ko.bed_body(H, Hk, p, sw))
    assert np.array_equal(kg.bed_encode(None, Hk, p), ko.bed_body(None, Hk, p))


@pytest.mark.gpu
def test_plan_bed_uses_resident_matrices():
    from kgw_testutil import synthetic_problem
    prob = synthetic_problem(N=300, p=210, K=10, mean_group=2.0, seed=9)
    plan = kg.Plan(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, seed=prob.seed)
    plan.run()
    hk, _ = plan.download()
    sw = kg.swap_flags(prob.p)
    body, ms = plan.encode_bed()
    tilde, _ = plan.encode_bed(include_genotypes=False)
    plan.close()
    assert ms >= 0.0
    assert np.array_equal(body, ko.bed_body(prob.H, hk, prob.p, sw))
    assert np.array_equal(tilde, ko.bed_body(None, hk, prob.p))
