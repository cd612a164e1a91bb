"""Pins the CPU oracle to the UNMODIFIED reference.

The fixtures under tests/golden/ were written by the reference binary itself
(oracle/_ref/snpknock2_probe, --n_threads 1).  With one thread the reference
feeds the unrelated haplotypes from a default-constructed boost taus88
(knockoffs.cpp:622-624), consumed haplotype-major in call order.  Replaying that
stream through oracle/kgw_oracle.c must reproduce the reference's latent paths
and haplotype-level knockoffs bit for bit (integer/bit work: exact equality).
"""
import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import GOLDEN_NAMES, load_golden


def test_golden_fixtures_present():
    assert len(GOLDEN_NAMES) >= 8


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_reproduces_reference_unrelated(name):
    prob, f = load_golden(name)
    res = ko.run_unrelated(prob, mode="taus88", taus_seed=None)
    assert np.array_equal(res.z, f["z_unrelated"].astype(np.int32)), "latent path Z differs from reference"
    assert np.array_equal(res.zk, f["zk_unrelated"].astype(np.int32)), "knockoff path Zk differs from reference"
    assert np.array_equal(res.hk, f["Hk"][prob.unrelated]), "knockoff alleles differ from reference"


def test_taus88_matches_boost_semantics():
    # default-constructed engine = seed 341 (linear_feedback_shift.hpp: default_seed)
    a = ko.taus88_stream(None, 5)
    b = ko.taus88_stream(341, 5)
    assert np.array_equal(a, b)
    assert len(set(a.tolist())) == 5
    # the low (32-k) state bits are discarded by the first step, so seeds 0 and 1 coincide
    # (linear_feedback_shift.hpp:90-98, :117-123); distinct high bits give distinct streams
    assert np.array_equal(ko.taus88_stream(0, 3), ko.taus88_stream(1, 3))
    assert not np.array_equal(ko.taus88_stream(2020, 3), ko.taus88_stream(2021 + 64, 3))


def test_philox_known_answers():
    # Random123 known-answer vectors for Philox4x32-10
    import ctypes as C
    L = ko.lib()
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        L.kgw_oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(x) for x in o) == want


def test_philox_numpy_matches_c():
    U = ko.philox_uniforms(2020, np.array([0, 7, 1999]), 37)
    for a, hap in enumerate((0, 7, 1999)):
        for s in range(3):
            for j in (0, 1, 2, 3, 4, 35, 36):
                assert U[a, s, j] == ko.philox_uniform(2020, hap, s, j)
    assert (U >= 0).all() and (U < 1).all()


def test_weighted_choice_semantics():
    import ctypes as C
    L = ko.lib()
    w = np.array([0.25, 0.0, 0.5, 0.25])
    pick = lambda u: L.kgw_oracle_weighted_choice(w.ctypes.data, len(w), C.c_double(u))
    assert pick(0.0) == 0
    assert pick(0.2499) == 0
    assert pick(0.25) == 2      # zero-weight entry is skipped (R < 0 is false)
    assert pick(0.75) == 3
    assert pick(0.999999) == 3
    z = np.zeros(3)
    assert L.kgw_oracle_weighted_choice(z.ctypes.data, 3, C.c_double(0.3)) == 2  # fallback: last index


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_reproduces_reference_related(name):
    """IBD-conditioned branch (FamilyBP + FamilyKnockoffs + sample_emission): with --n_threads 1 the
    reference seeds one taus88 with --seed (knockoffs.cpp:419-420) and walks the families in the order
    dumped by the probe.  Replaying that stream through oracle/kgw_oracle_family.c must reproduce the
    related haplotypes' knockoffs bit for bit and consume exactly as many draws as the reference made."""
    prob, f = load_golden(name)
    assert len(prob.fam_clusters) > 0
    rng, _ = ko.make_rng(prob, mode="taus88", taus_seed=prob.seed)
    seen = []
    for members, segs in prob.fam_clusters:
        r = ko.run_family(prob, members, segs, rng)
        assert np.array_equal(r.hk, f["Hk"][members]), f"family {members.tolist()}: knockoffs differ from reference"
        # Z / Zk agree on every IBD segment (the reference aborts otherwise, family_knockoffs.cpp:282-291)
        loc = {int(h): i for i, h in enumerate(members)}
        for (jmin, jmax, _, _, ids) in segs:
            rows = [loc[int(h)] for h in ids]
            assert (r.z[rows, jmin:jmax + 1] == r.z[rows[0], jmin:jmax + 1]).all()
            assert (r.zk[rows, jmin:jmax + 1] == r.zk[rows[0], jmin:jmax + 1]).all()
        seen += members.tolist()
    assert rng.n_draws == len(f["trace_related"])
    related = np.setdiff1d(np.arange(prob.N), prob.unrelated)
    assert sorted(seen) == related.tolist()
