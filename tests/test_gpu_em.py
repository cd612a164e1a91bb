"""--estimate-hmm on the device: kgw_estimate_hmm against the CPU oracle (oracle/kgw_oracle.c:
kgw_oracle_em, itself pinned to the unmodified reference binary on the toy example, tests/test_oracle_em.py) and against
the reference binary's own rates (tests/golden/em_toy_chr22_K10.npz).  gamma_j is a floating-point sum over haplotypes
accumulated with atomic adds, so the rates are compared to 1e-9 relative (the north star asks 1e-5 for floating-point
results); the number of EM steps must be the same."""
import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import GOLDEN, load_golden, synthetic_problem

pytestmark = pytest.mark.gpu
kg = pytest.importorskip("knockoffgwas_b200")


def test_em_matches_reference_binary_on_the_toy_example():
    prob, _ = load_golden("toy_chr22_K10_res2")
    g = np.load(GOLDEN / "em_toy_chr22_K10.npz")
    mu, steps, deltas, converged = kg.estimate_hmm(prob.H, prob.p, prob.K, prob.ref_local, prob.b)
    assert converged and steps == int(g["converged_after"])
    assert np.allclose(mu, g["mu"], rtol=1e-9, atol=0.0)
    assert np.allclose(deltas, g["deltas"], rtol=2e-3)
    assert mu.min() >= 1e-6 and mu.max() <= 1e-3


@pytest.mark.parametrize("N,p,K,W", [(60, 97, 6, 1), (80, 300, 33, 1), (50, 260, 50, 3), (40, 130, 100, 1), (64, 200, 200, 2)])
def test_em_matches_oracle(N, p, K, W):
    prob = synthetic_problem(N=max(N, K + 2), p=p, K=K, seed=K + W, W=W)
    want_mu, want_steps, want_deltas, want_conv = ko.run_em(prob)
    mu, steps, deltas, conv = kg.estimate_hmm(prob.H, prob.p, prob.K, prob.ref_local, prob.b, win_start=prob.win_start,
                                              win_end=prob.win_end)
    assert (steps, conv) == (want_steps, want_conv)
    assert np.allclose(mu, want_mu, rtol=1e-9, atol=0.0)
    assert np.allclose(deltas, want_deltas, rtol=1e-9)


def test_em_mirror_installs_the_rates():
    prob = synthetic_problem(N=40, p=80, K=8, seed=5)
    k = kg.Knockoffs(prob.H, prob.p, prob.K, genetic_dist=np.full(prob.p, 0.01))
    k.set_references(prob.ref_local)
    k.init_hmm(prob.K, 1.0, 1e-5)
    steps, deltas, conv = k.EM()
    assert 1 <= steps <= 20 and len(deltas) == steps
    assert k.mut_rates.min() >= 1e-6 and k.mut_rates.max() <= 1e-3
