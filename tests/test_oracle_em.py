"""--estimate-hmm: the CPU restatement of Knockoffs::EM (oracle/kgw_oracle.c: kgw_oracle_em) against
the unmodified reference binary run with --estimate-hmm on its toy chr22 example (tests/golden/em_toy_chr22_K10.npz,
made by tests/golden/make_golden_em.py).  The reference is built with -Ofast, so the mutation rates are required to
agree to 1e-12 relative (measured: 4e-16, 984 of 1000 bit-identical), the number of EM steps exactly.
The CUDA side of this row is not started; this pins the oracle it will be checked against."""
import numpy as np

import kgw_oracle as ko
from kgw_testutil import GOLDEN, load_golden


def test_em_restatement_matches_reference():
    prob, _ = load_golden("toy_chr22_K10_res2")
    g = np.load(GOLDEN / "em_toy_chr22_K10.npz")
    assert np.array_equal(prob.b, g["b"])
    mu, steps, deltas, converged = ko.run_em(prob)        # from init_hmm(K, rho): every mu_j = 1e-5
    assert converged and steps == int(g["converged_after"]) == len(g["iterations"])
    assert np.allclose(deltas, g["deltas"], rtol=2e-3)     # the reference prints three decimals
    assert np.allclose(mu, g["mu"], rtol=1e-12, atol=0.0)
    assert mu.min() >= 1e-6 and mu.max() <= 1e-3           # the clamp of EM_step (knockoffs.cpp:917-920)


def test_em_is_a_fixed_point_iteration_on_small_data():
    from kgw_testutil import synthetic_problem
    prob = synthetic_problem(N=40, p=60, K=6, seed=3)
    mu1, n1, d1, _ = ko.run_em(prob)
    assert 1 <= n1 <= 20 and len(d1) == n1 and np.all(np.diff(d1) < 0)          # the change shrinks from step to step
    mu2, n2, d2, conv2 = ko.run_em(prob, mu_init=mu1)                            # restarted at the result: converged at once
    assert conv2 and n2 == 1 and d2[0] < 5e-2
