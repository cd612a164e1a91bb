"""Audit mode of the unrelated sampler (kgw_generate_strict, csrc/kgw_strict.cu; SURVEY.md section 7, hard part 1).

The reference sums every quantity over the K states left to right in IEEE double; the product kernels sum over lanes as
a tree with fused multiply-adds.  kgw_generate_strict runs the reference's evaluation order on the device.  Bar:
  * strict mode == the strict-IEEE CPU restatement BIT FOR BIT -- latent paths, knockoff paths, knockoff alleles and
    every backward message (np.array_equal on float64, no tolerance), with the keyed generator and with the
    reference's own taus88 uniforms injected, on the golden fixtures (windows included) and on synthetic problems;
  * strict mode == what the unmodified reference binary wrote (tests/golden);
  * the fast path == strict mode on all of these ("a tree-order fast mode whose equality on the test sets is asserted").
"""
import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import GOLDEN_NAMES, load_golden, synthetic_problem

pytestmark = pytest.mark.gpu

kg = pytest.importorskip("knockoffgwas_b200")


def _args(prob):
    return (prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups)


def _kw(prob, **kw):
    return dict(win_start=prob.win_start, win_end=prob.win_end, unrelated=prob.unrelated, seed=prob.seed, **kw)


def _sub(prob, haps):
    return ko.Problem(H=prob.H, p=prob.p, K=prob.K, ref_local=prob.ref_local, ref_global=prob.ref_global, b=prob.b,
                      mu=prob.mu, groups=prob.groups, win_start=prob.win_start, win_end=prob.win_end,
                      unrelated=np.ascontiguousarray(haps, dtype=np.int32), seed=prob.seed)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_strict_mode_equals_oracle_bit_for_bit(name):
    prob, _ = load_golden(name)
    sub = _sub(prob, prob.unrelated[:: max(1, len(prob.unrelated) // 40)])
    res = ko.run_unrelated(sub, mode="philox", want_beta=True)
    hk, dbg = kg.generate_strict(*_args(sub), **_kw(sub, debug_paths=True, debug_beta=True))
    assert np.array_equal(dbg.z, res.z)
    assert np.array_equal(dbg.zk, res.zk)
    assert np.array_equal(hk[sub.unrelated], res.hk)
    assert np.array_equal(dbg.ckpt_locus, np.arange(sub.p))
    assert np.array_equal(dbg.beta_ckpt, res.beta), "a backward message differs in its last bit or more"
    # the fast path draws the same paths and alleles; its messages (tree sums, FMA) agree to rounding
    hk_f, dbg_f = kg.generate(*_args(sub), **_kw(sub, debug_paths=True, debug_beta=True))
    assert np.array_equal(dbg_f.z, dbg.z) and np.array_equal(dbg_f.zk, dbg.zk) and np.array_equal(hk_f, hk)
    if sub.W == 1:
        rows = dbg.beta_ckpt[:, dbg_f.ckpt_locus, :]
        assert np.allclose(dbg_f.beta_ckpt, rows, rtol=1e-9, atol=0)


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if "windows" not in n])
def test_strict_mode_reproduces_reference_binary(name):
    """Fed the uniforms the reference consumed (taus88, --n_threads 1), the audit mode returns what the reference wrote."""
    prob, f = load_golden(name)
    nu, p = len(prob.unrelated), prob.p
    raw = ko.taus88_stream(None, nu * 3 * p).astype(np.float64) / 4294967296.0
    U = np.zeros((prob.N, 3, p))
    U[prob.unrelated] = raw.reshape(nu, 3, p)
    hk, dbg = kg.generate_strict(*_args(prob), **_kw(prob, injected_u=U, debug_paths=True))
    assert np.array_equal(dbg.z, f["z_unrelated"].astype(np.int32))
    assert np.array_equal(dbg.zk, f["zk_unrelated"].astype(np.int32))
    assert np.array_equal(hk[prob.unrelated], f["Hk"][prob.unrelated])
    rest = np.setdiff1d(np.arange(prob.N), prob.unrelated)
    assert not hk[rest].any()


@pytest.mark.parametrize("K,mean_group,W,N,p", [(8, 1.0, 1, 64, 200), (50, 3.0, 1, 120, 400), (13, 6.0, 3, 70, 600),
                                                 (100, 2.0, 2, 150, 300), (2, 1.0, 1, 8, 40), (5, 50.0, 1, 40, 333)])
def test_strict_mode_on_synthetic_problems(K, mean_group, W, N, p):
    prob = synthetic_problem(N=N, p=p, K=K, seed=K + W, mean_group=mean_group, W=W, tiny_dist_frac=0.05)
    res = ko.run_unrelated(prob, mode="philox", want_beta=True)
    hk, dbg = kg.generate_strict(*_args(prob), **_kw(prob, debug_paths=True, debug_beta=True))
    assert np.array_equal(dbg.z, res.z) and np.array_equal(dbg.zk, res.zk) and np.array_equal(hk, res.hk)
    assert np.array_equal(dbg.beta_ckpt, res.beta)
    hk_f, dbg_f = kg.generate(*_args(prob), **_kw(prob, debug_paths=True))
    assert np.array_equal(dbg_f.z, dbg.z) and np.array_equal(dbg_f.zk, dbg.zk) and np.array_equal(hk_f, hk)


def test_strict_mode_argument_errors():
    prob = synthetic_problem(N=16, p=40, K=4)
    with pytest.raises(kg.KgwError):
        kg.generate_strict(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups[::-1].copy())
    bad = prob.ref_local.copy()
    bad[3, 0, 1] = prob.N
    with pytest.raises(kg.KgwError):
        kg.generate_strict(prob.H, prob.p, prob.K, bad, prob.b, prob.mu, prob.groups)
    hk, _ = kg.generate_strict(*_args(prob), unrelated=np.zeros(0, np.int32))
    assert not hk.any()


def test_strict_mode_has_no_bound_on_K():
    """The fast path keeps latent states in bytes (K <= 256, KGW_ERR_UNSUPPORTED beyond); the audit mode keeps them in
    int32 like the reference and takes any K."""
    prob = synthetic_problem(N=330, p=48, K=300, seed=5, mean_group=2.0)
    haps = np.arange(0, 330, 11, dtype=np.int32)
    sub = _sub(prob, haps)
    res = ko.run_unrelated(sub, mode="philox", want_beta=True)
    hk, dbg = kg.generate_strict(*_args(sub), **_kw(sub, debug_paths=True, debug_beta=True))
    assert np.array_equal(dbg.z, res.z) and np.array_equal(dbg.zk, res.zk) and np.array_equal(hk[haps], res.hk)
    assert np.array_equal(dbg.beta_ckpt, res.beta)
    assert dbg.z.max() > 255
    with pytest.raises(kg.KgwError):
        kg.generate(*_args(sub), **_kw(sub))


def test_knockoffs_mirror_audits_its_last_generate():
    prob = synthetic_problem(N=96, p=300, K=12, seed=3, mean_group=3.0)
    k = kg.Knockoffs(prob.H, prob.p, prob.K, seed=prob.seed)
    k.set_references(prob.ref_local)
    k.load_hmm(-np.log(np.maximum(prob.b, 1e-300)), prob.mu)
    k.set_groups(prob.groups)
    k.generate()
    assert len(k.audit(40)) == 0
    k.Hk[55, 7] ^= 0x10                       # a flipped knockoff allele is found when its row is among the sampled ones
    assert k.audit(96).tolist() == [55]
