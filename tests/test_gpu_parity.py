"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libkgw.so),
against (a) the golden vectors written by the unmodified reference binary and (b) the CPU
oracle on the same seeded inputs.

Bar (BASELINE.json north_star): latent paths and knockoff alleles bit-exact under the same
uniforms; backward probabilities within 1e-5 relative (we assert 1e-9, the kernel is fp64).
"""
import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import GOLDEN_NAMES, load_golden, synthetic_problem

pytestmark = pytest.mark.gpu

kg = pytest.importorskip("knockoffgwas_b200")


def _taus_injection(prob):
    """U[N][3][p]: the uniforms the reference consumed with --n_threads 1 (default-seeded taus88,
    haplotype-major, Z block / Zk block / Hk block; knockoffs.cpp:622-624, :1243-1500)."""
    nu, p = len(prob.unrelated), prob.p
    raw = ko.taus88_stream(None, nu * 3 * p).astype(np.float64) / 4294967296.0
    U = np.zeros((prob.N, 3, p))
    U[prob.unrelated] = raw.reshape(nu, 3, p)
    return U


def _run(prob, **kw):
    return kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups,
                       win_start=prob.win_start, win_end=prob.win_end, unrelated=prob.unrelated,
                       seed=prob.seed, **kw)


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if "windows" not in n])
def test_kernel_reproduces_reference_binary(name):
    """Kernel fed the reference's own taus88 uniforms == haplotype-level output of the reference."""
    prob, f = load_golden(name)
    hk, dbg = _run(prob, injected_u=_taus_injection(prob), debug_paths=True)
    assert np.array_equal(dbg.z, f["z_unrelated"].astype(np.int32)), "Z differs from the reference"
    assert np.array_equal(dbg.zk, f["zk_unrelated"].astype(np.int32)), "Zk differs from the reference"
    assert np.array_equal(hk[prob.unrelated], f["Hk"][prob.unrelated]), "Hk differs from the reference"
    # rows that were not requested stay untouched (zero), as generate_unrelated leaves them
    rest = np.setdiff1d(np.arange(prob.N), prob.unrelated)
    assert not hk[rest].any()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_kernel_matches_oracle_counter_rng(name):
    """Production generator (Philox keyed by seed/hap/stream/locus/window): device == host oracle."""
    prob, _ = load_golden(name)
    haps = prob.unrelated[:: max(1, len(prob.unrelated) // 300)]
    sub = ko.Problem(H=prob.H, p=prob.p, K=prob.K, ref_local=prob.ref_local, ref_global=prob.ref_global,
                     b=prob.b, mu=prob.mu, groups=prob.groups, win_start=prob.win_start, win_end=prob.win_end,
                     unrelated=haps, seed=prob.seed)
    res = ko.run_unrelated(sub, mode="philox")
    hk, dbg = _run(sub, debug_paths=True)
    assert np.array_equal(dbg.z, res.z)
    assert np.array_equal(dbg.zk, res.zk)
    assert np.array_equal(hk[haps], res.hk)


CASES = [
    # N, p, K, mean_group, tiny_dist_frac, W, options
    dict(N=40, p=97, K=3, mean_group=1.0),
    dict(N=40, p=131, K=10, mean_group=3.0),
    dict(N=70, p=257, K=16, mean_group=1.0, tiny_dist_frac=0.3),
    dict(N=70, p=300, K=29, mean_group=8.0),
    dict(N=90, p=200, K=50, mean_group=1.0),
    dict(N=90, p=200, K=50, mean_group=1.0, lanes=4),
    dict(N=90, p=200, K=50, mean_group=1.0, lanes=8, block=2),
    dict(N=90, p=230, K=50, mean_group=4.0, lanes=2, block=8),
    dict(N=70, p=200, K=25, mean_group=1.0, lanes=1),
    dict(N=90, p=260, K=50, mean_group=1.0, lanes=1),             # one lane per haplotype, 50 states, eight chunks
    dict(N=90, p=230, K=50, mean_group=4.0, lanes=1, block=8),
    dict(N=70, p=200, K=41, mean_group=2.0, lanes=1),             # padded: 41 of 50 states
    dict(N=150, p=170, K=100, mean_group=1.0, lanes=2),           # 50 states per lane on two lanes
    dict(N=150, p=170, K=87, mean_group=3.0, lanes=2),
    dict(N=220, p=130, K=200, mean_group=1.0, lanes=4),           # ... and on four
    dict(N=220, p=130, K=150, mean_group=2.0, lanes=4),
    dict(N=70, p=200, K=32, mean_group=3.0, lanes=1, block=8),
    dict(N=33, p=70, K=10, mean_group=1.0, lanes=1),
    dict(N=33, p=70, K=10, mean_group=2.0, lanes=8),
    dict(N=90, p=260, K=56, mean_group=2.0, tiny_dist_frac=0.5),
    dict(N=120, p=190, K=64, mean_group=5.0),
    dict(N=150, p=150, K=100, mean_group=1.0),
    dict(N=150, p=150, K=128, mean_group=30.0),
    dict(N=300, p=96, K=200, mean_group=1.0),
    dict(N=60, p=400, K=12, mean_group=4.0, W=3),
    dict(N=60, p=333, K=50, mean_group=2.0, W=4, block=2),
    dict(N=50, p=300, K=20, mean_group=1.0, block=8),
    dict(N=50, p=300, K=20, mean_group=6.0, block=2),
    dict(N=90, p=300, K=50, mean_group=1.0, block=4, seg=64),
    dict(N=90, p=301, K=50, mean_group=3.0, lanes=4, block=8, seg=128),
    dict(N=60, p=333, K=50, mean_group=2.0, W=4, block=2, seg=32),
    dict(N=50, p=257, K=20, mean_group=1.0, seg=96),
    dict(N=40, p=200, K=7, mean_group=1.0, lanes=2, block=4),
    dict(N=30, p=1, K=5, mean_group=1.0),
    dict(N=30, p=2, K=5, mean_group=1.0),
    dict(N=30, p=33, K=5, mean_group=50.0),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_kernel_matches_oracle_synthetic(case):
    c = dict(case)
    lanes, block, seg = c.pop("lanes", 0), c.pop("block", 0), c.pop("seg", 0)
    prob = synthetic_problem(seed=len(repr(case)), **c)
    res = ko.run_unrelated(prob, mode="philox", want_beta=True)
    hk, dbg = _run(prob, options=kg.Options(lanes_per_hap=lanes, block_loci=block, segment_loci=seg),
                   debug_paths=True, debug_beta=True)
    assert np.array_equal(dbg.z, res.z)
    assert np.array_equal(dbg.zk, res.zk)
    assert np.array_equal(hk, res.hk)
    # backward probabilities at the checkpoint loci (last window pass)
    loci = dbg.ckpt_locus
    ok = loci >= 0
    assert ok.any()
    want = res.beta[:, loci[ok], :]
    got = dbg.beta_ckpt[:, ok, :]
    assert np.allclose(got, want, rtol=1e-9, atol=1e-300), float(np.abs(got / want - 1).max())


def test_injected_uniform_edge_values():
    """u = 0, u just below 1 and zero-probability states follow weighted_choice exactly."""
    prob = synthetic_problem(N=48, p=64, K=7, mean_group=2.0, seed=5)
    rng = np.random.default_rng(1)
    U = rng.random((prob.N, 3, prob.p))
    U[::3] = 0.0
    U[1::3] = 1.0 - 2.0 ** -32
    U[2::3, :, ::2] = 0.0
    res = ko.run_unrelated(prob, mode="injected", injected=U)
    hk, dbg = _run(prob, injected_u=U, debug_paths=True)
    assert np.array_equal(dbg.z, res.z) and np.array_equal(dbg.zk, res.zk) and np.array_equal(hk, res.hk)


def test_shard_independence_and_determinism():
    """Rows do not depend on which other haplotypes share the launch (multi-GPU sharding),
    nor on the kernel geometry (lanes per haplotype, checkpoint spacing)."""
    prob = synthetic_problem(N=200, p=160, K=50, mean_group=1.0, seed=11)
    full, _ = _run(prob)
    again, _ = _run(prob)
    assert np.array_equal(full, again)
    part = np.arange(37, 200, 3, dtype=np.int32)
    sub, _ = kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups,
                         unrelated=part, seed=prob.seed)
    assert np.array_equal(sub[part], full[part])
    alt, _ = _run(prob, options=kg.Options(lanes_per_hap=4, block_loci=8))
    assert np.array_equal(alt, full)


def test_plan_reuse_across_partitions():
    """kgw_plan_set_groups == Knockoffs::set_partition without re-uploading H."""
    prob = synthetic_problem(N=64, p=200, K=10, mean_group=1.0, seed=3)
    coarse = synthetic_problem(N=64, p=200, K=10, mean_group=7.0, seed=3).groups
    plan = kg.Plan(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, seed=prob.seed)
    plan.run()
    hk0, _ = plan.download()
    plan.set_groups(coarse)
    ms = plan.run_timed(2)
    hk1, _ = plan.download()
    plan.close()
    assert ms > 0
    assert np.array_equal(hk0, ko.run_unrelated(prob, mode="philox").hk)
    prob.groups = coarse
    assert np.array_equal(hk1, ko.run_unrelated(prob, mode="philox").hk)


def test_host_replay_of_generator():
    for (hap, s, j) in [(0, 0, 0), (5, 1, 3), (1999, 2, 999), (699999, 1, 59999)]:
        assert kg.uniform(2020, hap, s, j) == ko.philox_uniform(2020, hap, s, j)


def test_error_behaviour():
    prob = synthetic_problem(N=16, p=40, K=4, seed=2)
    bad = prob.groups.copy()
    bad[10] = 99
    with pytest.raises(kg.KgwError):
        kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, bad)
    with pytest.raises(kg.KgwError):
        kg.generate(prob.H, prob.p, 300, np.zeros((16, 1, 300), np.int32), prob.b, prob.mu, prob.groups)


def test_kernel_reproduces_reference_binary_with_windows():
    """--windows > 0: the padded passes (knockoffs.cpp:1222-1226) overlap, so the reference consumes several uniforms
    for one (haplotype, stream, locus).  The oracle replays the reference's taus88 stream on the windows fixture (it
    reproduces the binary there, tests/test_oracle_vs_reference.py) and records every uniform under (haplotype, window
    pass, stream, locus); the kernel fed that record must return the reference binary's Z, Zk and Hk."""
    prob, f = load_golden("toy_chr22_K10_res1_windows")
    assert prob.W > 1
    U = np.zeros((prob.N, prob.W, 3, prob.p))
    res = ko.run_unrelated(prob, mode="taus88", record=U)
    assert np.array_equal(res.hk, f["Hk"][prob.unrelated]), "the oracle replay is not the reference's output"
    hk, dbg = _run(prob, injected_u=U, debug_paths=True)
    assert np.array_equal(dbg.z, f["z_unrelated"].astype(np.int32)), "Z differs from the reference"
    assert np.array_equal(dbg.zk, f["zk_unrelated"].astype(np.int32)), "Zk differs from the reference"
    assert np.array_equal(hk[prob.unrelated], f["Hk"][prob.unrelated]), "Hk differs from the reference"


@pytest.mark.parametrize("wide", ["0", "1"])
def test_one_lane_forms_match_oracle(wide, monkeypatch):
    """K = 50 with one lane per haplotype in both compiled forms: the slim shared-memory layout (one message-row buffer,
    knockoff-table rows in the ring, one pair of latent-path tiles; KGW_FORCE_WIDE=0) and the double-buffered one."""
    monkeypatch.setenv("KGW_FORCE_WIDE", wide)
    prob = synthetic_problem(N=120, p=333, K=50, seed=17, mean_group=2.5)
    res = ko.run_unrelated(prob, mode="philox")
    hk, dbg = _run(prob, options=kg.Options(lanes_per_hap=1, segment_loci=128), debug_paths=True)
    assert np.array_equal(dbg.z, res.z) and np.array_equal(dbg.zk, res.zk) and np.array_equal(hk, res.hk)
