"""GPU parity of the IBD-conditioned (related) branch -- FamilyBP + FamilyKnockoffs + sample_emission
(knockoffs.cpp:329-571, family_bp.cpp, family_knockoffs.cpp) -- called through the C ABI.

(a) keyed generator: CUDA == CPU oracle on every golden fixture's families;
(b) reference binary: the oracle replays the reference's own sequential taus88 stream (and is pinned to
    the reference bit for bit in tests/test_oracle_vs_reference.py) while recording each uniform at its
    (haplotype, stream, locus) key; fed those uniforms, the CUDA path must reproduce the reference's
    knockoffs of the related haplotypes bit for bit.
"""
import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu

kg = pytest.importorskip("knockoffgwas_b200")


def _families(prob):
    return [(m, [(s[0], s[1], s[4]) for s in segs]) for m, segs in prob.fam_clusters]


def _run(prob, **kw):
    return kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups,
                       win_start=prob.win_start, win_end=prob.win_end, unrelated=np.zeros(0, np.int32),
                       ref_global=prob.ref_global, seed=prob.seed, families=_families(prob), **kw)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_family_kernel_matches_oracle_counter_rng(name):
    prob, _ = load_golden(name)
    hk, dbg = _run(prob, debug_paths=True)
    rng, _ = ko.make_rng(prob, mode="philox")
    row = 0
    for members, segs in prob.fam_clusters:
        r = ko.run_family(prob, members, segs, rng)
        n = len(members)
        assert np.array_equal(dbg.fam_z[row:row + n], r.z), f"Z of family {members.tolist()}"
        assert np.array_equal(dbg.fam_zk[row:row + n], r.zk), f"Zk of family {members.tolist()}"
        assert np.array_equal(hk[members], r.hk), f"Hk of family {members.tolist()}"
        row += n
    related = np.concatenate([m for m, _ in prob.fam_clusters])
    rest = np.setdiff1d(np.arange(prob.N), related)
    assert not hk[rest].any()


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if "windows" not in n])
def test_family_kernel_reproduces_reference_binary(name):
    prob, f = load_golden(name)
    U = np.zeros((prob.N, 3, prob.p))
    rng, _ = ko.make_rng(prob, mode="taus88", taus_seed=prob.seed, record=U)
    for members, segs in prob.fam_clusters:
        r = ko.run_family(prob, members, segs, rng)
        assert np.array_equal(r.hk, f["Hk"][members])        # the oracle itself is pinned to the reference
    hk, _ = _run(prob, injected_u=U)
    related = np.concatenate([m for m, _ in prob.fam_clusters])
    assert np.array_equal(hk[related], f["Hk"][related]), "related knockoffs differ from the reference binary"


def test_whole_generate_call_matches_reference_binary():
    """Unrelated + related haplotypes in one kgw_generate() call == the reference's complete Hk."""
    prob, f = load_golden("toy_chr22_K10_res2")
    nu, p = len(prob.unrelated), prob.p
    U = np.zeros((prob.N, 3, p))
    rng, _ = ko.make_rng(prob, mode="taus88", taus_seed=prob.seed, record=U)
    for members, segs in prob.fam_clusters:
        ko.run_family(prob, members, segs, rng)
    raw = ko.taus88_stream(None, nu * 3 * p).astype(np.float64) / 4294967296.0   # unrelated: default-seeded engine
    U[prob.unrelated] = raw.reshape(nu, 3, p)
    hk, _ = kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, unrelated=prob.unrelated,
                        ref_global=prob.ref_global, seed=prob.seed, families=_families(prob), injected_u=U)
    assert np.array_equal(hk, f["Hk"])


def test_family_errors():
    prob, _ = load_golden("toy_chr21_K50_res5")
    fam = _families(prob)
    with pytest.raises(kg.KgwError):   # families need the global references
        kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, families=fam)
    bad = [(fam[0][0], [(fam[0][1][0][0], fam[0][1][0][1], np.array([0, 1], np.int32))])]
    with pytest.raises(kg.KgwError):   # a segment listing haplotypes outside its family
        kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, ref_global=prob.ref_global,
                    families=bad)
    with pytest.raises(kg.KgwError, match="also listed in `unrelated`"):   # the two branches would write the same rows
        kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, ref_global=prob.ref_global,
                    families=fam, unrelated=np.arange(prob.N, dtype=np.int32))
    from kgw_testutil import synthetic_problem
    wide = synthetic_problem(N=300, p=40, K=256, seed=3)   # K = 256 runs for unrelated haplotypes but not with families
    with pytest.raises(kg.KgwError, match="K > 255"):
        kg.generate(wide.H, wide.p, wide.K, wide.ref_local, wide.b, wide.mu, wide.groups, ref_global=wide.ref_global,
                    unrelated=np.arange(4, 300, dtype=np.int32),
                    families=[(np.array([0, 2], np.int32), [(5, 30, np.array([0, 2], np.int32))])])
    with pytest.raises(kg.KgwError, match="two families"):
        kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, ref_global=prob.ref_global,
                    families=fam + fam[:1], unrelated=np.zeros(0, np.int32))


def _synthetic_pairs(n_haps, p, K, pairs, mean_group=1.0, seed=3):
    from knockoffgwas_b200 import synth
    sp = synth.make_problem(n_haps, p, K, seed=seed, n_founders=60, mean_group=mean_group)
    synth.add_ibd_pairs(sp, pairs)
    prob = ko.Problem(H=sp.H, p=p, K=K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    return sp, prob


@pytest.mark.parametrize("K,mean_group", [(50, 1.0), (100, 4.0), (20, 1.0)])
def test_family_kernel_synthetic_pairs(K, mean_group):
    """Synthetic RaPID-style pairs (1-3 segments, SURVEY.md 8d C5 recipe), more families than resident warps'
    worth of work per CTA: Z, Zk and Hk of every family equal the oracle's.  Exercises the exact work skipping
    of the BP sweeps on chromosomes long enough for the perturbations to die out."""
    sp, prob = _synthetic_pairs(600, 3000, K, 24, mean_group)
    hk, dbg = kg.generate(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=np.zeros(0, np.int32),
                          ref_global=sp.ref, seed=sp.seed, families=sp.families, debug_paths=True)
    rng, _ = ko.make_rng(prob, mode="philox")
    row = 0
    for mem, segs in sp.families:
        r = ko.run_family(prob, mem, [(s[0], s[1], 0, 0, s[2]) for s in segs], rng)
        assert np.array_equal(dbg.fam_z[row:row + 2], r.z), f"Z of pair {mem.tolist()}"
        assert np.array_equal(dbg.fam_zk[row:row + 2], r.zk), f"Zk of pair {mem.tolist()}"
        assert np.array_equal(hk[mem], r.hk), f"Hk of pair {mem.tolist()}"
        row += 2


@pytest.mark.parametrize("size,K", [(3, 20), (4, 50), (6, 12)])
def test_family_kernel_synthetic_larger_families(size, K):
    """Families of 3-6 haplotypes whose segments involve changing subsets of the members (nodes with up to 5
    neighbours, several members joining / leaving at one junction): the generic junction paths, the special-range
    bookkeeping and the recomputed right messages against the oracle."""
    from knockoffgwas_b200 import synth
    sp = synth.make_problem(400, 1600, K, seed=21, n_founders=40, mean_dist_cm=0.02)
    synth.add_ibd_families(sp, 6, size=size, segments=5)
    assert len(sp.families) >= 4 and max(len(m) for m, _ in sp.families) >= 3
    prob = ko.Problem(H=sp.H, p=sp.p, K=K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    hk, dbg = kg.generate(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=np.zeros(0, np.int32),
                          ref_global=sp.ref, seed=sp.seed, families=sp.families, debug_paths=True)
    rng, _ = ko.make_rng(prob, mode="philox")
    row = 0
    for mem, segs in sp.families:
        r = ko.run_family(prob, mem, [(s[0], s[1], 0, 0, s[2]) for s in segs], rng)
        n = len(mem)
        assert np.array_equal(dbg.fam_z[row:row + n], r.z), f"Z of family {mem.tolist()}"
        assert np.array_equal(dbg.fam_zk[row:row + n], r.zk), f"Zk of family {mem.tolist()}"
        assert np.array_equal(hk[mem], r.hk), f"Hk of family {mem.tolist()}"
        row += n


@pytest.mark.parametrize("segs", [
    [(0, 40), (43, 45), (48, 48), (300, 599)],          # chromosome ends, a 3-locus gap, a one-locus segment
    [(0, 599)],                                         # one segment over the whole chromosome: no junction at all
    [(1, 1), (4, 5), (597, 598)],                       # junctions next to both ends
    [(0, 0), (2, 2), (599, 599)],
])
def test_family_kernel_segment_edge_cases(segs):
    """Hand-made pairs whose segments touch the chromosome ends, sit two loci apart (special ranges merge) or are one
    locus long: the stored spans of the right messages and the junction bookkeeping at the boundaries."""
    from knockoffgwas_b200 import synth
    sp = synth.make_problem(80, 600, 12, seed=17, n_founders=10, mean_dist_cm=0.05)
    ids = np.array([10, 33], dtype=np.int32)
    bits = np.unpackbits(sp.H, axis=1, bitorder="little")
    for j0, j1 in segs:
        bits[33, j0:j1 + 1] = bits[10, j0:j1 + 1]
    sp.H = np.packbits(bits, axis=1, bitorder="little")[:, :sp.H.shape[1]]
    for me, other in ((10, 33), (33, 10)):
        sp.ref[me][sp.ref[me] == other] = 79 if 79 not in sp.ref[me] else 78
        sp.ref[me].sort()
    fams = [(ids, [(j0, j1, ids) for j0, j1 in segs])]
    prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    hk, dbg = kg.generate(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=np.zeros(0, np.int32),
                          ref_global=sp.ref, seed=sp.seed, families=fams, debug_paths=True)
    rng, _ = ko.make_rng(prob, mode="philox")
    r = ko.run_family(prob, ids, [(j0, j1, 0, 0, ids) for j0, j1 in segs], rng)
    assert np.array_equal(dbg.fam_z, r.z) and np.array_equal(dbg.fam_zk, r.zk) and np.array_equal(hk[ids], r.hk)


@pytest.mark.parametrize("size,segments,p,K,max_subset,uniform_mu", [
    (12, 6, 1200, 20, 9, True),        # more than 8 members, nodes with up to 8 neighbours
    (32, 10, 1000, 12, 10, True),
    (100, 20, 800, 8, 12, True),       # a family as union-find over RaPID pairs builds them on biobank data
    (20, 8, 900, 40, 12, False),       # per-locus mutation rates (--hmm file): emissions formed with pow() on the device
])
def test_family_kernel_large_families(size, segments, p, K, max_subset, uniform_mu):
    """Families beyond the small class of the kernel (more than 8 haplotypes, nodes with more than 7 IBD neighbours): the
    reference has no limit (family_bp.cpp:41-61 pushes neighbours without bound, knockoffs.cpp:329-372 takes any family).
    One big family and a few pairs in the same call (both compiled classes launch), against the oracle."""
    from knockoffgwas_b200 import synth
    sp = synth.make_problem_fast(600, p, K, seed=71, n_founders=60, mean_dist_cm=0.05, mean_group=2.0)
    if not uniform_mu:
        sp.mu = np.random.default_rng(3).uniform(1e-6, 1e-3, size=p)
    synth.add_ibd_pairs(sp, 5)
    synth.add_big_family(sp, size, segments, max_subset=max_subset)
    mem, segs = sp.families[-1]
    assert len(mem) == size and (size > 8 or max(len(s[2]) for s in segs) > 8)
    prob = ko.Problem(H=sp.H, p=sp.p, K=K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    plan = kg.Plan(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=np.zeros(0, np.int32), ref_global=sp.ref,
                   seed=sp.seed, families=sp.families)
    assert plan.info()["launches_per_run"] == 5   # the big class's kernel + the four phase launches of the pairs' batch
    plan.run()
    hk, dbg = plan.download(debug_paths=True)
    plan.close()
    rng, _ = ko.make_rng(prob, mode="philox")
    row = 0
    for m, sl in sp.families:
        r = ko.run_family(prob, m, [(s[0], s[1], 0, 0, s[2]) for s in sl], rng)
        n = len(m)
        assert np.array_equal(dbg.fam_z[row:row + n], r.z), f"Z of the family of {n}"
        assert np.array_equal(dbg.fam_zk[row:row + n], r.zk), f"Zk of the family of {n}"
        assert np.array_equal(hk[m], r.hk), f"Hk of the family of {n}"
        row += n


def test_pair_class_in_several_batches(monkeypatch):
    """The pair class runs its phases over batches of as many pairs as have scratch slots (C5 at full size: ~1300 of a
    GPU's 4375 pairs at a time).  Forced here to 5 slots for 17 pairs -- four batches, the last one short -- at two
    partition resolutions on one plan: same Z, Zk, Hk as the oracle and as the one-batch run."""
    from knockoffgwas_b200 import synth
    sp = synth.make_problem_fast(400, 700, 20, seed=23, n_founders=40, mean_dist_cm=0.05, mean_group=1.0)
    synth.add_ibd_pairs(sp, 17)
    prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    coarse = (np.arange(sp.p) // 3).astype(np.int32)
    results = {}
    for slots in ("0", "5", "generic"):
        # "generic": the operation lists run in order inside family_kernel (the path taken when a pair's list does not
        # allow "all DRAWs, then the copies")
        monkeypatch.setenv("KGW_PAIR_SLOTS", "5" if slots == "generic" else slots)
        monkeypatch.setenv("KGW_PAIR_KO_GENERIC", "1" if slots == "generic" else "0")
        plan = kg.Plan(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=np.zeros(0, np.int32),
                       ref_global=sp.ref, seed=sp.seed, families=sp.families)
        out = []
        for groups in (sp.groups, coarse):
            plan.set_groups(groups)
            plan.run()
            hk, dbg = plan.download(debug_paths=True)
            out.append((hk.copy(), dbg.fam_z.copy(), dbg.fam_zk.copy()))
        plan.close()
        results[slots] = out
    for other in ("5", "generic"):
        for a, b in zip(results["0"], results[other]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
    rng, _ = ko.make_rng(prob, mode="philox")
    row = 0
    hk, fz, fzk = results["5"][0]
    for m, sl in sp.families:
        r = ko.run_family(prob, m, [(s[0], s[1], 0, 0, s[2]) for s in sl], rng)
        n = len(m)
        assert np.array_equal(fz[row:row + n], r.z) and np.array_equal(fzk[row:row + n], r.zk) and np.array_equal(hk[m], r.hk)
        row += n


@pytest.mark.parametrize("K,wide", [(50, "0"), (50, "1"), (100, "0"), (100, "1"), (24, "0"), (120, "0"), (200, "0")])
def test_pair_kernels_both_forms(monkeypatch, K, wide):
    """The pair draw kernel has two forms -- eight lanes per chain and two pairs per warp (long lists of pairs), sixteen
    lanes and one pair per warp (short lists) -- and lane widths for K up to 255; the launcher picks by list length.
    Both forced here on the same pairs, against the oracle."""
    from knockoffgwas_b200 import synth
    monkeypatch.setenv("KGW_PAIR_WIDE", wide)
    sp = synth.make_problem_fast(300, 500, K, seed=K, n_founders=40, mean_dist_cm=0.05, mean_group=1.5)
    synth.add_ibd_pairs(sp, 7)
    prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    hk, dbg = kg.generate(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=np.zeros(0, np.int32),
                          ref_global=sp.ref, seed=sp.seed, families=sp.families, debug_paths=True)
    rng, _ = ko.make_rng(prob, mode="philox")
    row = 0
    for m, sl in sp.families:
        r = ko.run_family(prob, m, [(s[0], s[1], 0, 0, s[2]) for s in sl], rng)
        n = len(m)
        assert np.array_equal(dbg.fam_z[row:row + n], r.z) and np.array_equal(dbg.fam_zk[row:row + n], r.zk)
        assert np.array_equal(hk[m], r.hk)
        row += n


def test_mixed_plan_reuse_with_shared_scratch(monkeypatch):
    """Unrelated haplotypes and pairs on ONE plan over three runs (two partitions, the first one again): with IBD families
    in the call the sampler's scratch lives inside the families' message buffer (the two branches never run at the same
    time), so every run must start from a clean slate.  Same rows as with separate buffers and as the oracle."""
    from knockoffgwas_b200 import synth
    sp = synth.make_problem_fast(400, 700, 20, seed=31, n_founders=40, mean_dist_cm=0.05, mean_group=1.0)
    synth.add_ibd_pairs(sp, 40)
    related = np.concatenate([m for m, _ in sp.families])
    unrelated = np.setdiff1d(np.arange(sp.N, dtype=np.int32), related)[:70].astype(np.int32)
    prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    coarse = (np.arange(sp.p) // 4).astype(np.int32)
    runs = {}
    for alias in ("1", "0"):
        monkeypatch.setenv("KGW_ALIAS_SCRATCH", alias)
        plan = kg.Plan(sp.H, sp.p, sp.K, prob.ref_local, sp.b, sp.mu, sp.groups, unrelated=unrelated, ref_global=sp.ref,
                       seed=sp.seed, families=sp.families)
        out = []
        for groups in (sp.groups, coarse, sp.groups):
            plan.set_groups(groups)
            plan.run()
            hk, _ = plan.download()
            out.append(hk.copy())
        plan.close()
        runs[alias] = out
    for a, b in zip(runs["1"], runs["0"]):
        assert np.array_equal(a, b)
    assert np.array_equal(runs["1"][0], runs["1"][2])
    assert not np.array_equal(runs["1"][0], runs["1"][1])
    r = ko.run_unrelated(prob, unrelated[:8])
    assert np.array_equal(runs["1"][0][unrelated[:8]], r.hk)
    rng, _ = ko.make_rng(prob, mode="philox")
    for m, sl in sp.families[:6]:
        rf = ko.run_family(prob, m, [(s[0], s[1], 0, 0, s[2]) for s in sl], rng)
        assert np.array_equal(runs["1"][0][m], rf.hk)
