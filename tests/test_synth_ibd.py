"""CPU checks of the synthetic IBD generator (knockoffgwas_b200.synth.add_ibd_pairs, SURVEY.md 8d C5 recipe)
and of the oracle's related branch on it."""
import numpy as np

import kgw_oracle as ko
from knockoffgwas_b200 import synth


def test_pairs_are_identical_on_segments_and_not_references():
    sp = synth.make_problem(400, 1500, 10, seed=11, n_founders=30)
    synth.add_ibd_pairs(sp, 20)
    assert len(sp.families) == 20
    bits = np.unpackbits(sp.H, axis=1, bitorder="little")[:, :sp.p]
    seen = set()
    for mem, segs in sp.families:
        a, c = (int(x) for x in mem)
        assert a < c and a // 2 != c // 2                       # different samples
        assert not ({a, c} & seen)                              # disjoint pairs
        seen |= {a, c}
        assert 1 <= len(segs) <= 3
        last = -10
        for j0, j1, ids in segs:
            assert 0 < j0 <= j1 < sp.p - 1 and j0 > last + 2    # sorted, disjoint, not touching
            last = j1
            assert sp.bp[j1] - sp.bp[j0] >= 100_000
            assert np.array_equal(ids, mem)
            assert np.array_equal(bits[a, j0:j1 + 1], bits[c, j0:j1 + 1])
        assert c not in sp.ref[a] and a not in sp.ref[c]
        assert np.all(np.diff(sp.ref[a]) > 0) and np.all(np.diff(sp.ref[c]) > 0)


def test_oracle_runs_synthetic_pair_and_shares_paths_on_segments():
    sp = synth.make_problem(200, 600, 8, seed=4, n_founders=20)
    synth.add_ibd_pairs(sp, 3)
    prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    rng, _ = ko.make_rng(prob, mode="philox")
    for mem, segs in sp.families:
        r = ko.run_family(prob, mem, [(s[0], s[1], 0, 0, s[2]) for s in segs], rng)
        assert r.bp_sweeps >= 4
        for j0, j1, _ in segs:                                   # one latent path / knockoff path inside a segment
            assert np.array_equal(r.z[0, j0:j1 + 1], r.z[1, j0:j1 + 1])
            assert np.array_equal(r.zk[0, j0:j1 + 1], r.zk[1, j0:j1 + 1])


def test_larger_families_are_tidy_and_identical_on_segments():
    sp = synth.make_problem(300, 1200, 10, seed=2, n_founders=30, mean_dist_cm=0.02)
    synth.add_ibd_families(sp, 5, size=4, segments=5)
    bits = np.unpackbits(sp.H, axis=1, bitorder="little")[:, :sp.p]
    assert sp.families
    for mem, segs in sp.families:
        assert len({int(m) // 2 for m in mem}) == len(mem)               # one haplotype per sample
        spans = sorted((j0, j1) for j0, j1, _ in segs)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert b0 > a1 + 1                                              # neither overlapping nor touching
        keys = [(len(ids), tuple(int(x) for x in ids), j0) for j0, _, ids in segs]
        assert keys == sorted(keys)                                         # std::set<IbdSeg> order
        for j0, j1, ids in segs:
            assert set(int(x) for x in ids) <= set(int(x) for x in mem) and len(ids) >= 2
            assert sp.bp[j1] - sp.bp[j0] >= 100_000
            for other in ids[1:]:
                assert np.array_equal(bits[ids[0], j0:j1 + 1], bits[other, j0:j1 + 1])
        for me in mem:
            assert not (set(int(x) for x in sp.ref[me]) & set(int(x) for x in mem))
