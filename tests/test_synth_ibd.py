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
