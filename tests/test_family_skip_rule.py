"""The related-branch kernel (knockoffgwas_b200/csrc/kgw_family.cuh) recomputes a loopy-BP message only in the first
iteration, around interval boundaries, or when its own-chain input changed bits in the same sweep.  The claim that
this is EXACT (every skipped message would have been recomputed to the bits already stored, adding 0 to the
convergence sum) is audited here on the CPU: the oracle computes every message as the reference does and counts
the ones the rule would skip and the ones among them that would have changed ("violations").  Also pins that the
audit does not disturb the oracle's own results."""
import numpy as np

import kgw_oracle as ko
from kgw_testutil import GOLDEN_NAMES, load_golden
from knockoffgwas_b200 import synth


def _audit(prob, families):
    rng, _ = ko.make_rng(prob, mode="philox")
    plain = [ko.run_family(prob, m, s, rng) for m, s in families]
    ko.family_audit(True)
    try:
        rng, _ = ko.make_rng(prob, mode="philox")
        audited = [ko.run_family(prob, m, s, rng) for m, s in families]
        st = ko.family_audit_stats()
    finally:
        ko.family_audit(False)
    for a, b in zip(plain, audited):
        assert np.array_equal(a.z, b.z) and np.array_equal(a.zk, b.zk) and np.array_equal(a.hk, b.hk)
        assert a.bp_sweeps == b.bp_sweeps
    return st


def test_rule_is_exact_on_reference_fixtures():
    tot = {"messages": 0, "skippable": 0, "violations": 0}
    for name in ("toy_chr22_K10_res2", "toy_chr21_K50_res5"):
        prob, _ = load_golden(name)
        st = _audit(prob, prob.fam_clusters)
        for k in tot:
            tot[k] += st[k]
    assert tot["violations"] == 0
    assert tot["skippable"] > 0.5 * tot["messages"]          # families of 2 and 3 on 1000-SNP chromosomes: ~70 %


def test_rule_is_exact_on_synthetic_pairs_and_larger_families():
    sp = synth.make_problem(300, 4000, 20, seed=8, n_founders=40)
    synth.add_ibd_pairs(sp, 6)
    synth.add_ibd_families(sp, 4, size=4, segments=5)
    prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups, win_start=None, win_end=None, seed=sp.seed)
    fams = [(m, [(s[0], s[1], 0, 0, s[2]) for s in segs]) for m, segs in sp.families]
    st = _audit(prob, fams)
    assert st["violations"] == 0
    assert st["skippable"] > 0.7 * st["messages"]            # long chromosomes: the perturbations die out
