"""Parity at the shapes BASELINE.json's configs are quoted on ("what's weak" 1):
  * C3: p = 60 000 SNPs, K = 50, with the chromosome cut into segments automatically (sweep 0, thousands of
    power-of-two rescales, segment-sized scratch) -- rows at warp / slot edges against the oracle;
  * C5: K = 100, p = 60 000, related pairs and families of 3-6 with the unrelated haplotypes in the same call;
  * C4: six partition levels through kgw_plan_set_groups on one resident plan.
Reference lines: knockoffs.cpp:636-700 (unrelated worker), family_bp.cpp:79-230, family_knockoffs.cpp:77-293."""
import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import oracle_families, oracle_unrelated_rows

pytestmark = pytest.mark.gpu
kg = pytest.importorskip("knockoffgwas_b200")


def _problem(sp, groups=None):
    return ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=sp.ref[:, None, :], ref_global=sp.ref, b=sp.b, mu=sp.mu,
                      groups=sp.groups if groups is None else groups, win_start=None, win_end=None, seed=sp.seed)


def test_c3_shape_auto_segmented_rows_match_oracle():
    from knockoffgwas_b200 import synth
    N, p, K = 12000, 60000, 50
    sp = synth.make_problem_fast(N, p, K, seed=31)
    plan = kg.Plan(sp.H, p, K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, seed=sp.seed)
    info = plan.info()
    assert info["segment_loci"] < p, "this test is meant to run the segmented path (sweep 0)"
    plan.run()
    hk, _ = plan.download()
    plan.close()
    G = 32 // info["lanes_per_hap"]
    slots = info["grid"] * (info["block"] // 32)
    rng = np.random.default_rng(5)
    rows = {0, 1, G - 1, G, G + 1, 2 * G - 1, N - 1, N - G, N - G - 1, (slots - 1) * G, min(N - 1, slots * G - 1),
            min(N - 1, slots * G)}
    rows |= {int(x) for x in rng.choice(N, size=40 - len(rows), replace=False)}
    rows = np.array(sorted(r for r in rows if 0 <= r < N), dtype=np.int32)
    want = oracle_unrelated_rows(_problem(sp), rows, chunk=2)
    bad = rows[(hk[rows] != want).any(axis=1)]
    assert len(bad) == 0, f"rows differing from the oracle at p = 60 000 ({info['segment_loci']} loci per segment): {bad.tolist()}"


def test_c5_shape_mixed_call_matches_oracle():
    from knockoffgwas_b200 import synth
    N, p, K = 600, 60000, 100
    sp = synth.make_problem_fast(N, p, K, seed=41, n_founders=200, mean_group=3.0)
    synth.add_ibd_pairs(sp, 16)
    for size in (3, 4, 5, 6, 4):
        synth.add_ibd_families(sp, 1, size=size, segments=4, seed=9 + size)
    sizes = [len(m) for m, _ in sp.families]
    assert sizes.count(2) >= 16 and sum(1 for s in sizes if s >= 3) >= 4
    related = np.concatenate([m for m, _ in sp.families])
    unrelated = np.setdiff1d(np.arange(N, dtype=np.int32), related).astype(np.int32)
    hk, dbg = kg.generate(sp.H, p, K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, unrelated=unrelated, ref_global=sp.ref,
                          seed=sp.seed, families=sp.families, debug_paths=True)
    prob = _problem(sp)
    row = 0
    for (mem, _), r in zip(sp.families, oracle_families(prob, sp.families)):
        n = len(mem)
        assert np.array_equal(dbg.fam_z[row:row + n], r.z), f"Z of family {mem.tolist()}"
        assert np.array_equal(dbg.fam_zk[row:row + n], r.zk), f"Zk of family {mem.tolist()}"
        assert np.array_equal(hk[mem], r.hk), f"Hk of family {mem.tolist()}"
        row += n
    pick = unrelated[:: max(1, len(unrelated) // 24)]
    want = oracle_unrelated_rows(prob, pick, chunk=2)
    assert np.array_equal(hk[pick], want), "unrelated rows of the mixed call differ from the oracle"


def test_c4_six_partition_levels_on_one_plan():
    """configs[3]: six resolutions (single-SNP groups up to ~50 SNPs per group) through kgw_plan_set_groups on ONE
    plan -- H, references and scratch stay resident, only the per-locus knockoff tables change
    (Knockoffs::set_partition, knockoffs.cpp:91-107; the loop of main.cpp:159-181)."""
    from knockoffgwas_b200 import synth
    N, p, K = 3000, 4000, 50
    sp = synth.make_problem_fast(N, p, K, seed=51)
    synth.add_ibd_pairs(sp, 6)
    related = np.concatenate([m for m, _ in sp.families])
    unrelated = np.setdiff1d(np.arange(N, dtype=np.int32), related).astype(np.int32)
    rows = unrelated[:: len(unrelated) // 48]
    grng = np.random.default_rng(77)
    plan = kg.Plan(sp.H, p, K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, unrelated=unrelated, ref_global=sp.ref,
                   seed=sp.seed, families=sp.families)
    seen = []
    for level, mean_group in enumerate((1.0, 2.0, 5.0, 10.0, 25.0, 50.0)):
        groups = synth.make_groups(p, mean_group, grng)
        plan.set_groups(groups)
        plan.run()
        hk, _ = plan.download()
        prob = _problem(sp, groups)
        want = oracle_unrelated_rows(prob, rows, chunk=4)
        assert np.array_equal(hk[rows], want), f"level {level} (mean group {mean_group}): unrelated rows differ from the oracle"
        for (mem, _), r in zip(sp.families, oracle_families(prob, sp.families)):
            assert np.array_equal(hk[mem], r.hk), f"level {level}: related pair {mem.tolist()} differs from the oracle"
        seen.append(hk[rows].copy())
    plan.close()
    assert any(not np.array_equal(seen[0], s) for s in seen[1:]), "the partition had no effect on the knockoffs"
