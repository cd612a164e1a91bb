#!/usr/bin/env python
"""Development helper: device time of the related branch on a golden fixture (families only) next to the
unrelated branch on the same number of haplotypes.  usage: python scripts/fam_timing.py [fixture]"""
import sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle")); sys.path.insert(0, str(REPO / "tests"))
import knockoffgwas_b200 as kg
from kgw_testutil import load_golden
name = sys.argv[1] if len(sys.argv) > 1 else "toy_chr21_K50_res0"
prob, _ = load_golden(name)
fams = [(m, [(s[0], s[1], s[4]) for s in segs]) for m, segs in prob.fam_clusters]
n_rel = sum(len(m) for m, _ in fams)
sizes = np.bincount([len(m) for m, _ in fams])
print(name, "families", len(fams), "members", n_rel, "size histogram", sizes.tolist(), "p", prob.p, "K", prob.K)
for label, unrel, fam in (("related only", np.zeros(0, np.int32), fams), ("unrelated only", prob.unrelated[:n_rel], None)):
    plan = kg.Plan(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, unrelated=unrel,
                   ref_global=prob.ref_global, seed=prob.seed, families=fam)
    plan.run(); plan.sync()
    ms = plan.run_timed(3)
    n = n_rel if fam else len(unrel)
    print(f"  {label:16s} {ms:9.3f} ms  {n * prob.p / (ms * 1e-3):.3e} hap-SNPs/s  info={plan.info()}")
    plan.close()
