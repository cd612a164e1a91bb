#!/usr/bin/env python
"""Development helper: run the C2 workload several times per kernel form and report how many rows differ between runs."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth
N, P, K = int(os.environ.get("N", 20000)), int(os.environ.get("P", 10000)), 50
sp = synth.make_problem_fast(N, P, K, seed=2020)
def run():
    plan = kg.Plan(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, seed=sp.seed)
    plan.run(); hk, _ = plan.download(); plan.close(); return hk
ref = None
for form in ("0", "1"):
    os.environ["KGW_FORCE_WIDE"] = form
    for it in range(int(os.environ.get("REPS", 4))):
        hk = run()
        if ref is None: ref = hk
        bad = np.flatnonzero((hk != ref).any(axis=1))
        print(f"form wide={form} run {it}: {len(bad)} rows differ from the first run", bad[:8].tolist(), flush=True)
