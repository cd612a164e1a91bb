#!/usr/bin/env python
"""Development helper: small reproducer (several runs of one plan, rows differing from the first run)."""
import os, sys
from pathlib import Path
import numpy as np
R = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(R)]
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth
N, P, K = int(os.environ.get("N", 4000)), int(os.environ.get("P", 512)), 50
sp = synth.make_problem_fast(N, P, K, seed=2020, n_founders=200)
plan = kg.Plan(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, seed=sp.seed)
outs = []
for it in range(int(os.environ.get("REPS", 3))):
    plan.run()
    hk, _ = plan.download()
    outs.append(hk.copy())
    print(f"run {it}: {int((hk != outs[0]).any(axis=1).sum())} rows differ from run 0", flush=True)
print(plan.info())
plan.close()
