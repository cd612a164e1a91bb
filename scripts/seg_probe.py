#!/usr/bin/env python
"""Development helper: time of the sampler kernel against the segment length (C5-shaped unrelated shard)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth
N, p, K = int(os.environ.get("N", 19688)), 60000, int(os.environ.get("K", 100))
prob = synth.make_problem_fast(N, p, K, seed=2022)
for wide in ("0", "1"):
    os.environ["KGW_FORCE_WIDE"] = wide
    for seg in (0, 30016, 15008, 7520, 3744):
        pl = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, options=kg.Options(segment_loci=seg))
        pl.run(); pl.sync()
        ms = pl.run_timed(1)
        i = pl.info()
        print("wide=" + wide, "seg", seg, "->", i["segment_loci"], round(ms, 1), "ms  grid", i["grid"], "scratch GB", round(i["scratch_bytes"] / 1e9, 1), flush=True)
        pl.close()
        kg.release_workspace()
