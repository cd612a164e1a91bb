#!/usr/bin/env python
"""Development helper: sampler time by lanes per haplotype (unrelated shard, K and N from the environment)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth
N, p, K = int(os.environ.get("N", 19688)), int(os.environ.get("P", 60000)), int(os.environ.get("K", 100))
prob = synth.make_problem_fast(N, p, K, seed=2022)
for lanes in [int(x) for x in os.environ.get("LANES", "0,2,4").split(",")]:
    for wide in os.environ.get("WIDE", "x").split(","):
        if wide in "01": os.environ["KGW_FORCE_WIDE"] = wide
        else: os.environ.pop("KGW_FORCE_WIDE", None)
        pl = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, options=kg.Options(lanes_per_hap=lanes))
        pl.run(); pl.sync()
        ms = pl.run_timed(1)
        i = pl.info()
        print(f"N={N} K={K} lanes={lanes} wide={wide}: {ms:.1f} ms  {N*p/ms/1e6:.2f}e9 hap-SNP/s  L={i['lanes_per_hap']} S={i['states_per_lane']} grid={i['grid']} seg={i['segment_loci']}", flush=True)
        pl.close(); kg.release_workspace()
