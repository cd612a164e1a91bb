#!/usr/bin/env python
"""Development helper (--K 50 --pairs-per-gpu 0: the same for configs[2], 700k unrelated haplotypes): one GPU's FULL eighth of BASELINE.json configs[4] (350k samples = 700k haplotypes, K=100, 60 000
SNPs, 10 % related): 78 750 unrelated haplotypes + 4 375 IBD pairs in ONE generate() call (bench.py's c5_shard leg runs a
quarter of this so that the default run stays within minutes).
usage: python scripts/c5_full_shard.py [--scale 1.0]"""
import argparse, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--parts", action="store_true", help="also time the two branches on their own")
ap.add_argument("--K", type=int, default=100)
ap.add_argument("--pairs-per-gpu", type=int, default=4375, help="IBD pairs per GPU's eighth (0: configs[2], unrelated haplotypes only)")
ap.add_argument("--devices", type=int, default=1, help="GPUs of this process (kgw_options.devices); with --scale 8: the whole of configs[4]")
a = ap.parse_args()
N, p, K = int(87500 * a.scale) // 2 * 2, 60000, a.K
pairs = int(a.pairs_per_gpu * a.scale)
t0 = time.time()
prob = synth.make_problem_fast(N, p, K, seed=2022, rho=1.0, lam=1e-3)
if pairs > 0:
    synth.add_ibd_pairs(prob, pairs)
fams = prob.families if pairs > 0 else None
related = np.concatenate([m for m, _ in prob.families]) if pairs > 0 else np.zeros(0, np.int32)
unrelated = np.setdiff1d(np.arange(N, dtype=np.int32), related).astype(np.int32)
print(f"{len(unrelated)} unrelated + {len(related)} related haplotypes ({pairs} pairs) x {p} SNPs, K={K}; built in {time.time() - t0:.0f} s", flush=True)
opts = kg.Options(devices=tuple(range(a.devices))) if a.devices > 1 else None
t0 = time.time()
plan = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, unrelated=unrelated, ref_global=prob.ref,
               seed=prob.seed, families=fams, options=opts)
print(f"plan created in {time.time() - t0:.1f} s ({a.devices} device(s))", flush=True)
info = plan.info()
plan.run(); plan.sync()
ms = plan.run_timed(2)   # CUDA events on every device's stream, the slowest device counts (kgw_plan_run_timed)
print(f"one generate() call: {ms:.1f} ms = {N * p / (ms * 1e-3):.3e} hap-SNPs/s; launches per call {info['launches_per_run']}, "
      f"sampler {info['lanes_per_hap']} lanes x {info['states_per_lane']} states, segment {info['segment_loci']} loci, scratch {info['scratch_bytes'] / 1e9:.1f} GB", flush=True)
t0 = time.time()
hk, _ = plan.download()
print(f"download of the {hk.nbytes / 1e9:.2f} GB of knockoff rows: {time.time() - t0:.1f} s", flush=True)
plan.close()
if a.parts:
    for name, kw in (("unrelated only", dict(unrelated=unrelated)),
                     ("related only", dict(unrelated=np.zeros(0, np.int32), ref_global=prob.ref, families=prob.families))):
        pl = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, **kw)
        pl.run(); pl.sync()
        t = pl.run_timed(1)
        i2 = pl.info()
        n = len(unrelated) if name.startswith("unrel") else len(related)
        print(f"  {name}: {t:.1f} ms = {n * p / (t * 1e-3):.3e} hap-SNPs/s (launches {i2['launches_per_run']}, segment {i2['segment_loci']})", flush=True)
        pl.close()
    kg.release_workspace()
filled = int((hk[np.concatenate([unrelated, related])] != 0).any(axis=1).sum())
print(f"rows written: {filled} of {N}")
