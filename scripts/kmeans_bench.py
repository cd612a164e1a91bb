#!/usr/bin/env python
"""Development helper: one iteration of the bifurcating k-means' data passes (kgw_kinship_distances + kgw_kinship_centroids)
on a thinned matrix of biobank width.
usage: python scripts/kmeans_bench.py [--haps N] [--snps P]"""
import argparse, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import knockoffgwas_b200 as kg

ap = argparse.ArgumentParser()
ap.add_argument("--haps", type=int, default=700000)
ap.add_argument("--snps", type=int, default=6000)
a = ap.parse_args()
N, P = a.haps, a.snps
rng = np.random.default_rng(3)
H = rng.integers(0, 256, size=(N, (P + 7) // 8), dtype=np.uint8)
t0 = time.time()
km = kg.KinshipMatrix(H, P)
print(f"{N} haplotypes x {P} thinned SNPs uploaded in {time.time() - t0:.2f} s", flush=True)
members = np.arange(N, dtype=np.int32)
mu0, mu1 = km.centroids(members, (members % 2).astype(np.int32))
for it in range(3):
    t0 = time.perf_counter()
    d0, d1 = km.distances(members, mu0, mu1)
    t1 = time.perf_counter()
    mu0, mu1 = km.centroids(members, (d0 >= d1).astype(np.int32))
    t2 = time.perf_counter()
    print(f"iteration {it}: distances {1e3 * (t1 - t0):.1f} ms, centroids {1e3 * (t2 - t1):.1f} ms (calls, transfers included)", flush=True)
km.close()
