#!/usr/bin/env python
"""Development helper: device time of the .bed body kernel (kgw_plan_encode_bed) on random matrices of a given size.
usage: python scripts/bed_bench.py [--haps N] [--snps p]"""
import argparse, sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import knockoffgwas_b200 as kg

ap = argparse.ArgumentParser()
ap.add_argument("--haps", type=int, default=87500)
ap.add_argument("--snps", type=int, default=60000)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
N, p, K = a.haps, a.snps, 8
rng = np.random.default_rng(1)
rb = (p + 7) // 8
H = rng.integers(0, 256, size=(N, rb), dtype=np.uint8)
ref = rng.integers(0, N, size=(N, 1, K), dtype=np.int32)
b = np.full(p, 0.9); mu = np.full(p, 1e-3); groups = np.arange(p, dtype=np.int32)
plan = kg.Plan(H, p, K, ref, b, mu, groups, unrelated=np.zeros(0, np.int32), seed=1)   # H given as packed rows
body, _ = plan.encode_bed()
ms = []
for _ in range(a.iters):
    body, t = plan.encode_bed(out=body)
    ms.append(t)
tot = body.nbytes + 2 * N * rb
print(f"{N} x {p}: kernel {min(ms):.3f} ms (median {np.median(ms):.3f}), {tot / (min(ms) * 1e-3) / 1e12:.2f} TB/s of {tot / 1e9:.2f} GB")
