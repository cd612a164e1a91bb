#!/usr/bin/env python
"""Development helper: device time of the related (IBD) branch on synthetic pairs (synth.add_ibd_pairs,
SURVEY.md 8d C5 recipe), next to the unrelated branch on the same haplotypes; optional check of the
first families against the CPU oracle.
usage: python scripts/fam_bench.py [--haps N] [--snps p] [--K K] [--pairs P] [--check C] [--mean-group g]"""
import argparse, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle")); sys.path.insert(0, str(REPO / "tests"))
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--haps", type=int, default=4000)
ap.add_argument("--snps", type=int, default=10000)
ap.add_argument("--K", type=int, default=50)
ap.add_argument("--pairs", type=int, default=400)
ap.add_argument("--check", type=int, default=0, help="families compared with the oracle")
ap.add_argument("--mean-group", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--unrelated", action="store_true", help="also time the unrelated branch on the same haplotypes")
ap.add_argument("--mixed", type=int, default=0, help="also time both branches in one call (that many unrelated haplotypes), one after the other and overlapped")
a = ap.parse_args()

t0 = time.time()
sp = synth.make_problem(a.haps, a.snps, a.K, seed=3, n_founders=200, mean_group=a.mean_group)
synth.add_ibd_pairs(sp, a.pairs)
fams = sp.families
nseg = np.bincount([len(s) for _, s in fams])
members = np.concatenate([m for m, _ in fams])
print(f"{a.haps} haplotypes x {a.snps} SNPs, K={a.K}, {len(fams)} pairs, segments per pair histogram {nseg.tolist()}, "
      f"built in {time.time() - t0:.1f} s", flush=True)
ref3 = sp.ref[:, None, :]
runs = [("related only", np.zeros(0, np.int32), fams)]
if a.unrelated:
    runs.append(("unrelated only", np.sort(members).astype(np.int32), None))
if a.mixed:
    rest = np.setdiff1d(np.arange(sp.N), members)[:a.mixed].astype(np.int32)
    runs.append(("unrelated part", rest, None))
    runs.append(("both, sequential", rest, fams))
    runs.append(("both, overlapped", rest, fams))
import os
for label, unrel, fam in runs:
    if label.startswith("both"):
        os.environ["KGW_FAMILY_OVERLAP"] = "1" if "overlapped" in label else "0"
    plan = kg.Plan(sp.H, sp.p, sp.K, ref3, sp.b, sp.mu, sp.groups, unrelated=unrel, ref_global=sp.ref, seed=sp.seed, families=fam)
    plan.run(); plan.sync()
    ms = plan.run_timed(a.iters)
    n = (len(members) if fam else 0) + (len(unrel) if (not fam or label.startswith("both")) else 0)
    if label == "unrelated only":
        n = len(unrel)
    print(f"  {label:16s} {ms:10.3f} ms  {n * sp.p / (ms * 1e-3):.3e} hap-SNPs/s  ({n} haplotypes)", flush=True)
    if fam and a.check:
        import kgw_oracle as ko
        hk, dbg = plan.download(debug_paths=True)
        prob = ko.Problem(H=sp.H, p=sp.p, K=sp.K, ref_local=ref3, ref_global=sp.ref, b=sp.b, mu=sp.mu, groups=sp.groups,
                          win_start=None, win_end=None, seed=sp.seed)
        rng, _ = ko.make_rng(prob, mode="philox")
        row, bad = 0, 0
        for f, (mem, segs) in enumerate(fams[:a.check]):
            r = ko.run_family(prob, mem, [(s[0], s[1], 0, 0, s[2]) for s in segs], rng)
            ok = (np.array_equal(dbg.fam_z[row:row + 2], r.z) and np.array_equal(dbg.fam_zk[row:row + 2], r.zk)
                  and np.array_equal(hk[mem], r.hk))
            bad += 0 if ok else 1
            if not ok:
                dz = np.argwhere(dbg.fam_z[row:row + 2] != r.z)
                print(f"    family {f} {mem.tolist()} differs from the oracle; first Z mismatches {dz[:4].tolist()} of {len(dz)}")
            row += 2
        print(f"  oracle check: {a.check - bad} of {a.check} families bit-exact (Z, Zk, Hk)", flush=True)
    plan.close()
