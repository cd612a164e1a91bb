#!/usr/bin/env python
"""Golden vectors for reference selection (KinshipComputer::assign_references, kinship_computer.cpp:837-931):
runs the UNMODIFIED reference with --compute-references on the toy chr22 data and stores what
assign_references reads -- H, the kinship clusters in the order of `<out>_clust.txt` (kinship_computer.cpp:693-711),
metadata.related_families per haplotype, the windows -- and what it returned, as written to
`<out>_lref.txt` / `<out>_ref.txt` (kinship.cpp:355-414).  Two cases: several clusters with one window, and
the leave-window-out distances of --windows.  Output: tests/golden/refs_*.npz (committed).
Needs /root/reference and oracle/_ref: build container only."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1] / "oracle"))
import kgw_oracle as ko  # noqa: E402
from make_golden import DATA, PROBE  # noqa: E402


def read_refs(path: Path):
    lines = [ln for ln in open(path)]
    N, W = int(lines[0].split()[1]), int(lines[1].split()[1])
    body = [ln for ln in lines[2:] if not ln.startswith("#")]
    K = len(body[0].split())
    return np.array([[int(x) for x in ln.split()] for ln in body], dtype=np.int32).reshape(N, W, K)


def run(tmp: Path, tag: str, K: int, cmin: int, cmax: int, windows: str):
    out = tmp / tag
    out.mkdir(parents=True)
    args = [str(PROBE), "--bgen", str(DATA / "haplotypes/example_chr22"), "--keep", str(DATA / "qc/samples_qc.txt"),
            "--extract", str(DATA / "qc/qc_chr22.txt"), "--map", str(DATA / "maps/genetic_map_chr22.txt"),
            "--part", str(DATA / "partitions/example_chr22.txt"), "--ibd", str(DATA / "ibd/example_chr22.txt"),
            "--K", str(K), "--cluster_size_min", str(cmin), "--cluster_size_max", str(cmax), "--n_threads", "1",
            "--seed", "2020", "--compute-references", "--generate-knockoffs", "--resolution", "6",
            "--hmm-rho", "1", "--hmm-lambda", "1e-3", "--windows", windows, "--out", str(out / "ko_chr22")]
    env = dict(os.environ, KGW_PROBE_DIR=str(out / "dump"), KGW_PROBE_TRACE="0")
    r = subprocess.run(args, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
        raise SystemExit("reference run failed: " + tag)
    prob, _, _ = ko.load_probe_dump(out / "dump" / "call0")
    clusters = [np.array([int(x) for x in ln.split()], dtype=np.int32) for ln in open(out / "ko_chr22_clust.txt") if ln.strip()]
    lref = read_refs(out / "ko_chr22_lref.txt")
    gref = read_refs(out / "ko_chr22_ref.txt")[:, 0, :]
    fam_sizes = np.array([len(f) for f in prob.families], dtype=np.int32)
    fam_members = np.concatenate(prob.families).astype(np.int32)
    np.savez_compressed(HERE / f"refs_{tag}.npz", p=prob.p, K=K, H=prob.H, win_start=prob.win_start, win_end=prob.win_end,
                        clust_sizes=np.array([len(c) for c in clusters], dtype=np.int32), clust_members=np.concatenate(clusters),
                        fam_sizes=fam_sizes, fam_members=fam_members, lref=lref, gref=gref)
    print(f"refs_{tag}: N={prob.N} p={prob.p} K={K} W={prob.W} clusters={[len(c) for c in clusters]} "
          f"related={int((fam_sizes > 1).sum())}")


def main():
    tmp = Path(tempfile.mkdtemp(prefix="kgw_golden_refs_"))
    try:
        run(tmp, "toy_chr22_K10_clusters", 10, 200, 600, "0")
        run(tmp, "toy_chr22_K10_windows", 10, 500, 2000, "1000000")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
