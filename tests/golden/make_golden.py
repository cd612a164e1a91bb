#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs oracle/_ref/snpknock2_probe (reference objects + the --wrap probe of
oracle/ref_shims/wrap_probe.cpp) on the reference's own toy data
(/root/reference/data, the inputs of examples/example_1_generate_knockoffs.sh)
with --n_threads 1, and stores for every Knockoffs::generate() call:
  * the inputs exactly as the unchanged host code hands them to the hot path
    (H, references, b, mu, groups, windows, unrelated / family lists, IBD segments),
  * the haplotype-level knockoffs Hk the reference produced,
  * the latent paths z / zk of the unrelated haplotypes and the raw draw trace of
    the related ones (from the weighted_choice trace).

Needs /root/reference and oracle/_ref (run `bash oracle/build_ref.sh` first); it
therefore only runs in the build container.  The .npz outputs are committed.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "oracle"))
import kgw_oracle as ko  # noqa: E402

REF = Path(os.environ.get("KGW_REFERENCE_DIR", "/root/reference"))
DATA = REF / "data"
PROBE = REPO / "oracle" / "_ref" / "snpknock2_probe"
OUT = Path(__file__).resolve().parent


def run_ref(tmp: Path, chrom: int, K: int, extra: list[str], tag: str, refs_from: Path | None):
    out = tmp / f"{tag}"
    out.mkdir(parents=True, exist_ok=True)
    args = [
        str(PROBE),
        "--bgen", str(DATA / f"haplotypes/example_chr{chrom}"),
        "--keep", str(DATA / "qc/samples_qc.txt"),
        "--extract", str(DATA / f"qc/qc_chr{chrom}.txt"),
        "--map", str(DATA / f"maps/genetic_map_chr{chrom}.txt"),
        "--part", str(DATA / f"partitions/example_chr{chrom}.txt"),
        "--ibd", str(DATA / f"ibd/example_chr{chrom}.txt"),
        "--K", str(K), "--cluster_size_min", "1000", "--cluster_size_max", "10000",
        "--n_threads", "1", "--seed", "2020", "--generate-knockoffs",
        "--out", str(out / f"ko_chr{chrom}"),
    ] + extra
    if refs_from is None:
        args += ["--compute-references"]
    else:
        args += ["--ref", str(refs_from / f"ko_chr{chrom}_ref.txt"), "--lref", str(refs_from / f"ko_chr{chrom}_lref.txt")]
    env = dict(os.environ, KGW_PROBE_DIR=str(out / "dump"), KGW_PROBE_TRACE="1")
    r = subprocess.run(args, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
        raise SystemExit(f"reference run failed: {tag}")
    return out


def save_call(dump_dir: Path, name: str, note: str):
    prob, Hk, trace = ko.load_probe_dump(dump_dir)
    nu, p = len(prob.unrelated), prob.p
    # draw layout per unrelated haplotype (generate_unrelated_worker, knockoffs.cpp:654-681):
    # for each window: Z over the padded range, Zk over the loci of the padded range's groups,
    # Hk over the core range.  Later windows overwrite the padded overlap of z / zk.
    W = prob.W
    grp_first = {}
    grp_last = {}
    for j, g in enumerate(prob.groups):
        grp_first.setdefault(int(g), j)
        grp_last[int(g)] = j
    layout = []
    for w in range(W):
        js = int(prob.win_start[w - 1] if w > 0 else prob.win_start[w])
        je = int(prob.win_end[w + 1] if w < W - 1 else prob.win_end[w])
        ks = grp_first[int(prob.groups[js])]
        ke = grp_last[int(prob.groups[je - 1])] + 1
        layout.append((js, je, ks, ke, int(prob.win_start[w]), int(prob.win_end[w])))
    per_hap = sum((je - js) + (ke - ks) + (ce - cs) for (js, je, ks, ke, cs, ce) in layout)
    n_unrel_draws = nu * per_hap
    tr_u = trace[len(trace) - n_unrel_draws:].reshape(nu, per_hap, 2)
    z = np.zeros((nu, p), dtype=np.uint8)
    zk = np.zeros((nu, p), dtype=np.uint8)
    hk_tr = np.zeros((nu, p), dtype=np.uint8)
    t = 0
    for (js, je, ks, ke, cs, ce) in layout:
        assert (tr_u[:, t:t + je - js, 0] == prob.K).all()
        z[:, js:je] = tr_u[:, t:t + je - js, 1]; t += je - js
        assert (tr_u[:, t:t + ke - ks, 0] == prob.K).all()
        zk[:, ks:ke] = tr_u[:, t:t + ke - ks, 1]; t += ke - ks
        assert (tr_u[:, t:t + ce - cs, 0] == 2).all()
        hk_tr[:, cs:ce] = tr_u[:, t:t + ce - cs, 1]; t += ce - cs
    assert (hk_tr == ko.unpack_bits(Hk[prob.unrelated], p)).all()
    tr_rel = trace[: len(trace) - n_unrel_draws].astype(np.int16)
    fam_sizes = np.array([len(f) for f in prob.families], dtype=np.int32)
    fam_members = np.concatenate(prob.families).astype(np.int32) if prob.families else np.zeros(0, np.int32)
    segs = []
    for (hap, jmin, jmax, bpmin, bpmax, ids) in prob.ibd_segments:
        segs += [hap, jmin, jmax, bpmin, bpmax, len(ids)] + [int(x) for x in ids]
    np.savez_compressed(
        OUT / f"{name}.npz",
        note=np.array(note), p=p, K=prob.K, W=prob.W, seed=prob.seed,
        H=prob.H, Hk=Hk, ref_local=prob.ref_local, ref_global=prob.ref_global, b=prob.b, mu=prob.mu,
        groups=prob.groups, win_start=prob.win_start, win_end=prob.win_end, unrelated=prob.unrelated,
        fam_sizes=fam_sizes, fam_members=fam_members, ibd_segments=np.array(segs, dtype=np.int32),
        bp=prob.bp, z_unrelated=z, zk_unrelated=zk, trace_related=tr_rel,
        fam_clusters=np.fromfile(dump_dir / "fam_clusters.i32", dtype=np.int32))
    print(f"  {name}.npz  N={prob.N} p={p} K={prob.K} W={prob.W} groups={prob.groups[-1] + 1} "
          f"unrelated={nu} related_draws={len(tr_rel)}")


def main():
    if not PROBE.exists():
        raise SystemExit("build oracle/_ref first: bash oracle/build_ref.sh")
    tmp = Path(tempfile.mkdtemp(prefix="kgw_golden_"))
    try:
        hmm = ["--hmm-rho", "1", "--hmm-lambda", "1e-3", "--windows", "0"]
        # (1) toy chr22, K=10, all seven resolutions (module_2 settings, example_1 script)
        a = run_ref(tmp, 22, 10, hmm, "chr22_K10", None)
        for r, call in ((0, 0), (2, 2), (4, 4), (6, 6)):
            save_call(a / "dump" / f"call{call}", f"toy_chr22_K10_res{r}",
                      f"example_1 settings, chr22, K=10, resolution {r}, --n_threads 1")
        # (2) toy chr21, K=50 (the metric's K), finest and a coarse resolution
        b = run_ref(tmp, 21, 50, hmm, "chr21_K50", None)
        for r in (0, 5):
            save_call(b / "dump" / f"call{r}", f"toy_chr21_K50_res{r}",
                      f"chr21, K=50, resolution {r}, --n_threads 1")
        # (3) multi-window mode: --windows 1000000 bp on chr22, K=10
        c = run_ref(tmp, 22, 10, ["--hmm-rho", "1", "--hmm-lambda", "1e-3", "--windows", "1000000",
                                  "--resolution", "1"], "chr22_K10_win", None)
        save_call(c / "dump" / "call0", "toy_chr22_K10_res1_windows",
                  "chr22, K=10, --windows 1000000, resolution 1, --n_threads 1")
        # (4) --hmm file path (load_hmm: b_0 = exp(-rec_0), mu clamped), re-using (1)'s references
        hmm_file = a / "ko_chr22_hmm.txt"
        d = run_ref(tmp, 22, 10, ["--hmm", str(hmm_file), "--windows", "0", "--resolution", "3"],
                    "chr22_K10_hmmfile", a)
        save_call(d / "dump" / "call0", "toy_chr22_K10_res3_hmmfile",
                  "chr22, K=10, --hmm <file written by run 1>, resolution 3, --n_threads 1")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
