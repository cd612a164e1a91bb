"""TEST / BENCH INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference binary built into
oracle/_ref/ (oracle/build_ref.sh) on synthetic inputs.

Writes a problem in the reference's own input formats (SURVEY.md Appendix B:
`--haps` PREFIX.haps/.legend/.sample haps_reader.cpp:37-132, `--map` metadata.cpp:431-439,
`--part` metadata.cpp:495-543, `--ref/--lref` kinship.cpp:35-231), runs
`snpknock2_probe ... --generate-knockoffs --n_threads T`, and returns the wall time of
Knockoffs::generate() as measured by the std::chrono wrapper of oracle/ref_shims/wrap_probe.cpp.
Only bench.py's cpu_baseline / `--impl reference` legs and tests/ may import this module.
"""
from __future__ import annotations

import os
import re
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PROBE = HERE / "_ref" / "snpknock2_probe"
TIMER = HERE / "_ref" / "snpknock2_timer"   # timer around Knockoffs::generate only (no weighted_choice wrapper): the timed arm


def timed_binary() -> Path:
    return TIMER if TIMER.exists() and os.access(TIMER, os.X_OK) else PROBE


def available() -> bool:
    return PROBE.exists() and os.access(PROBE, os.X_OK)


def write_inputs(d: Path, H: np.ndarray, p: int, ref: np.ndarray, groups: np.ndarray, cm: np.ndarray,
                 bp: np.ndarray, ref_local: np.ndarray | None = None) -> dict:
    """H: uint8 [N][row_bytes] packed rows (N even: haplotypes 2s, 2s+1 belong to sample s)."""
    d = Path(d)
    d.mkdir(parents=True, exist_ok=True)
    N = H.shape[0]
    assert N % 2 == 0, "the reference handles diploid samples (two haplotypes each)"
    bits = np.unpackbits(H, axis=1, bitorder="little")[:, :p]          # [N][p]
    # .haps: one row per SNP, one space-separated 0/1 token per haplotype (haps_reader.cpp:107-110)
    txt = np.full((p, 2 * N), ord(" "), dtype=np.uint8)
    txt[:, 0::2] = bits.T + ord("0")
    txt[:, -1] = ord("\n")
    (d / "syn.haps").write_bytes(txt.tobytes())
    ids = [f"rs{j}" for j in range(p)]
    with open(d / "syn.legend", "w") as f:                              # header + id bp a0 a1
        f.write("id position a0 a1\n")
        f.writelines(f"{ids[j]} {int(bp[j])} A G\n" for j in range(p))
    with open(d / "syn.sample", "w") as f:                              # 2 header lines (metadata.cpp:131-147)
        f.write("ID_1 ID_2 missing sex\n0 0 0 D\n")
        f.writelines(f"s{s} s{s} 0 1\n" for s in range(N // 2))
    with open(d / "syn.map", "w") as f:                                 # header + 4 columns (metadata.cpp:431-439)
        f.write("Chromosome Position(bp) Rate(cM/Mb) Map(cM)\n")
        f.writelines(f"chr1 {int(bp[j])} 1.0 {float(cm[j])!r}\n" for j in range(p))
    with open(d / "syn.part", "w") as f:                                # header `SNP res_1`, then `id g`
        f.write("SNP res_1\n")
        f.writelines(f"{ids[j]} {int(groups[j])}\n" for j in range(p))
    K = ref.shape[-1]
    if ref_local is None:
        ref_local = ref.reshape(N, 1, K)
    W = ref_local.shape[1]
    with open(d / "syn_ref.txt", "w") as f:                             # kinship.cpp:140-231
        f.write(f"# {N}\n# 1\n")
        for i in range(N):
            f.write(f"# ID s{i // 2} {i % 2} chr 1\n" + " ".join(map(str, ref[i])) + "\n")
    with open(d / "syn_lref.txt", "w") as f:                            # kinship.cpp:35-138
        f.write(f"# {N}\n# {W}\n")
        for i in range(N):
            for w in range(W):
                f.write(f"# ID s{i // 2} {i % 2} chr 1 window {w}\n" + " ".join(map(str, ref_local[i, w])) + "\n")
    return dict(prefix=str(d / "syn"), map=str(d / "syn.map"), part=str(d / "syn.part"),
                ref=str(d / "syn_ref.txt"), lref=str(d / "syn_lref.txt"))


def run_reference(files: dict, out_dir: Path, *, rho: float, lam: float, n_threads: int, seed: int = 2020,
                  dump_dir: Path | None = None, trace: bool = False, timeout: float = 3600.0) -> dict:
    """Runs the reference CLI the way module_2_knockoffs.sh does (:41-58), with --haps input and
    preloaded references.  Returns {'seconds': generate() wall time, 'threads': T, 'stdout': ...}."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    args = [str(PROBE), "--haps", files["prefix"], "--map", files["map"], "--part", files["part"],
            "--ref", files["ref"], "--lref", files["lref"], "--hmm-rho", repr(float(rho)),
            "--hmm-lambda", repr(float(lam)), "--windows", "0", "--resolution", "0", "--seed", str(seed),
            "--n_threads", str(n_threads), "--generate-knockoffs", "--out", str(out_dir / "ko"),
            "--log", str(out_dir / "log")]
    env = dict(os.environ)
    env.pop("KGW_PROBE_DIR", None)
    env.pop("KGW_PROBE_TRACE", None)
    if dump_dir is not None:
        env["KGW_PROBE_DIR"] = str(dump_dir)
        if trace:
            env["KGW_PROBE_TRACE"] = "1"
    r = subprocess.run(args, env=env, capture_output=True, text=True, timeout=timeout, cwd=str(out_dir))
    if r.returncode != 0:
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    m = re.search(r"KGW_PROBE generate .*threads=(\d+) seconds=([0-9.]+)", r.stderr)
    if not m:
        raise RuntimeError("reference run printed no KGW_PROBE timing line:\n" + r.stderr[-2000:])
    return dict(seconds=float(m.group(2)), threads=int(m.group(1)), stdout=r.stdout)


def run_reference_all_resolutions(files: dict, out_dir: Path, *, rho: float, lam: float, n_threads: int,
                                  seed: int = 2020, timeout: float = 3600.0) -> list:
    """Same CLI without --resolution: the reference calls Knockoffs::generate() once per column of
    the --part file (main.cpp:159-181).  Returns the generate() wall times, one per call."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    args = [str(timed_binary()), "--haps", files["prefix"], "--map", files["map"], "--part", files["part"],
            "--ref", files["ref"], "--lref", files["lref"], "--hmm-rho", repr(float(rho)),
            "--hmm-lambda", repr(float(lam)), "--windows", "0", "--seed", str(seed),
            "--n_threads", str(n_threads), "--generate-knockoffs", "--out", str(out_dir / "ko"),
            "--log", str(out_dir / "log")]
    env = dict(os.environ)
    env.pop("KGW_PROBE_DIR", None)
    env.pop("KGW_PROBE_TRACE", None)
    r = subprocess.run(args, env=env, capture_output=True, text=True, timeout=timeout, cwd=str(out_dir))
    if r.returncode != 0:
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return [float(x) for x in re.findall(r"KGW_PROBE generate .*seconds=([0-9.]+)", r.stderr)]


def time_reference(H, p, ref, groups, cm, bp, *, rho, lam, n_threads, workdir: Path | None = None) -> dict:
    """write_inputs + run_reference in a scratch directory; returns run_reference's dict."""
    with tempfile.TemporaryDirectory(prefix="kgw_ref_", dir=workdir) as tmp:
        files = write_inputs(Path(tmp) / "in", H, p, ref, groups, cm, bp)
        return run_reference(files, Path(tmp) / "out", rho=rho, lam=lam, n_threads=n_threads)
