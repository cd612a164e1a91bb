"""Shared helpers for the test-suite: golden fixture loading and synthetic problems."""
from __future__ import annotations

from pathlib import Path

import numpy as np

import kgw_oracle as ko

GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_NAMES = sorted(pth.stem for pth in GOLDEN.glob("toy_*.npz"))


def load_golden(name: str):
    """Returns (problem, fixture dict).  The fixture was produced by tests/golden/make_golden.py
    from the unmodified reference binary."""
    f = np.load(GOLDEN / f"{name}.npz")
    fam, t = [], 0
    for s in f["fam_sizes"]:
        fam.append(f["fam_members"][t:t + s].astype(np.int32))
        t += s
    raw, segs, t = f["ibd_segments"], [], 0
    while t < len(raw):
        hap, jmin, jmax, bpmin, bpmax, n = (int(x) for x in raw[t:t + 6])
        segs.append((hap, jmin, jmax, bpmin, bpmax, raw[t + 6:t + 6 + n].copy()))
        t += 6 + n
    prob = ko.Problem(H=f["H"], p=int(f["p"]), K=int(f["K"]), ref_local=f["ref_local"], ref_global=f["ref_global"],
                      b=f["b"], mu=f["mu"], groups=f["groups"], win_start=f["win_start"], win_end=f["win_end"],
                      unrelated=f["unrelated"], families=fam, ibd_segments=segs, bp=f["bp"], seed=int(f["seed"]),
                      fam_clusters=ko.parse_fam_clusters(f["fam_clusters"]) if "fam_clusters" in f else [])
    return prob, f


def synthetic_problem(N=64, p=200, K=8, seed=0, mean_group=1.0, rho=1.0, lam=1e-3, n_founders=12,
                      mean_dist_cm=0.05, tiny_dist_frac=0.0, W=1):
    """Small Li-Stephens-mosaic problem in the layout of SURVEY.md A.7 (numpy default_rng(seed))."""
    rng = np.random.default_rng(seed)
    freq = rng.uniform(0.05, 0.5, size=p)
    founders = (rng.random((n_founders, p)) < freq).astype(np.uint8)
    d = rng.exponential(mean_dist_cm, size=p)
    if tiny_dist_frac > 0:
        d[rng.random(p) < tiny_dist_frac] = 1e-9
    d = np.maximum(d, 1e-9)
    bits = np.zeros((N, p), dtype=np.uint8)
    for i in range(N):
        cur = rng.integers(n_founders)
        sw = rng.random(p) < (1.0 - np.exp(-d))
        for j in range(p):
            if sw[j]:
                cur = rng.integers(n_founders)
            bits[i, j] = founders[cur, j]
    bits ^= (rng.random((N, p)) < 1e-3).astype(np.uint8)
    H = ko.pack_bits(bits)
    ref = np.zeros((N, K), dtype=np.int32)
    for i in range(N):
        cand = np.setdiff1d(np.arange(N), [i])
        ref[i] = np.sort(rng.choice(cand, size=K, replace=False))
    b = np.exp(-rho * d)
    b[0] = 0.0
    mu = np.full(p, lam)
    if mean_group <= 1.0:
        groups = np.arange(p, dtype=np.int32)
    else:
        cuts = rng.random(p) < 1.0 / mean_group
        cuts[0] = False
        groups = np.cumsum(cuts).astype(np.int32)
    if W > 1:
        # window starts at group boundaries, as metadata.cpp:86-108 guarantees
        cand = np.flatnonzero(np.diff(groups) > 0) + 1
        starts = np.sort(rng.choice(cand, size=W - 1, replace=False))
        win_start = np.concatenate([[0], starts]).astype(np.int32)
        win_end = np.concatenate([starts, [p]]).astype(np.int32)
        ref_local = np.stack([ref] * W, axis=1)
        for w in range(1, W):
            for i in range(N):
                cand_i = np.setdiff1d(np.arange(N), [i])
                ref_local[i, w] = np.sort(rng.choice(cand_i, size=K, replace=False))
    else:
        win_start, win_end, ref_local = None, None, ref[:, None, :]
    return ko.Problem(H=H, p=p, K=K, ref_local=ref_local, ref_global=ref, b=b, mu=mu, groups=groups,
                      win_start=win_start, win_end=win_end, seed=2020 + seed)


def host_threads() -> int:
    import os
    return max(1, min(64, (os.cpu_count() or 1)))


def oracle_unrelated_rows(prob, haps, chunk=32, mode="philox"):
    """Packed knockoff rows of `haps` from the CPU oracle, run on every host core (the ctypes call releases the
    GIL; the keyed generator makes the chunks independent).  Returns uint8 [len(haps)][row_bytes]."""
    from concurrent.futures import ThreadPoolExecutor
    haps = np.ascontiguousarray(haps, dtype=np.int32)
    out = np.zeros((len(haps), prob.row_bytes), dtype=np.uint8)

    def work(c0):
        out[c0:c0 + chunk] = ko.run_unrelated(prob, haps=haps[c0:c0 + chunk], mode=mode).hk

    with ThreadPoolExecutor(max_workers=host_threads()) as ex:
        list(ex.map(work, range(0, len(haps), chunk)))
    return out


def oracle_families(prob, families):
    """[(FamilyResult)] of the CPU oracle for `families` = [(members, [(j_min, j_max, ids), ...])], one host thread
    per family (keyed generator: no state is carried from one family to the next)."""
    from concurrent.futures import ThreadPoolExecutor

    def work(fam):
        mem, segs = fam
        rng, _ = ko.make_rng(prob, mode="philox")
        return ko.run_family(prob, mem, [(s[0], s[1], 0, 0, s[-1]) for s in segs], rng)

    with ThreadPoolExecutor(max_workers=host_threads()) as ex:
        return list(ex.map(work, families))
