"""Deterministic synthetic inputs for the hot path: Li-Stephens mosaic
haplotypes, in-block sorted reference lists, HMM tables as `init_hmm(rho, lambda)` builds them
(snpknock2/src/knockoffs.cpp:1508-1525) and contiguous variant groups.  Used by bench.py and
the tests; nothing here touches the device."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class SynthProblem:
    H: np.ndarray            # uint8 [N][row_bytes], chaplotype layout (LSB-first, chaplotype.h:25)
    p: int
    K: int
    ref: np.ndarray          # int32 [N][K], sorted ascending (knockoffs.cpp:116,140)
    b: np.ndarray            # float64 [p]
    mu: np.ndarray           # float64 [p]
    groups: np.ndarray       # int32 [p]
    cm: np.ndarray           # float64 [p] cumulative genetic position (> 0)
    bp: np.ndarray           # int64 [p] strictly increasing physical position
    seed: int = 2020
    families: list = field(default_factory=list)
    segments: list = field(default_factory=list)

    @property
    def N(self) -> int:
        return int(self.H.shape[0])


def make_groups(p: int, mean_group: float, rng) -> np.ndarray:
    """Contiguous groups with geometric sizes of the given mean (1 = singletons)."""
    if mean_group <= 1.0:
        return np.arange(p, dtype=np.int32)
    cuts = rng.random(p) < 1.0 / mean_group
    cuts[0] = False
    return np.cumsum(cuts).astype(np.int32)


def make_references(N: int, K: int, rng, block: int = 5000) -> np.ndarray:
    """K distinct non-self indices inside the haplotype's block of `block` consecutive
    haplotypes (mimics --cluster_size_max), sorted ascending."""
    ref = np.empty((N, K), dtype=np.int32)
    for b0 in range(0, N, block):
        b1 = min(N, b0 + block)
        m = b1 - b0
        if m - 1 < K:
            raise ValueError(f"block of {m} haplotypes cannot supply K={K} references")
        for i in range(b0, b1):
            c = rng.choice(m - 1, size=K, replace=False)
            c += (c >= i - b0)
            c.sort()
            ref[i] = c + b0
    return ref


def make_problem(n_haps: int, p: int, K: int, seed: int = 2020, mean_group: float = 1.0, rho: float = 1.0,
                 lam: float = 1e-3, n_founders: int = 2000, mean_dist_cm: float = 0.005, flip: float = 1e-3,
                 block: int = 5000, hap_seed: int | None = None, chunk: int = 1024) -> SynthProblem:
    """`seed` fixes everything that is per-SNP (allele frequencies, founders, map, groups);
    `hap_seed` (default = seed) fixes the haplotype rows and their reference lists, so that the
    ranks of a multi-GPU run can draw disjoint shards of one population."""
    rng = np.random.default_rng(seed)
    freq = rng.uniform(0.05, 0.5, size=p)
    founders = (rng.random((n_founders, p)) < freq).astype(np.uint8)
    d = np.maximum(rng.exponential(mean_dist_cm, size=p), 1e-9)
    d[0] = 0.0
    cm = 1.0 + np.cumsum(d)
    bp = 1_000_000 + np.cumsum(np.maximum(1, np.rint(d * 1e6).astype(np.int64)))
    groups = make_groups(p, mean_group, rng)
    b = np.exp(-rho * d)
    b[0] = 0.0                                        # knockoffs.cpp:1513
    mu = np.full(p, float(lam))

    hrng = np.random.default_rng([seed, 7919 if hap_seed is None else hap_seed])
    psw = (1.0 - np.exp(-d)).astype(np.float32)
    row_bytes = (p + 7) // 8
    H = np.empty((n_haps, row_bytes), dtype=np.uint8)
    cols = np.arange(p, dtype=np.int32)
    for c0 in range(0, n_haps, chunk):
        m = min(chunk, n_haps - c0)
        sw = hrng.random((m, p), dtype=np.float32) < psw
        sw[:, 0] = True
        ids = hrng.integers(0, n_founders, size=(m, p), dtype=np.int32)
        pos = np.where(sw, cols, 0)
        np.maximum.accumulate(pos, axis=1, out=pos)
        cur = np.take_along_axis(ids, pos, axis=1)
        bits = founders[cur, cols[None, :]]
        bits ^= (hrng.random((m, p), dtype=np.float32) < flip).astype(np.uint8)
        H[c0:c0 + m] = np.packbits(bits, axis=1, bitorder="little")
    ref = make_references(n_haps, K, hrng, block=min(block, n_haps))
    return SynthProblem(H=H, p=p, K=K, ref=ref, b=b, mu=mu, groups=groups, cm=cm, bp=bp, seed=seed)


def make_problem_fast(n_haps: int, p: int, K: int, seed: int = 2020, mean_group: float = 1.0, rho: float = 1.0,
                      lam: float = 1e-3, n_founders: int = 2000, mean_dist_cm: float = 0.005, flip: float = 1e-3,
                      block: int = 5000, hap_seed: int | None = None, chunk: int = 8192) -> SynthProblem:
    """Same population model as make_problem, generated about two orders of magnitude faster for the biobank-shaped
    cases (C3 / C5 shards: 87 500 x 60 000 cells would take minutes cell by cell): founder switches are drawn per
    64-locus block of the packed rows instead of per locus (rows are gathered from the packed founder matrix), and
    the sporadic allele flips are drawn as a sparse list of positions.  Per-SNP quantities (map, HMM tables, groups)
    are those of make_problem with the same `seed`; the haplotype rows differ."""
    rng = np.random.default_rng(seed)
    freq = rng.uniform(0.05, 0.5, size=p)
    founders = (rng.random((n_founders, p)) < freq).astype(np.uint8)
    d = np.maximum(rng.exponential(mean_dist_cm, size=p), 1e-9)
    d[0] = 0.0
    cm = 1.0 + np.cumsum(d)
    bp = 1_000_000 + np.cumsum(np.maximum(1, np.rint(d * 1e6).astype(np.int64)))
    groups = make_groups(p, mean_group, rng)
    b = np.exp(-rho * d)
    b[0] = 0.0                                        # knockoffs.cpp:1513
    mu = np.full(p, float(lam))

    hrng = np.random.default_rng([seed, 7919 if hap_seed is None else hap_seed, 64])
    nblk = (p + 63) // 64
    row_bytes = (p + 7) // 8
    fpad = np.zeros((n_founders, nblk * 64), dtype=np.uint8)
    fpad[:, :p] = founders
    F64 = np.packbits(fpad, axis=1, bitorder="little").view(np.uint64).reshape(n_founders, nblk)
    dpad = np.zeros(nblk * 64)
    dpad[:p] = d
    psw = (1.0 - np.exp(-dpad.reshape(nblk, 64).sum(axis=1))).astype(np.float32)
    H = np.empty((n_haps, row_bytes), dtype=np.uint8)
    blk = np.arange(nblk, dtype=np.int32)
    for c0 in range(0, n_haps, chunk):
        m = min(chunk, n_haps - c0)
        sw = hrng.random((m, nblk), dtype=np.float32) < psw
        sw[:, 0] = True
        ids = hrng.integers(0, n_founders, size=(m, nblk), dtype=np.int32)
        pos = np.where(sw, blk, 0)
        np.maximum.accumulate(pos, axis=1, out=pos)
        cur = np.take_along_axis(ids, pos, axis=1)
        rows = F64[cur, blk[None, :]]
        nf = int(hrng.binomial(m * p, flip))
        if nf:
            fr = hrng.integers(0, m, size=nf)
            fc = hrng.integers(0, p, size=nf)
            np.bitwise_xor.at(rows, (fr, fc >> 6), np.uint64(1) << (fc & 63).astype(np.uint64))
        H[c0:c0 + m] = rows.view(np.uint8).reshape(m, nblk * 8)[:, :row_bytes]
    ref = make_references(n_haps, K, hrng, block=min(block, n_haps))
    return SynthProblem(H=H, p=p, K=K, ref=ref, b=b, mu=mu, groups=groups, cm=cm, bp=bp, seed=seed)


def add_ibd_pairs(prob: SynthProblem, n_pairs: int, seed: int = 5, mean_len_cm: float = 20.0,
                  min_len_bp: int = 100_000, max_segments: int = 3) -> SynthProblem:
    """Synthetic RaPID-style IBD: `n_pairs` disjoint pairs of haplotypes from
    different samples, 1..max_segments segments per pair, length ~ Exp(mean_len_cm) and at least
    `min_len_bp`; the alleles of the first member are copied onto the second on every segment so that
    H is truly identical there, and neither member keeps the other as a reference.  Segments of a
    pair are disjoint and do not touch, so IbdCluster::tidy (ibd.cpp:99-189) leaves them as they are;
    they are listed in the std::set<IbdSeg> order (ibd.cpp:482-: same index set -> ascending j_min).
    Fills prob.families = [(members, [(j_min, j_max, ids), ...]), ...] as knockoffgwas_b200.api takes them."""
    rng = np.random.default_rng([prob.seed, seed])
    N, p = prob.N, prob.p
    if 2 * n_pairs > N // 2:
        raise ValueError("too many pairs for this population")
    samples = rng.choice(N // 2, size=2 * n_pairs, replace=False)
    bits = None
    fams = []
    for q in range(n_pairs):
        i1 = int(2 * samples[2 * q] + rng.integers(2))
        i2 = int(2 * samples[2 * q + 1] + rng.integers(2))
        a, c = (i1, i2) if i1 < i2 else (i2, i1)
        nseg = int(rng.integers(1, max_segments + 1))
        segs, tries = [], 0
        while len(segs) < nseg and tries < 50:
            tries += 1
            j0 = int(rng.integers(1, p - 2))
            end_cm = prob.cm[j0] + rng.exponential(mean_len_cm)
            j1 = int(min(p - 2, np.searchsorted(prob.cm, end_cm)))
            while j1 < p - 2 and prob.bp[j1] - prob.bp[j0] < min_len_bp:
                j1 += 1
            if prob.bp[j1] - prob.bp[j0] < min_len_bp:
                continue
            if any(not (j1 + 2 < s0 or j0 > s1 + 2) for s0, s1 in segs):
                continue
            segs.append((j0, j1))
        segs.sort()
        ids = np.array([a, c], dtype=np.int32)
        for j0, j1 in segs:
            ra = np.unpackbits(prob.H[a], bitorder="little")
            rc = np.unpackbits(prob.H[c], bitorder="little")
            rc[j0:j1 + 1] = ra[j0:j1 + 1]
            prob.H[c] = np.packbits(rc, bitorder="little")
        for me, other in ((a, c), (c, a)):          # relatives are never references (kinship_computer.cpp:196-203)
            hit = np.flatnonzero(prob.ref[me] == other)
            if len(hit):
                cand = other
                while cand == other or cand == me or cand in prob.ref[me]:
                    cand = int(rng.integers(max(0, me - 2000), min(N, me + 2000)))
                prob.ref[me, hit[0]] = cand
                prob.ref[me].sort()
        fams.append((ids, [(j0, j1, ids) for j0, j1 in segs]))
    prob.families = fams
    return prob


def write_bgen_image(H: np.ndarray, p: int, compression: int = 0, level: int = 1) -> np.ndarray:
    """Phased BGEN v1.2 image (layout 2, 8-bit probabilities) of the packed haplotype matrix H [2n][ceil(p/8)]:
    what `qctool`/UK Biobank ship and BgenReader (bgen_reader.cpp) ingests.  compression 0 = none, 1 = zlib.
    Used by bench.py to time the ingest kernel on the benchmark's own population."""
    import struct
    import zlib
    bits = np.unpackbits(np.ascontiguousarray(H, dtype=np.uint8), axis=1, bitorder="little")[:, :p]
    nh = bits.shape[0]
    n = nh // 2
    vals = np.where(bits.T == 0, 255, 0).astype(np.uint8)          # [p][2n]: P(first allele) as 8 bits
    head = struct.pack("<IHBB", n, 2, 2, 2) + bytes([2]) * n + bytes([1, 8])
    parts = [struct.pack("<I", 20), struct.pack("<III", 20, p, n), b"bgen", struct.pack("<I", compression | (2 << 2))]
    for j in range(p):
        rsid = b"rs%d" % j
        ident = struct.pack("<H", 0) + struct.pack("<H", len(rsid)) + rsid + struct.pack("<H", 2) + b"22"
        ident += struct.pack("<IH", 1000 + j, 2) + struct.pack("<I", 1) + b"A" + struct.pack("<I", 1) + b"G"
        raw = head + vals[j].tobytes()
        if compression == 0:
            blk = struct.pack("<I", len(raw)) + raw
        else:
            z = zlib.compress(raw, level)
            blk = struct.pack("<II", len(z) + 4, len(raw)) + z
        parts.append(ident + blk)
    return np.frombuffer(b"".join(parts), dtype=np.uint8)


def add_ibd_families(prob: SynthProblem, n_families: int, size: int = 3, segments: int = 4, seed: int = 9,
                     min_len_bp: int = 100_000) -> SynthProblem:
    """Families of `size` haplotypes (different samples) sharing `segments` IBD segments, each among a random subset of
    at least two members: nodes with several neighbours, members joining and leaving at different loci.  Segments of
    a family neither overlap nor touch (two loci apart at least) and span >= `min_len_bp`, so they are already what
    IbdCluster::tidy (ibd.cpp:99-189) would return; order: fewer indices first, then index set, then j_min
    (IbdSeg::operator<, ibd.cpp:482-509).  Alleles of the subset's first member are copied onto the others."""
    rng = np.random.default_rng([prob.seed, seed, size])
    N, p = prob.N, prob.p
    taken = {int(x) for m, _ in prob.families for x in m}
    free_samples = [s for s in rng.permutation(N // 2) if 2 * s not in taken and 2 * s + 1 not in taken]
    if len(free_samples) < n_families * size:
        raise ValueError("not enough unrelated samples left")
    fams = list(prob.families)
    for f in range(n_families):
        mem = np.sort(np.array([2 * free_samples[f * size + t] + int(rng.integers(2)) for t in range(size)], dtype=np.int32))
        cuts = np.sort(rng.choice(np.arange(2, p - 2), size=2 * segments, replace=False))
        segs = []
        for q in range(segments):
            j0, j1 = int(cuts[2 * q]) + 1, int(cuts[2 * q + 1]) - 1
            if j1 - j0 < 3 or prob.bp[j1] - prob.bp[j0] < min_len_bp:
                continue
            ids = np.sort(rng.choice(mem, size=int(rng.integers(2, size + 1)), replace=False)).astype(np.int32)
            src = np.unpackbits(prob.H[ids[0]], bitorder="little")
            for other in ids[1:]:
                row = np.unpackbits(prob.H[other], bitorder="little")
                row[j0:j1 + 1] = src[j0:j1 + 1]
                prob.H[other] = np.packbits(row, bitorder="little")
            segs.append((j0, j1, ids))
        if not segs:
            continue
        linked = set()
        for _, _, ids in segs:
            linked |= {int(x) for x in ids}
        mem = np.array(sorted(linked), dtype=np.int32)     # members without any segment are not part of the family
        segs.sort(key=lambda sg: (len(sg[2]), tuple(int(x) for x in sg[2]), sg[0]))
        for me in mem:                                      # relatives are never references
            for other in mem:
                hit = np.flatnonzero(prob.ref[me] == other)
                if me != other and len(hit):
                    cand = int(other)
                    while cand in mem or cand in prob.ref[me]:
                        cand = int(rng.integers(max(0, me - 2000), min(N, me + 2000)))
                    prob.ref[me, hit[0]] = cand
                    prob.ref[me].sort()
        fams.append((mem, segs))
    prob.families = fams
    return prob


def add_big_family(prob: SynthProblem, size: int, segments: int, max_subset: int = 12, seed: int = 13,
                   min_len_bp: int = 100_000) -> SynthProblem:
    """One IBD family of `size` haplotypes (different samples) with `segments` disjoint segments.  Every member has a
    home segment (so that all of them are linked), segments then take random extra members up to `max_subset`: nodes
    with many neighbours, members joining and leaving at every boundary -- the shapes union-find over RaPID pairs
    produces on biobank data (metadata.cpp:242-293).  Segments neither overlap nor touch, are at least `min_len_bp`
    long and are listed in IbdSeg order (fewer indices first, then index set, then j_min; ibd.cpp:482-509)."""
    rng = np.random.default_rng([prob.seed, seed, size, segments])
    N, p = prob.N, prob.p
    taken = {int(x) for m, _ in prob.families for x in m}
    free_samples = [s for s in rng.permutation(N // 2) if 2 * s not in taken and 2 * s + 1 not in taken]
    if len(free_samples) < size:
        raise ValueError("not enough unrelated samples left")
    mem = np.sort(np.array([2 * free_samples[t] + int(rng.integers(2)) for t in range(size)], dtype=np.int32))
    home = np.array_split(rng.permutation(mem), segments)
    for _ in range(200):
        cuts = np.sort(rng.choice(np.arange(2, p - 2), size=2 * segments, replace=False))
        ok = all(cuts[2 * q + 1] - 1 - (cuts[2 * q] + 1) >= 3 and prob.bp[cuts[2 * q + 1] - 1] - prob.bp[cuts[2 * q] + 1] >= min_len_bp
                 for q in range(segments))
        if ok:
            break
    else:
        raise ValueError("could not place the segments: chromosome too short for min_len_bp")
    segs = []
    for q in range(segments):
        j0, j1 = int(cuts[2 * q]) + 1, int(cuts[2 * q + 1]) - 1
        ids = set(int(x) for x in home[q])
        extra = int(rng.integers(0, max(1, max_subset - len(ids) + 1)))
        others = [int(x) for x in rng.permutation(mem) if int(x) not in ids][:extra]
        ids |= set(others)
        if len(ids) < 2:
            ids.add(int(next(x for x in mem if int(x) not in ids)))
        ids = np.array(sorted(ids), dtype=np.int32)
        src = np.unpackbits(prob.H[ids[0]], bitorder="little")
        for other in ids[1:]:
            row = np.unpackbits(prob.H[other], bitorder="little")
            row[j0:j1 + 1] = src[j0:j1 + 1]
            prob.H[other] = np.packbits(row, bitorder="little")
        segs.append((j0, j1, ids))
    segs.sort(key=lambda sg: (len(sg[2]), tuple(int(x) for x in sg[2]), sg[0]))
    memset = set(int(x) for x in mem)
    for me in mem:                                      # relatives are never references
        bad = [t for t in range(prob.ref.shape[1]) if int(prob.ref[me, t]) in memset]
        for t in bad:
            cand = int(me)
            while cand in memset or cand in prob.ref[me]:
                cand = int(rng.integers(0, N))
            prob.ref[me, t] = cand
        if bad:
            prob.ref[me].sort()
    prob.families = list(prob.families) + [(mem, segs)]
    return prob
