"""ctypes binding of the CPU oracle (oracle/libkgw_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(knockoffgwas_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

STREAM_Z, STREAM_ZK, STREAM_HK, STREAM_FAMILY = 0, 1, 2, 3
RNG_INJECTED, RNG_TAUS88, RNG_PHILOX = 0, 1, 2


class _Taus(C.Structure):
    _fields_ = [("v1", C.c_uint32), ("v2", C.c_uint32), ("v3", C.c_uint32)]


class _Problem(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("p", C.c_int32), ("K", C.c_int32), ("W", C.c_int32),
        ("row_bytes", C.c_int64),
        ("H", C.c_void_p), ("ref_local", C.c_void_p), ("ref_global", C.c_void_p),
        ("b", C.c_void_p), ("mu", C.c_void_p), ("groups", C.c_void_p),
        ("win_start", C.c_void_p), ("win_end", C.c_void_p),
    ]


class _Rng(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("injected", C.c_void_p), ("taus", _Taus),
        ("seed", C.c_uint64), ("n_draws", C.c_int64), ("record", C.c_void_p),
        ("n_windows", C.c_int32), ("reserved", C.c_int32),
    ]


class _Family(C.Structure):
    _fields_ = [("n_members", C.c_int32), ("members", C.c_void_p), ("n_segs", C.c_int32),
                ("seg_jmin", C.c_void_p), ("seg_jmax", C.c_void_p), ("seg_off", C.c_void_p), ("seg_ids", C.c_void_p)]


def build(force: bool = False) -> Path:
    so = _HERE / "libkgw_oracle.so"
    srcs = [_HERE / n for n in ("kgw_oracle.c", "kgw_oracle_family.c", "kgw_rng.c", "kgw_oracle.h", "kgw_rng.h")]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "libkgw_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.kgw_oracle_unrelated.restype = C.c_int
        _LIB.kgw_oracle_family_run.restype = C.c_int
        _LIB.kgw_oracle_weighted_choice.restype = C.c_int
        _LIB.kgw_oracle_weighted_choice.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _LIB.kgw_taus88_seed.argtypes = [C.POINTER(_Taus), C.c_uint32]
        _LIB.kgw_taus88_next.argtypes = [C.POINTER(_Taus)]
        _LIB.kgw_taus88_next.restype = C.c_uint32
        _LIB.kgw_oracle_philox_uniform.restype = C.c_double
        _LIB.kgw_oracle_philox_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        _LIB.kgw_oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


@dataclass
class Problem:
    """Everything the unchanged host code hands to the hot path."""
    H: np.ndarray                 # uint8 [N][row_bytes], LSB-first
    p: int
    K: int
    ref_local: np.ndarray         # int32 [N][W][K]
    ref_global: np.ndarray        # int32 [N][K]
    b: np.ndarray                 # float64 [p]
    mu: np.ndarray                # float64 [p]
    groups: np.ndarray            # int32 [p]
    win_start: np.ndarray = None  # int32 [W]
    win_end: np.ndarray = None    # int32 [W]
    unrelated: np.ndarray = None  # int32 sorted haplotype ids
    families: list = field(default_factory=list)   # list of int32 arrays (haplotype ids)
    ibd_segments: list = field(default_factory=list)  # list of (hap, j_min, j_max, bp_min, bp_max, ids)
    # families in the reference's single-thread processing order with their TIDIED segments
    # (IbdCluster::tidy, ibd.cpp:99-189, run by the unchanged host code): list of
    # (members int32[], [(j_min, j_max, bp_min, bp_max, ids int32[]), ...])
    fam_clusters: list = field(default_factory=list)
    bp: np.ndarray = None
    seed: int = 2020
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.H = np.ascontiguousarray(self.H, dtype=np.uint8)
        self.N = int(self.H.shape[0])
        self.row_bytes = int(self.H.shape[1])
        if self.win_start is None:
            self.win_start = np.array([0], dtype=np.int32)
            self.win_end = np.array([self.p], dtype=np.int32)
        self.W = int(len(self.win_start))
        self.ref_local = np.ascontiguousarray(self.ref_local, dtype=np.int32).reshape(self.N, self.W, self.K)
        self.ref_global = np.ascontiguousarray(self.ref_global, dtype=np.int32).reshape(self.N, self.K)
        self.b = np.ascontiguousarray(self.b, dtype=np.float64)
        self.mu = np.ascontiguousarray(self.mu, dtype=np.float64)
        self.groups = np.ascontiguousarray(self.groups, dtype=np.int32)
        self.win_start = np.ascontiguousarray(self.win_start, dtype=np.int32)
        self.win_end = np.ascontiguousarray(self.win_end, dtype=np.int32)
        if self.unrelated is None:
            self.unrelated = np.arange(self.N, dtype=np.int32)
        self.unrelated = np.ascontiguousarray(self.unrelated, dtype=np.int32)

    def _c(self) -> _Problem:
        P = _Problem()
        P.N, P.p, P.K, P.W = self.N, self.p, self.K, self.W
        P.row_bytes = self.row_bytes
        P.H = self.H.ctypes.data
        P.ref_local = self.ref_local.ctypes.data
        P.ref_global = self.ref_global.ctypes.data
        P.b = self.b.ctypes.data
        P.mu = self.mu.ctypes.data
        P.groups = self.groups.ctypes.data
        P.win_start = self.win_start.ctypes.data
        P.win_end = self.win_end.ctypes.data
        return P


def parse_fam_clusters(raw) -> list:
    """Inverse of the fam_clusters.i32 record written by oracle/ref_shims/wrap_probe.cpp."""
    raw = np.asarray(raw, dtype=np.int32)
    if len(raw) == 0:
        return []
    out, t = [], 1
    for _ in range(int(raw[0])):
        n = int(raw[t]); members = raw[t + 1:t + 1 + n].copy(); t += 1 + n
        ns = int(raw[t]); t += 1
        segs = []
        for _ in range(ns):
            jmin, jmax, bpmin, bpmax, m = (int(x) for x in raw[t:t + 5])
            segs.append((jmin, jmax, bpmin, bpmax, raw[t + 5:t + 5 + m].copy()))
            t += 5 + m
        out.append((members, segs))
    assert t == len(raw)
    return out


def load_probe_dump(d) -> tuple[Problem, np.ndarray, np.ndarray | None]:
    """Load a directory written by oracle/_ref/snpknock2_probe (wrap_probe.cpp).
    Returns (problem, Hk_reference uint8 [N][row_bytes], trace int32 [n][2] or None)."""
    d = Path(d)
    meta = {}
    for line in (d / "meta.txt").read_text().splitlines():
        k, v = line.split(None, 1)
        meta[k] = v
    N, p, K, W = (int(meta[k]) for k in ("N", "p", "K", "W"))
    rb = (p + 7) // 8
    rd = lambda name, dt: np.fromfile(d / name, dtype=dt)
    fam_off = rd("fam_offsets.i32", np.int32)
    fam_mem = rd("fam_members.i32", np.int32)
    families = [fam_mem[fam_off[f]:fam_off[f + 1]].copy() for f in range(len(fam_off) - 1)]
    raw = rd("ibd_segments.i32", np.int32)
    segs, t = [], 0
    while t < len(raw):
        hap, jmin, jmax, bpmin, bpmax, n = (int(x) for x in raw[t:t + 6])
        segs.append((hap, jmin, jmax, bpmin, bpmax, raw[t + 6:t + 6 + n].copy()))
        t += 6 + n
    clusters = parse_fam_clusters(rd("fam_clusters.i32", np.int32)) if (d / "fam_clusters.i32").exists() else []
    prob = Problem(
        H=rd("H.u8", np.uint8).reshape(N, rb), p=p, K=K,
        ref_local=rd("ref_local.i32", np.int32).reshape(N, W, K),
        ref_global=rd("ref_global.i32", np.int32).reshape(N, K),
        b=rd("b.f64", np.float64), mu=rd("mu.f64", np.float64), groups=rd("groups.i32", np.int32),
        win_start=rd("win_start.i32", np.int32), win_end=rd("win_end.i32", np.int32),
        unrelated=rd("unrelated.i32", np.int32), families=families, ibd_segments=segs,
        bp=rd("bp.i32", np.int32), seed=int(meta["seed"]), meta=meta, fam_clusters=clusters)
    Hk = rd("Hk.u8", np.uint8).reshape(N, rb)
    trace = None
    if (d / "trace.i32").exists() and meta.get("traced") == "1":
        trace = rd("trace.i32", np.int32).reshape(-1, 2)
    return prob, Hk, trace


def taus88_stream(seed: int | None, n: int) -> np.ndarray:
    """n raw 32-bit outputs of boost taus88; seed None = default-constructed (341)."""
    g = _Taus()
    L = lib()
    L.kgw_taus88_seed(C.byref(g), 341 if seed is None else int(seed) & 0xFFFFFFFF)
    out = np.empty(n, dtype=np.uint32)
    for i in range(n):
        out[i] = L.kgw_taus88_next(C.byref(g))
    return out


def philox_uniform(seed: int, hap: int, stream: int, locus: int) -> float:
    return lib().kgw_oracle_philox_uniform(int(seed), int(hap), int(stream), int(locus))


def philox_uniforms(seed: int, haps: np.ndarray, p: int) -> np.ndarray:
    """U[len(haps)][3][p] of the keyed generator (numpy restatement, vectorised)."""
    haps = np.asarray(haps, dtype=np.uint32)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    nblk = (p + 3) // 4
    c0 = np.broadcast_to(np.arange(nblk, dtype=np.uint32)[None, None, :], (len(haps), 3, nblk)).copy()
    c1 = np.broadcast_to(haps[:, None, None], c0.shape).copy()
    c2 = np.broadcast_to(np.arange(3, dtype=np.uint32)[None, :, None], c0.shape).copy()
    c3 = np.zeros_like(c0)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = p0.astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    words = np.stack([c0, c1, c2, c3], axis=-1).reshape(len(haps), 3, nblk * 4)[:, :, :p]
    return words.astype(np.float64) / 4294967296.0


@dataclass
class Result:
    z: np.ndarray
    zk: np.ndarray
    hk: np.ndarray
    beta: np.ndarray | None
    n_draws: int
    taus_state: tuple | None = None


def run_unrelated(prob: Problem, haps=None, *, mode="philox", seed=None, injected=None,
                  taus_seed=None, taus_state=None, want_beta=False, record=None) -> Result:
    """Run the restated unrelated-haplotype sampler for `haps` (default: prob.unrelated).

    mode: "philox" (key = seed), "injected" (U[N][3][p], indexed by absolute haplotype id),
          "taus88" (sequential replay of the reference's stream; taus_seed None = boost default)."""
    L = lib()
    haps = prob.unrelated if haps is None else np.ascontiguousarray(haps, dtype=np.int32)
    n = len(haps)
    R, keep = make_rng(prob, mode=mode, seed=seed, injected=injected, taus_seed=taus_seed, taus_state=taus_state,
                       record=record)   # `record`: float64 [N][3][p] or [N][W][3][p], receives every uniform handed out
    z = np.zeros((n, prob.p), dtype=np.int32)
    zk = np.zeros((n, prob.p), dtype=np.int32)
    hk = np.zeros((n, prob.row_bytes), dtype=np.uint8)
    beta = np.zeros((n, prob.p, prob.K), dtype=np.float64) if want_beta else None
    P = prob._c()
    rc = L.kgw_oracle_unrelated(C.byref(P), C.byref(R), C.c_void_p(haps.ctypes.data), C.c_int32(n),
                                C.c_void_p(z.ctypes.data), C.c_void_p(zk.ctypes.data),
                                C.c_void_p(hk.ctypes.data),
                                C.c_void_p(beta.ctypes.data if want_beta else None))
    if rc != 0:
        raise RuntimeError(f"kgw_oracle_unrelated failed with {rc}")
    return Result(z, zk, hk, beta, int(R.n_draws), (R.taus.v1, R.taus.v2, R.taus.v3))


def make_rng(prob: Problem, mode="philox", seed=None, injected=None, taus_seed=None, taus_state=None, record=None):
    """Uniform source shared by run_unrelated / run_family.  Returns (_Rng, keep-alive object).
    record: optional float64 [N][3][p] array that receives every uniform at its (hap, stream, locus) key."""
    L = lib()
    R = _Rng()
    keep = None
    if record is not None:
        assert record.dtype == np.float64 and record.flags.c_contiguous
        if record.shape == (prob.N, prob.W, 3, prob.p):
            R.n_windows = prob.W
        else:
            assert record.shape == (prob.N, 3, prob.p)
        R.record = record.ctypes.data
    if mode == "injected":
        keep = np.ascontiguousarray(injected, dtype=np.float64)
        if keep.shape == (prob.N, prob.W, 3, prob.p):
            R.n_windows = prob.W
        else:
            assert keep.shape == (prob.N, 3, prob.p)
        R.mode, R.injected = RNG_INJECTED, keep.ctypes.data
    elif mode == "taus88":
        R.mode = RNG_TAUS88
        if taus_state is not None:
            R.taus.v1, R.taus.v2, R.taus.v3 = taus_state
        else:
            L.kgw_taus88_seed(C.byref(R.taus), 341 if taus_seed is None else int(taus_seed) & 0xFFFFFFFF)
    elif mode == "philox":
        R.mode, R.seed = RNG_PHILOX, int(prob.seed if seed is None else seed)
    else:
        raise ValueError(mode)
    return R, keep


@dataclass
class FamilyResult:
    members: np.ndarray
    z: np.ndarray
    zk: np.ndarray
    hk: np.ndarray
    bp_sweeps: int


def run_family(prob: Problem, members, segments, rng: "_Rng") -> FamilyResult:
    """Restated generate_related_cluster for one family.  `segments`: tidied
    [(j_min, j_max, bp_min, bp_max, ids), ...] in the reference's std::set order; `rng` from make_rng
    (sequential modes carry their state across calls)."""
    L = lib()
    members = np.ascontiguousarray(members, dtype=np.int32)
    n = len(members)
    jmin = np.array([s[0] for s in segments], dtype=np.int32)
    jmax = np.array([s[1] for s in segments], dtype=np.int32)
    ids = [np.asarray(s[4], dtype=np.int32) for s in segments]
    off = np.concatenate([[0], np.cumsum([len(x) for x in ids])]).astype(np.int32)
    flat = np.concatenate(ids).astype(np.int32) if ids else np.zeros(0, np.int32)
    F = _Family()
    F.n_members, F.members, F.n_segs = n, members.ctypes.data, len(segments)
    F.seg_jmin, F.seg_jmax, F.seg_off, F.seg_ids = jmin.ctypes.data, jmax.ctypes.data, off.ctypes.data, flat.ctypes.data
    z = np.zeros((n, prob.p), dtype=np.int32)
    zk = np.zeros((n, prob.p), dtype=np.int32)
    hk = np.zeros((n, prob.row_bytes), dtype=np.uint8)
    sweeps = C.c_int64(0)
    P = prob._c()
    rc = L.kgw_oracle_family_run(C.byref(P), C.byref(rng), C.byref(F), C.c_void_p(z.ctypes.data),
                                 C.c_void_p(zk.ctypes.data), C.c_void_p(hk.ctypes.data), C.byref(sweeps))
    if rc != 0:
        raise RuntimeError(f"kgw_oracle_family_run failed with {rc}")
    return FamilyResult(members, z, zk, hk, int(sweeps.value))


def unpack_bits(rows: np.ndarray, p: int) -> np.ndarray:
    return np.unpackbits(rows, axis=1, bitorder="little")[:, :p]


def pack_bits(bits: np.ndarray) -> np.ndarray:
    return np.packbits(bits.astype(np.uint8), axis=1, bitorder="little")


def bed_body(H, Hk, p: int, swap=None) -> np.ndarray:
    """Bytes of `<out>.bed` after the magic, restating writeBED / writeBED_variant
    (snpknock2/src/plink_interface.cpp:10-35, 67-103): per SNP j the columns (H, Hk), or (Hk, H) where
    swap[j]; with H None the p columns of Hk (write_binary(basename, Hk), :240-246).  A column holds
    ceil(N/2/4) bytes; sample s = haplotypes 2s, 2s+1 -> bits 2(s%4), 2(s%4)+1 of byte s/4:
    both alleles set -> 11, exactly one -> 10 (the HIGH bit), none -> 00 (:20-28)."""
    def columns(M):
        bits = unpack_bits(np.asarray(M, dtype=np.uint8), p).astype(np.uint8)   # [N][p]
        N = bits.shape[0]
        assert N % 2 == 0
        b0, b1 = bits[0::2], bits[1::2]                                         # [N/2][p]
        code = (b0 & b1) | ((b0 | b1) << 1)                                     # 2-bit genotype codes
        ns = N // 2
        nb = (ns + 3) // 4
        pad = np.zeros((4 * nb, p), dtype=np.uint8)
        pad[:ns] = code
        q = pad.reshape(nb, 4, p)
        byte = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)       # [nb][p]
        return np.ascontiguousarray(byte.T)                                     # [p][nb]
    ck = columns(Hk)
    if H is None:
        return ck
    ch = columns(H)
    sw = np.zeros(p, dtype=bool) if swap is None else np.asarray(swap).astype(bool)
    out = np.empty((2 * p, ck.shape[1]), dtype=np.uint8)
    out[0::2] = np.where(sw[:, None], ck, ch)
    out[1::2] = np.where(sw[:, None], ch, ck)
    return out


def thin_haplotypes(H, p: int, compression: int = 10) -> tuple[np.ndarray, int]:
    """The thinned copy of the chromosome that KinshipComputer measures distances on (Dataset(metadata,
    num_threads, compression), kinship_computer.cpp:19; Metadata::thin keeps the SNPs with j % compression == 0,
    metadata.cpp:563-580; main.cpp:101 fixes compression = 10): packed rows of ceil(p / compression) bits."""
    bits = unpack_bits(np.asarray(H, dtype=np.uint8), p)[:, ::compression]
    return pack_bits(bits), bits.shape[1]


def find_neighbors(H, p: int, K: int, clusters, family_of, compression: int = 10) -> np.ndarray:
    """KinshipComputer::findNeighbors for every haplotype with one window (kinship_computer.cpp:174-233): within
    the haplotype's kinship cluster, the K members with the smallest Hamming distance on the thinned chromosome
    (chaplotype::hamming over whole bytes, chaplotype.cpp:62-66), members of the same IBD family pushed to
    distance 1e6 (:197-199), ties kept in cluster order (multimap insertion order and the strict
    `dij < top_key` test, :207-218).  Returns int32 [N][K] in (distance, cluster position) order, -1 where a
    cluster has fewer than K+1 members.  family_of[i] = label shared by the haplotypes of
    metadata.related_families[i], negative for singletons.
    (With --windows > 0 the reference applies the full-resolution window bounds to the thinned rows,
    distance_hap_window :78-104, and reads past the end of the byte vectors; that mode has no defined result
    and is not restated.)"""
    Ht, pt = thin_haplotypes(H, p, compression)
    N = Ht.shape[0]
    bits = unpack_bits(Ht, pt).astype(np.int32)
    family_of = np.asarray(family_of)
    ref = np.full((N, K), -1, dtype=np.int32)
    for members in clusters:
        members = np.asarray(members, dtype=np.int64)
        n = len(members)
        fam = family_of[members]
        related = (fam[:, None] == fam[None, :]) & (fam[:, None] >= 0)
        X = bits[members]
        D = (X @ (1 - X).T + (1 - X) @ X.T).astype(np.int64)
        D[related] = 1000000
        D[np.arange(n), np.arange(n)] = np.iinfo(np.int64).max     # `if(i==j) continue`
        order = np.argsort(D, axis=1, kind="stable")[:, :min(K, n - 1)]
        ref[members, :order.shape[1]] = members[order]
    return ref


def combine_references_families(ref_in: np.ndarray, families) -> np.ndarray:
    """KinshipComputer::combine_references_families (kinship_computer.cpp:765-835): every list sorted ascending;
    the members of an IBD family share the intersection of their lists, topped up round-robin with each
    member's next-closest reference (in the ORIGINAL distance order) until it holds K.
    families[i] = metadata.related_families[i] (list of haplotype ids, itself included)."""
    if ref_in.ndim == 2:
        return combine_references_families(ref_in[:, None, :], families)[:, 0, :]
    N, W, K = ref_in.shape
    ref = np.sort(ref_in, axis=2)
    for w in range(W):
        done = np.zeros(N, dtype=bool)
        for i1 in range(N):
            if done[i1]:
                continue
            fam = [int(x) for x in families[i1]]
            done[i1] = True
            for i2 in fam:
                done[i2] = True
            if len(fam) <= 1:
                continue
            inter = list(ref[i1, w])
            for i2 in fam:
                if i2 == i1:
                    continue
                s2 = set(int(x) for x in ref[i2, w])
                inter = [x for x in inter if int(x) in s2]
            inter = sorted(set(int(x) for x in inter))
            i = k = 0
            while len(inter) < K:
                cand = int(ref_in[fam[i], w, k])
                if cand not in inter:
                    inter.append(cand)
                    inter.sort()
                if i + 1 < len(fam):
                    i += 1
                else:
                    i, k = 0, k + 1
            for i2 in fam:
                ref[i2, w] = inter
    return ref


def family_labels(families) -> np.ndarray:
    """haplotype -> label shared within metadata.related_families[i] (the smallest member), -1 for singletons."""
    return np.array([min(int(x) for x in f) if len(f) > 1 else -1 for f in families], dtype=np.int32)



# ---------------------------------------------------------------------------------------------------------
# Phased BGEN -> haplotype matrix.  Restates BgenReader::read_worker / prob_to_hap
# (snpknock2/src/bgen_reader.cpp:48-63, 139-261) on top of the BGEN v1.2 container (layout 2): the genfile::bgen
# library the reference links is vendored there (snpknock2/include/bgen); its read_probs() hands prob_to_hap, per
# sample, [P(h1 = first allele), P(h1 = second), P(h2 = first), P(h2 = second)] with P = stored / (2^B - 1).
def _zstd_decompress(buf: bytes, raw_len: int) -> bytes:
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
    out = C.create_string_buffer(raw_len)
    got = z.ZSTD_decompress(out, raw_len, buf, len(buf))
    if got != raw_len:
        raise ValueError("zstd block did not inflate to its stated length")
    return out.raw


def bgen_variants(image: bytes):
    """[(rsid, raw probability block bytes)] in file order."""
    import struct
    import zlib
    d = bytes(image)
    offset, hlen, M, N = struct.unpack_from("<IIII", d, 0)
    (flags,) = struct.unpack_from("<I", d, 4 + hlen - 4)
    comp, layout = flags & 3, (flags >> 2) & 15
    if layout != 2:
        raise ValueError("layout 2 only")
    out, pos = [], 4 + offset
    for _ in range(M):
        (l,) = struct.unpack_from("<H", d, pos); pos += 2 + l
        (l,) = struct.unpack_from("<H", d, pos); rsid = d[pos + 2:pos + 2 + l].decode(); pos += 2 + l
        (l,) = struct.unpack_from("<H", d, pos); pos += 2 + l
        (_, nall) = struct.unpack_from("<IH", d, pos); pos += 6
        for _a in range(nall):
            (l,) = struct.unpack_from("<I", d, pos); pos += 4 + l
        (c,) = struct.unpack_from("<I", d, pos); pos += 4
        if comp == 0:
            raw = d[pos:pos + c]
        else:
            (dl,) = struct.unpack_from("<I", d, pos)
            blk = d[pos + 4:pos + c]
            raw = zlib.decompress(blk) if comp == 1 else _zstd_decompress(blk, dl)
            if len(raw) != dl:
                raise ValueError("block length mismatch")
        pos += c
        out.append((rsid, raw))
    return out, N


def bgen_read(image, sample_idx, variant_of_col) -> np.ndarray:
    """uint8 [2 len(sample_idx)][ceil(p/8)]: H[2i + a][j] = prob_to_hap of haplotype a of file sample
    sample_idx[i] at file variant variant_of_col[j]; raises ValueError("Incorrect input") where the reference
    throws (probability not exactly 0 / 1, missing sample)."""
    import struct
    variants, N = bgen_variants(image)
    sample_idx = np.asarray(sample_idx, dtype=np.int64)
    p = len(variant_of_col)
    bits = np.zeros((2 * len(sample_idx), p), dtype=np.uint8)
    for j, v in enumerate(variant_of_col):
        raw = variants[int(v)][1]
        n, nall, pmin, pmax = struct.unpack_from("<IHBB", raw, 0)
        ploidy = np.frombuffer(raw, dtype=np.uint8, count=n, offset=8)
        phased, B = raw[8 + n], raw[9 + n]
        if n != N or nall != 2 or pmin != 2 or pmax != 2 or phased != 1 or not (1 <= B <= 32):
            raise ValueError("not phased diploid biallelic data")
        stream = np.unpackbits(np.frombuffer(raw, dtype=np.uint8, offset=10 + n), bitorder="little")
        vals = stream[:2 * n * B].reshape(2 * n, B).astype(np.uint64)
        vals = (vals << np.arange(B, dtype=np.uint64)).sum(axis=1)          # little-endian bit stream, B bits each
        ones = np.uint64((1 << B) - 1)
        for a in (0, 1):
            q = vals[2 * sample_idx + a]
            if np.any(ploidy[sample_idx] != 2) or np.any((q != 0) & (q != ones)):
                raise ValueError("Incorrect input")                            # bgen_reader.cpp:54-60
            bits[2 * np.arange(len(sample_idx)) + a, j] = (q == 0)            # P(first allele) == 0 -> second allele -> 1
    return pack_bits(bits)


def bgen_write(bits: np.ndarray, rsids, B: int = 8, compression: int = 1, missing=None, bad_prob=None, zlevel: int = -1) -> bytes:
    """Synthetic phased BGEN v1.2 image (layout 2) for tests: bits uint8 [2 n][p]; compression 0 none / 1 zlib;
    `missing` = (sample, variant) flagged missing, `bad_prob` = (haplotype, variant) given probability 1/2."""
    import struct
    import zlib
    nh, p = bits.shape
    n = nh // 2
    body = b""
    for j in range(p):
        vals = np.where(bits[:, j] == 0, (1 << B) - 1, 0).astype(np.uint64)
        if bad_prob is not None and bad_prob[1] == j:
            vals[bad_prob[0]] = (1 << B) // 2
        ploidy = np.full(n, 2, dtype=np.uint8)
        if missing is not None and missing[1] == j:
            ploidy[missing[0]] = 0x82
            vals[2 * missing[0]:2 * missing[0] + 2] = 0
        st = ((vals[:, None] >> np.arange(B, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8).reshape(-1)
        raw = struct.pack("<IHBB", n, 2, 2, 2) + ploidy.tobytes() + bytes([1, B]) + np.packbits(st, bitorder="little").tobytes()
        ident = b"".join(struct.pack("<H", len(x)) + x for x in (b"", rsids[j].encode(), b"22"))
        ident += struct.pack("<IH", 1000 + j, 2) + struct.pack("<I", 1) + b"A" + struct.pack("<I", 1) + b"G"
        if compression == 0:
            blk = struct.pack("<I", len(raw)) + raw
        else:
            z = zlib.compress(raw, zlevel)   # level 0 writes stored blocks, small inputs fixed-Huffman ones
            blk = struct.pack("<II", len(z) + 4, len(raw)) + z
        body += ident + blk
    flags = compression | (2 << 2)
    header = struct.pack("<III", 20, p, n) + b"bgen" + struct.pack("<I", flags)
    return struct.pack("<I", 20) + header + body


def family_audit(enable: bool) -> None:
    """Switch the audit of the CUDA kernel's BP work-skipping rule on / off (and zero its counters)."""
    lib().kgw_oracle_family_audit(C.c_int(1 if enable else 0))


def family_audit_stats() -> dict:
    out = (C.c_int64 * 3)()
    lib().kgw_oracle_family_audit_get(out)
    return {"messages": int(out[0]), "skippable": int(out[1]), "violations": int(out[2])}


# ---------------------------------------------------------------------------------------------------------
# `.haps` text -> haplotype matrix.  Restates the data loop of HapsReader::read (snpknock2/src/haps_reader.cpp:93-124)
# with sutils::tokenize (utils.cpp:234-246).
def haps_tokenize(line: bytes) -> list:
    import re
    toks = [t for t in re.split(rb"[ \t]+", line) if t != b""]       # maximal runs of non-blank characters
    if toks and toks[-1].endswith(b"\r"):                               # utils.cpp:243-244, last token only
        toks[-1] = toks[-1][:-1]
    return toks


def haps_read(text, sample_idx, row_of_col) -> np.ndarray:
    """uint8 [2 len(sample_idx)][ceil(p/8)]: H[2i + a][j] = (token 2 sample_idx[i] + a of file row row_of_col[j] == "1");
    raises IndexError where the reference would index past the end of `tokens`."""
    lines = bytes(text).split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                                                     # std::getline yields no row after the last '\n'
    sample_idx = np.asarray(sample_idx, dtype=np.int64)
    bits = np.zeros((2 * len(sample_idx), len(row_of_col)), dtype=np.uint8)
    for j, r in enumerate(row_of_col):
        toks = haps_tokenize(lines[int(r)])
        for i, s in enumerate(sample_idx):
            bits[2 * i, j] = toks[2 * s] == b"1"
            bits[2 * i + 1, j] = toks[2 * s + 1] == b"1"
    return pack_bits(bits)


def run_em(prob: Problem, mu_init=None):
    """Restated Knockoffs::EM (--estimate-hmm): returns (mu after EM, steps run, per-step deltas, converged)."""
    L = lib()
    L.kgw_oracle_em.restype = C.c_int
    p = prob.p
    mu0 = np.ascontiguousarray(np.full(p, 1e-5) if mu_init is None else mu_init, dtype=np.float64)
    out = np.zeros(p, dtype=np.float64)
    deltas = np.zeros(20, dtype=np.float64)
    it = C.c_int32(0)
    P = prob._c()
    rc = L.kgw_oracle_em(C.byref(P), C.c_void_p(mu0.ctypes.data), C.c_void_p(out.ctypes.data), C.byref(it),
                         C.c_void_p(deltas.ctypes.data))
    if rc < 0:
        raise RuntimeError(f"kgw_oracle_em failed with {rc}")
    n = int(it.value)
    return out, n, deltas[:max(n, 1)].copy(), rc == 0


# ---- kinship k-means data passes (test infrastructure; kinship_computer.cpp:239-250, 314-336, 368-405) -------------
def kinship_distances(H, p, members, mu0, mu1, use=None):
    """update_assignments' distances restated: for member i (a haplotype index) the squared distance of BOTH haplotypes
    of its sample to each centroid, one accumulator per haplotype walked in locus order (distance(), :239-250), then
    distance(i0) + distance(i1) (:327-328).  Returns (d0, d1)."""
    bits = np.unpackbits(np.ascontiguousarray(H, dtype=np.uint8), axis=1, bitorder="little")[:, :p].astype(np.float64)
    members = np.asarray(members, dtype=np.int64)
    use = np.ones(p, dtype=bool) if use is None else np.asarray(use).astype(bool)
    out = []
    for mu in (np.asarray(mu0, dtype=np.float64), np.asarray(mu1, dtype=np.float64)):
        tot = []
        for h in (0, 1):
            rows = bits[2 * (members // 2) + h]
            acc = np.zeros(len(members))
            for j in range(p):                      # sequential in j: the order of the reference's single accumulator
                if use[j]:
                    d = rows[:, j] - mu[j]
                    acc = acc + d * d
            tot.append(acc)
        out.append(tot[0] + tot[1])
    return out[0], out[1]


def kinship_centroids(H, p, members, assign):
    """update_centroids restated (:368-405): sums of 0.0 / 1.0 over the members of each side, divided by the side's size."""
    bits = np.unpackbits(np.ascontiguousarray(H, dtype=np.uint8), axis=1, bitorder="little")[:, :p].astype(np.float64)
    members = np.asarray(members, dtype=np.int64)
    assign = np.asarray(assign)
    mus = []
    for side in (0, 1):
        sel = members[assign == side]
        with np.errstate(divide="ignore", invalid="ignore"):
            mus.append(bits[sel].sum(axis=0) / float(len(sel)) if len(sel) else np.full(p, np.nan))
    return mus[0], mus[1]
