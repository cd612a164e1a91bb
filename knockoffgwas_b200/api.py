"""ctypes binding of the C ABI (include/kgw.h) and a host-side mirror of the reference's
`Knockoffs` interface (snpknock2/src/knockoffs.h:22-68) for the hot path.

No CPU fallback: every compute call goes through libkgw.so and fails loudly if the CUDA
extension is missing or no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB = None

KGW_ABI_VERSION = 4
STREAM_Z, STREAM_ZK, STREAM_HK, STREAM_FAMILY = 0, 1, 2, 3


class KgwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kgw error {code}: {msg}")
        self.code = code


class _Families(C.Structure):
    _fields_ = [("n_families", C.c_int32), ("reserved", C.c_int32), ("fam_off", C.c_void_p), ("fam_members", C.c_void_p),
                ("seg_off", C.c_void_p), ("seg_jmin", C.c_void_p), ("seg_jmax", C.c_void_p),
                ("seg_ids_off", C.c_void_p), ("seg_ids", C.c_void_p)]


class _Problem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("N", C.c_int32), ("p", C.c_int32), ("K", C.c_int32),
        ("W", C.c_int32), ("reserved0", C.c_int32), ("row_bytes", C.c_int64),
        ("H", C.c_void_p), ("ref_local", C.c_void_p), ("ref_global", C.c_void_p),
        ("b", C.c_void_p), ("mu", C.c_void_p), ("groups", C.c_void_p),
        ("win_start", C.c_void_p), ("win_end", C.c_void_p), ("unrelated", C.c_void_p),
        ("n_unrelated", C.c_int32), ("reserved1", C.c_int32), ("seed", C.c_uint64),
        ("injected_u", C.c_void_p), ("families", C.POINTER(_Families)),
    ]


class _Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("block_loci", C.c_int32), ("lanes_per_hap", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("stream", C.c_void_p), ("reserved", C.c_int32), ("segment_loci", C.c_int32),
                ("n_devices", C.c_int32), ("reserved2", C.c_int32), ("devices", C.c_void_p)]


class _PlanInfo(C.Structure):
    _fields_ = [("launches_per_run", C.c_int32), ("grid", C.c_int32), ("block", C.c_int32),
                ("lanes_per_hap", C.c_int32), ("states_per_lane", C.c_int32), ("block_loci", C.c_int32),
                ("segment_loci", C.c_int32), ("reserved", C.c_int32), ("scratch_bytes", C.c_int64), ("row_stride_bytes", C.c_int64), ("smem_bytes", C.c_int64)]


class _Debug(C.Structure):
    _fields_ = [("z", C.c_void_p), ("zk", C.c_void_p), ("beta_ckpt", C.c_void_p), ("ckpt_locus", C.c_void_p),
                ("n_ckpt", C.c_int32), ("reserved", C.c_int32), ("fam_z", C.c_void_p), ("fam_zk", C.c_void_p)]


@dataclass
class Options:
    device: int = 0
    block_loci: int = 0
    lanes_per_hap: int = 0
    ctas_per_sm: int = 0
    stream: int = 0   # cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream); 0 = plan-owned
    segment_loci: int = 0  # 0 = auto; multiple of 32: loci per segment of the backward/forward sweeps
    skip_sweeps: int = 0  # development timing experiments only: bit s set = skip sweep s+1 (results invalid)
    devices: tuple = ()   # several GPUs in one process (kgw_options.devices): CUDA ordinals; empty = `device` only

    def _c(self) -> _Options:
        o = _Options()
        if len(self.devices) > 0:
            self._dev = np.ascontiguousarray(self.devices, dtype=np.int32)   # kept alive by this object
            o.n_devices, o.devices = len(self._dev), self._dev.ctypes.data
        o.device, o.block_loci, o.lanes_per_hap, o.ctas_per_sm = self.device, self.block_loci, self.lanes_per_hap, self.ctas_per_sm
        o.stream = self.stream or None
        o.reserved = self.skip_sweeps
        o.segment_loci = self.segment_loci
        return o


def lib_path() -> Path:
    import os
    alt = os.environ.get("KGW_LIB")   # development: an alternative build of the same library (knockoffgwas_b200/build.py)
    return Path(alt) if alt else _PKG / "lib" / "libkgw.so"


def load_library():
    """Load libkgw.so (built in-tree by knockoffgwas_b200.build / __graft_entry__.build)."""
    global _LIB
    if _LIB is None:
        pth = lib_path()
        if not pth.exists():
            raise KgwError(-3, f"{pth} is missing: build the CUDA extension first "
                               f"(python -m knockoffgwas_b200.build); there is no CPU fallback")
        L = C.CDLL(str(pth))
        L.kgw_abi_version.restype = C.c_int32
        L.kgw_device_count.restype = C.c_int32
        L.kgw_estimate_hmm.restype = C.c_int32
        L.kgw_estimate_hmm.argtypes = [C.POINTER(_Problem), C.POINTER(_Options), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
        L.kgw_partition_work.restype = C.c_int32
        L.kgw_partition_work.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.kgw_last_error.restype = C.c_char_p
        L.kgw_uniform.restype = C.c_double
        L.kgw_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.kgw_generate.restype = C.c_int32
        L.kgw_generate.argtypes = [C.POINTER(_Problem), C.POINTER(_Options), C.c_void_p, C.POINTER(_Debug)]
        L.kgw_generate_strict.restype = C.c_int32
        L.kgw_generate_strict.argtypes = [C.POINTER(_Problem), C.c_int32, C.c_void_p, C.POINTER(_Debug)]
        L.kgw_plan_create.restype = C.c_int32
        L.kgw_plan_create.argtypes = [C.POINTER(_Problem), C.POINTER(_Options), C.POINTER(C.c_void_p)]
        L.kgw_plan_set_groups.restype = C.c_int32
        L.kgw_plan_set_groups.argtypes = [C.c_void_p, C.c_void_p]
        L.kgw_plan_run.restype = C.c_int32
        L.kgw_plan_run.argtypes = [C.c_void_p]
        L.kgw_plan_run_timed.restype = C.c_int32
        L.kgw_plan_run_timed.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
        L.kgw_plan_sync.restype = C.c_int32
        L.kgw_plan_sync.argtypes = [C.c_void_p]
        L.kgw_plan_download.restype = C.c_int32
        L.kgw_plan_download.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Debug)]
        L.kgw_plan_device_hk.restype = C.c_void_p
        L.kgw_plan_device_hk.argtypes = [C.c_void_p]
        L.kgw_plan_info.restype = C.c_int32
        L.kgw_plan_info.argtypes = [C.c_void_p, C.POINTER(_PlanInfo)]
        L.kgw_release_workspace.restype = None
        L.kgw_plan_destroy.restype = None
        L.kgw_plan_destroy.argtypes = [C.c_void_p]
        L.kgw_bed_column_bytes.restype = C.c_int64
        L.kgw_bed_column_bytes.argtypes = [C.c_int32]
        L.kgw_plan_encode_bed.restype = C.c_int32
        L.kgw_plan_encode_bed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_float)]
        L.kgw_bgen_read.restype = C.c_int32
        L.kgw_bgen_read.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int64, C.c_void_p, C.POINTER(C.c_float)]
        L.kgw_haps_read.restype = C.c_int32
        L.kgw_haps_read.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64,
                                    C.c_void_p, C.POINTER(C.c_float)]
        L.kgw_assign_references.restype = C.c_int32
        L.kgw_assign_references.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_float)]
        L.kgw_bed_encode.restype = C.c_int32
        L.kgw_bed_encode.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.kgw_kinship_create.restype = C.c_int32
        L.kgw_kinship_create.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.kgw_kinship_destroy.restype = None
        L.kgw_kinship_destroy.argtypes = [C.c_void_p]
        L.kgw_kinship_distances.restype = C.c_int32
        L.kgw_kinship_distances.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kgw_kinship_centroids.restype = C.c_int32
        L.kgw_kinship_centroids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        if L.kgw_abi_version() != KGW_ABI_VERSION:
            raise KgwError(-1, "libkgw.so ABI version mismatch")
        _LIB = L
    return _LIB


def device_count() -> int:
    return int(load_library().kgw_device_count())


def partition_work(n_unrelated: int, family_sizes, n_devices: int):
    """kgw_partition_work: (cut int32 [n_devices+1] into the sorted unrelated list, device of every family)."""
    sizes = np.asarray(family_sizes, dtype=np.int64)
    off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(sizes)]), dtype=np.int32)
    cut = np.zeros(n_devices + 1, dtype=np.int32)
    dev = np.zeros(max(1, len(sizes)), dtype=np.int32)
    _check(load_library().kgw_partition_work(int(n_unrelated), len(sizes), off.ctypes.data, int(n_devices), cut.ctypes.data,
                                             dev.ctypes.data))
    return cut, dev[:len(sizes)]


def abi_version() -> int:
    return load_library().kgw_abi_version()


def uniform(seed: int, hap: int, stream: int, locus: int) -> float:
    """Host replay of the device generator."""
    return load_library().kgw_uniform(int(seed), int(hap), int(stream), int(locus))


def release_workspace():
    """Free the scratch slabs parked by destroyed plans (kgw_release_workspace)."""
    load_library().kgw_release_workspace()


def _check(rc: int):
    if rc != 0:
        raise KgwError(rc, (load_library().kgw_last_error() or b"").decode())


class _Held:
    """Keeps the numpy buffers behind a kgw_problem alive."""

    def __init__(self, H, p, K, ref_local, b, mu, groups, win_start=None, win_end=None, unrelated=None,
                 ref_global=None, seed=2020, injected_u=None, families=None):
        self.H = np.ascontiguousarray(H, dtype=np.uint8)
        assert self.H.ndim == 2
        self.N, self.row_bytes = (int(x) for x in self.H.shape)
        self.p, self.K = int(p), int(K)
        if win_start is None:
            win_start, win_end = [0], [self.p]
        self.win_start = np.ascontiguousarray(win_start, dtype=np.int32)
        self.win_end = np.ascontiguousarray(win_end, dtype=np.int32)
        self.W = len(self.win_start)
        self.ref_local = np.ascontiguousarray(ref_local, dtype=np.int32).reshape(self.N, self.W, self.K)
        self.ref_global = None if ref_global is None else np.ascontiguousarray(ref_global, dtype=np.int32).reshape(self.N, self.K)
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        self.groups = np.ascontiguousarray(groups, dtype=np.int32)
        assert self.b.shape == (self.p,) and self.mu.shape == (self.p,) and self.groups.shape == (self.p,)
        if unrelated is None:
            # every haplotype without IBD relatives (metadata.unrelated, metadata.cpp:242-293)
            unrelated = np.arange(self.N, dtype=np.int32)
            if families:
                related = np.concatenate([np.asarray(m, dtype=np.int32) for m, _ in families])
                unrelated = np.setdiff1d(unrelated, related).astype(np.int32)
        self.unrelated = np.ascontiguousarray(unrelated, dtype=np.int32)
        self.seed = int(seed)
        self.injected_u = None
        if injected_u is not None:
            self.injected_u = np.ascontiguousarray(injected_u, dtype=np.float64)
            assert self.injected_u.shape in ((self.N, self.W, 3, self.p),) + (((self.N, 3, self.p),) if self.W == 1 else ())
        self.fam = None
        if families:
            # families: [(members, [(j_min, j_max, ..., ids) | (j_min, j_max, ids), ...]), ...] -- member lists in
            # metadata.related_families order, TIDIED segments in the reference's std::set<IbdSeg> order
            i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
            mem = [i32(m) for m, _ in families]
            segs = [sg for _, sl in families for sg in sl]
            ids = [i32(sg[-1]) for sg in segs]
            self._fam_arrays = dict(
                fam_off=i32(np.concatenate([[0], np.cumsum([len(m) for m in mem])])),
                fam_members=i32(np.concatenate(mem)),
                seg_off=i32(np.concatenate([[0], np.cumsum([len(sl) for _, sl in families])])),
                seg_jmin=i32([sg[0] for sg in segs]), seg_jmax=i32([sg[1] for sg in segs]),
                seg_ids_off=i32(np.concatenate([[0], np.cumsum([len(x) for x in ids])])),
                seg_ids=i32(np.concatenate(ids)) if ids else i32([]))
            self.fam = _Families()
            self.fam.n_families = len(families)
            for k, v in self._fam_arrays.items():
                setattr(self.fam, k, v.ctypes.data)
            self.n_family_members = int(len(self._fam_arrays["fam_members"]))
            self.family_members = self._fam_arrays["fam_members"]

    def c(self) -> _Problem:
        P = _Problem()
        P.abi_version = KGW_ABI_VERSION
        P.N, P.p, P.K, P.W = self.N, self.p, self.K, self.W
        P.row_bytes = self.row_bytes
        P.H = self.H.ctypes.data
        P.ref_local = self.ref_local.ctypes.data
        P.ref_global = None if self.ref_global is None else self.ref_global.ctypes.data
        P.b, P.mu, P.groups = self.b.ctypes.data, self.mu.ctypes.data, self.groups.ctypes.data
        P.win_start, P.win_end = self.win_start.ctypes.data, self.win_end.ctypes.data
        P.unrelated = self.unrelated.ctypes.data
        P.n_unrelated = len(self.unrelated)
        P.seed = self.seed
        P.injected_u = None if self.injected_u is None else self.injected_u.ctypes.data
        P.families = C.pointer(self.fam) if self.fam is not None else None
        return P


@dataclass
class DebugOut:
    z: np.ndarray | None = None
    zk: np.ndarray | None = None
    beta_ckpt: np.ndarray | None = None
    ckpt_locus: np.ndarray | None = None
    fam_z: np.ndarray | None = None     # [family members in input order][p]
    fam_zk: np.ndarray | None = None


def _make_debug(held: _Held, want_paths: bool, want_beta: bool):
    if not (want_paths or want_beta):
        return None, None
    n = len(held.unrelated)
    out = DebugOut()
    d = _Debug()
    if want_paths:
        out.z = np.zeros((n, held.p), dtype=np.int32)
        out.zk = np.zeros((n, held.p), dtype=np.int32)
        d.z, d.zk = out.z.ctypes.data, out.zk.ctypes.data
        if held.fam is not None:
            out.fam_z = np.zeros((held.n_family_members, held.p), dtype=np.int32)
            out.fam_zk = np.zeros((held.n_family_members, held.p), dtype=np.int32)
            d.fam_z, d.fam_zk = out.fam_z.ctypes.data, out.fam_zk.ctypes.data
    if want_beta:
        cap = held.p
        out.beta_ckpt = np.zeros((n, cap, held.K), dtype=np.float64)
        out.ckpt_locus = np.full(cap, -1, dtype=np.int32)
        d.beta_ckpt, d.ckpt_locus, d.n_ckpt = out.beta_ckpt.ctypes.data, out.ckpt_locus.ctypes.data, cap
    return d, out


def _finish_debug(d, out, held):
    if out is not None and out.beta_ckpt is not None:
        n = int(d.n_ckpt)
        # the library wrote rows with stride n (its own nblk), re-view accordingly
        flat = out.beta_ckpt.reshape(-1)[: len(held.unrelated) * n * held.K]
        out.beta_ckpt = flat.reshape(len(held.unrelated), n, held.K).copy()
        out.ckpt_locus = out.ckpt_locus[:n].copy()
    return out


def generate(H, p, K, ref_local, b, mu, groups, *, win_start=None, win_end=None, unrelated=None, ref_global=None,
             seed=2020, injected_u=None, families=None, options: Options | None = None, debug_paths=False,
             debug_beta=False):
    """One-shot drop-in for Knockoffs::generate() (host buffers in, host buffer out).
    Returns (Hk uint8 [N][row_bytes], DebugOut | None)."""
    L = load_library()
    held = _Held(H, p, K, ref_local, b, mu, groups, win_start, win_end, unrelated, ref_global, seed, injected_u,
                 families)
    P = held.c()
    o = (options or Options())._c()
    hk = np.zeros_like(held.H)
    d, out = _make_debug(held, debug_paths, debug_beta)
    _check(L.kgw_generate(C.byref(P), C.byref(o), hk.ctypes.data, C.byref(d) if d is not None else None))
    return hk, _finish_debug(d, out, held)


def generate_strict(H, p, K, ref_local, b, mu, groups, *, win_start=None, win_end=None, unrelated=None, seed=2020,
                    injected_u=None, device=0, debug_paths=False, debug_beta=False):
    """kgw_generate_strict(): the unrelated sampler in the reference's summation order (audit mode; one thread per
    haplotype, slow).  Same arguments and outputs as generate(); debug_beta returns every backward row [n][p][K]."""
    L = load_library()
    held = _Held(H, p, K, ref_local, b, mu, groups, win_start, win_end, unrelated, None, seed, injected_u, None)
    P = held.c()
    hk = np.zeros_like(held.H)
    d, out = _make_debug(held, debug_paths, debug_beta)
    _check(L.kgw_generate_strict(C.byref(P), C.c_int32(device), hk.ctypes.data, C.byref(d) if d is not None else None))
    return hk, _finish_debug(d, out, held)


def generate_into(H, p, K, ref_local, b, mu, groups, *, out, win_start=None, win_end=None, unrelated=None,
                  ref_global=None, seed=2020, options: Options | None = None):
    """kgw_generate() straight on caller buffers (no copies on the Python side when the arrays are
    already contiguous and of the right dtype -- e.g. views of pinned torch tensors); `out` is
    uint8 [N][row_bytes] and receives the rows listed in `unrelated`."""
    L = load_library()
    held = _Held(H, p, K, ref_local, b, mu, groups, win_start, win_end, unrelated, ref_global, seed, None)
    assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == held.H.shape
    P = held.c()
    o = (options or Options())._c()
    _check(L.kgw_generate(C.byref(P), C.byref(o), out.ctypes.data, None))
    return out


def estimate_hmm(H, p, K, ref_local, b, *, mu_init=None, win_start=None, win_end=None, options: Options | None = None):
    """kgw_estimate_hmm (Knockoffs::EM, knockoffs.cpp:706-960): returns (mu after EM, steps run, per-step deltas, converged)."""
    L = load_library()
    p = int(p)
    held = _Held(H, p, K, ref_local, b, np.full(p, 1e-5), np.arange(p, dtype=np.int32), win_start, win_end)
    P = held.c()
    o = (options or Options())._c()
    mu0 = np.ascontiguousarray(np.full(p, 1e-5) if mu_init is None else mu_init, dtype=np.float64)
    out = np.zeros(p, dtype=np.float64)
    deltas = np.zeros(20, dtype=np.float64)
    it = C.c_int32(0)
    rc = L.kgw_estimate_hmm(C.byref(P), C.byref(o), mu0.ctypes.data, out.ctypes.data, C.byref(it), deltas.ctypes.data)
    if rc < 0:
        _check(rc)
    n = int(it.value)
    return out, n, deltas[:max(n, 1)].copy(), rc == 0


def swap_flags(num_snps: int, seed: int = 2019) -> np.ndarray:
    """The column-order coin flips of plink::write_binary (plink_interface.cpp:267-274):
    std::mt19937_64(2019) through std::uniform_int_distribution<int>(0, 1).  With a 64-bit engine libstdc++
    (GCC >= 11, bits/uniform_int_dist.h) maps one draw x to the high half of the 128-bit product x * 2,
    i.e. the top bit of x."""
    n, w, m, r = 312, 64, 156, 31
    a = 0xB5026F5AA96619E9
    mask = (1 << 64) - 1
    mt = [0] * n
    mt[0] = seed & mask
    for i in range(1, n):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & mask
    idx = n
    out = np.zeros(num_snps, dtype=np.uint8)
    lower, upper = (1 << r) - 1, mask ^ ((1 << r) - 1)
    for j in range(num_snps):
        if idx >= n:
            for i in range(n):
                x = (mt[i] & upper) | (mt[(i + 1) % n] & lower)
                xa = x >> 1
                if x & 1:
                    xa ^= a
                mt[i] = mt[(i + m) % n] ^ xa
            idx = 0
        y = mt[idx]
        idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        out[j] = (y & mask) >> 63
    return out


def bed_encode(H, Hk, p, swap=None, device=0):
    """One-shot kgw_bed_encode from host matrices; H may be None for the Hk-only file."""
    L = load_library()
    Hk = np.ascontiguousarray(Hk, dtype=np.uint8)
    N, row_bytes = Hk.shape
    both = H is not None
    if both:
        H = np.ascontiguousarray(H, dtype=np.uint8)
        assert H.shape == Hk.shape
    nb = (N // 2 + 3) // 4
    bed = np.empty((2 * p if both else p, nb), dtype=np.uint8)
    sw = None
    if both:
        sw = np.ascontiguousarray(swap_flags(p) if swap is None else swap, dtype=np.uint8)
    _check(L.kgw_bed_encode(N, int(p), row_bytes, H.ctypes.data if both else None, Hk.ctypes.data,
                            None if sw is None else sw.ctypes.data, int(device), bed.ctypes.data))
    return bed


def assign_references(H, p, K, clusters, family_of=None, compression=10, device=0):
    """kgw_assign_references: K nearest haplotypes inside the kinship cluster (findNeighbors,
    kinship_computer.cpp:174-233) for every haplotype; returns (int32 [N][K] in (distance, cluster position)
    order, device milliseconds).  `clusters` = list of haplotype-id arrays in the reference's cluster order."""
    L = load_library()
    H = np.ascontiguousarray(H, dtype=np.uint8)
    N, row_bytes = H.shape
    off = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int32)
    mem = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32) for c in clusters]), dtype=np.int32)
    fam = None if family_of is None else np.ascontiguousarray(family_of, dtype=np.int32)
    out = np.empty((N, int(K)), dtype=np.int32)
    ms = C.c_float()
    _check(L.kgw_assign_references(N, int(p), row_bytes, H.ctypes.data, int(K), int(compression), len(clusters),
                                   off.ctypes.data, mem.ctypes.data, None if fam is None else fam.ctypes.data,
                                   int(device), out.ctypes.data, C.byref(ms)))
    return out, float(ms.value)


class KinshipMatrix:
    """The thinned rows of all chromosomes resident in HBM for the data passes of the bifurcating k-means
    (kgw_kinship_*; KinshipComputer::update_assignments / update_centroids, kinship_computer.cpp:314-405)."""

    def __init__(self, H, p, device=0):
        self._L = load_library()
        self.H = np.ascontiguousarray(H, dtype=np.uint8)
        self.N, self.row_bytes = (int(x) for x in self.H.shape)
        self.p = int(p)
        self._h = C.c_void_p()
        _check(self._L.kgw_kinship_create(self.N, self.p, self.row_bytes, self.H.ctypes.data, int(device), C.byref(self._h)))

    def distances(self, members, mu0, mu1, use=None):
        """(d0, d1): distance of every member's sample (both haplotypes) to the two centroids, sequential fp64 sums."""
        mem = np.ascontiguousarray(members, dtype=np.int32)
        a = np.ascontiguousarray(mu0, dtype=np.float64)
        b = np.ascontiguousarray(mu1, dtype=np.float64)
        assert a.shape == (self.p,) and b.shape == (self.p,)
        u = None if use is None else np.ascontiguousarray(use, dtype=np.uint8)
        d0, d1 = np.empty(len(mem)), np.empty(len(mem))
        _check(self._L.kgw_kinship_distances(self._h, mem.ctypes.data, len(mem), a.ctypes.data, b.ctypes.data,
                                             None if u is None else u.ctypes.data, d0.ctypes.data, d1.ctypes.data))
        return d0, d1

    def centroids(self, members, assign):
        """(mu0, mu1): column means of the members with assign 0 / 1."""
        mem = np.ascontiguousarray(members, dtype=np.int32)
        asg = np.ascontiguousarray(assign, dtype=np.int32)
        mu0, mu1 = np.empty(self.p), np.empty(self.p)
        _check(self._L.kgw_kinship_centroids(self._h, mem.ctypes.data, asg.ctypes.data, len(mem), mu0.ctypes.data, mu1.ctypes.data, None, None))
        return mu0, mu1

    def close(self):
        if self._h:
            self._L.kgw_kinship_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BgenReader:
    """Host mirror of the reference's BgenReader (bgen_reader.h, bgen_reader.cpp:8-80): opens `<prefix>.bgen`
    (+ `<prefix>.sample`, two header lines then `ID_1 ID_2 ...`, metadata.cpp:131-147), keeps the variant ids of
    the file, and `read(sample_ids, snp_ids)` returns the packed haplotype matrix uint8 [2 len(sample_ids)]
    [ceil(len(snp_ids)/8)] in chaplotype layout through kgw_bgen_read.  The id bookkeeping (match_indices,
    the rsid look-ups) is host work as in the reference, with dictionaries instead of std::find."""

    def __init__(self, prefix, sample_file=None):
        import struct
        self.path = str(prefix) if str(prefix).endswith(".bgen") else str(prefix) + ".bgen"
        self.image = np.fromfile(self.path, dtype=np.uint8)
        d = self.image.tobytes()
        if len(d) < 24:
            raise KgwError(-1, "BGEN image shorter than its header")
        offset, hlen, M, N = struct.unpack_from("<IIII", d, 0)
        self.n_variants, self.n_samples = int(M), int(N)
        rsid, pos = [], 4 + offset
        for _ in range(M):                      # variant identifying data (BGEN v1.2): id, rsid, chromosome, position, alleles
            (l,) = struct.unpack_from("<H", d, pos); pos += 2 + l
            (l,) = struct.unpack_from("<H", d, pos); rsid.append(d[pos + 2:pos + 2 + l].decode()); pos += 2 + l
            (l,) = struct.unpack_from("<H", d, pos); pos += 2 + l
            (_, nall) = struct.unpack_from("<IH", d, pos); pos += 6
            for _a in range(nall):
                (l,) = struct.unpack_from("<I", d, pos); pos += 4 + l
            (c,) = struct.unpack_from("<I", d, pos); pos += 4 + c
        self.variant_ids = rsid
        sf = sample_file if sample_file is not None else self.path[:-5] + ".sample"
        self.sample_ids = None
        try:
            with open(sf) as fh:
                rows = [ln.split() for ln in fh.read().splitlines() if ln.strip()]
            self.sample_ids = [r[0] for r in rows[2:]]
        except OSError:
            pass
        if self.sample_ids is not None and len(self.sample_ids) != self.n_samples:
            raise KgwError(-1, "Error loading BGEN file: the number of samples does not match.")   # bgen_reader.cpp:31-34

    def read(self, sample_ids=None, snp_ids=None, device=0, host_threads=0, return_ms=False):
        if sample_ids is None:
            sidx = np.arange(self.n_samples, dtype=np.int32)
        else:
            if self.sample_ids is None:
                raise KgwError(-1, "sample ids requested but no .sample file was found")
            where = {}
            for i, sid in enumerate(self.sample_ids):
                where.setdefault(sid, i)
            try:
                sidx = np.array([where[str(x)] for x in sample_ids], dtype=np.int32)
            except KeyError as e:
                raise KgwError(-1, f"sample {e.args[0]} is not in the BGEN sample file") from None
        if snp_ids is None:
            vcol = np.arange(self.n_variants, dtype=np.int32)
        else:
            first = {}
            for v, rid in enumerate(self.variant_ids):
                first.setdefault(rid, v)
            try:
                vcol = np.array([first[str(x)] for x in snp_ids], dtype=np.int32)
            except KeyError:
                raise KgwError(-1, "Error loading BGEN file. Some requested variants were not found.") from None   # :238-241
        return read_bgen_image(self.image, sidx, vcol, device=device, host_threads=host_threads, return_ms=return_ms)


def read_bgen_image(image, sample_idx, variant_of_col, device=0, host_threads=0, return_ms=False):
    """kgw_bgen_read on an in-memory .bgen image: uint8 [2 n_samples][ceil(p/8)] chaplotype rows."""
    L = load_library()
    image = np.ascontiguousarray(image, dtype=np.uint8)
    sidx = np.ascontiguousarray(sample_idx, dtype=np.int32)
    vcol = np.ascontiguousarray(variant_of_col, dtype=np.int32)
    p = len(vcol)
    H = np.zeros((2 * len(sidx), (p + 7) // 8), dtype=np.uint8)
    ms = C.c_float()
    _check(L.kgw_bgen_read(image.ctypes.data, image.nbytes, len(sidx), sidx.ctypes.data, p, vcol.ctypes.data, int(host_threads),
                           int(device), H.shape[1], H.ctypes.data, C.byref(ms)))
    return (H, float(ms.value)) if return_ms else H


class HapsReader:
    """Host mirror of the reference's HapsReader (haps_reader.h, haps_reader.cpp:8-35): `<prefix>.haps` (one row per SNP),
    `<prefix>.legend` (header, then the variant id in the first column), `<prefix>.sample` (two header lines, then
    `ID_1 ...`); `read(sample_ids, snp_ids)` returns the packed haplotype matrix through kgw_haps_read."""

    def __init__(self, prefix):
        prefix = str(prefix)
        self.text = np.fromfile(prefix + ".haps", dtype=np.uint8)
        with open(prefix + ".legend") as fh:
            self.variant_ids = [ln.split()[0] for ln in fh.read().splitlines()[1:] if ln.strip()]
        with open(prefix + ".sample") as fh:
            self.sample_ids = [ln.split()[0] for ln in fh.read().splitlines()[2:] if ln.strip()]

    def read(self, sample_ids=None, snp_ids=None, device=0, return_ms=False):
        if sample_ids is None:
            sidx = np.arange(len(self.sample_ids), dtype=np.int32)
        else:
            where = {}
            for i, sid in enumerate(self.sample_ids):
                where.setdefault(sid, i)
            try:
                sidx = np.array([where[str(x)] for x in sample_ids], dtype=np.int32)
            except KeyError as e:
                raise KgwError(-1, f"Error parsing HAPS file. Could not find sample {e.args[0]}") from None   # haps_reader.cpp:82-84
        if snp_ids is None:
            rows = np.arange(len(self.variant_ids), dtype=np.int32)
        else:
            first = {}
            for r, vid in enumerate(self.variant_ids):
                first.setdefault(vid, r)
            try:
                rows = np.array([first[str(x)] for x in snp_ids], dtype=np.int32)
            except KeyError as e:
                raise KgwError(-1, f"variant {e.args[0]} is not in the legend") from None
        return read_haps_text(self.text, sidx, rows, device=device, return_ms=return_ms)


def read_haps_text(text, sample_idx, row_of_col, device=0, return_ms=False):
    """kgw_haps_read on an in-memory .haps file: uint8 [2 n_samples][ceil(p/8)] chaplotype rows."""
    L = load_library()
    text = np.ascontiguousarray(np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text, dtype=np.uint8)
    sidx = np.ascontiguousarray(sample_idx, dtype=np.int32)
    rows = np.ascontiguousarray(row_of_col, dtype=np.int32)
    p = len(rows)
    H = np.zeros((2 * len(sidx), (p + 7) // 8), dtype=np.uint8)
    ms = C.c_float()
    _check(L.kgw_haps_read(text.ctypes.data, text.nbytes, len(sidx), sidx.ctypes.data, p, rows.ctypes.data, int(device),
                           H.shape[1], H.ctypes.data, C.byref(ms)))
    return (H, float(ms.value)) if return_ms else H


def combine_references_families(ref, families):
    """Host step after kgw_assign_references, KinshipComputer::combine_references_families
    (kinship_computer.cpp:765-835): lists sorted ascending; the members of an IBD family share the intersection of
    their lists, topped up round-robin with each member's next-closest reference until it holds K.
    ref: int32 [N][K] in distance order; families[i] = metadata.related_families[i]."""
    ref = np.asarray(ref, dtype=np.int32)
    N, K = ref.shape
    out = np.sort(ref, axis=1)
    done = np.zeros(N, dtype=bool)
    for i1 in range(N):
        if done[i1]:
            continue
        fam = [int(x) for x in families[i1]]
        done[i1] = True
        done[fam] = True
        if len(fam) <= 1:
            continue
        inter = set(int(x) for x in out[i1])
        for i2 in fam:
            if i2 != i1:
                inter &= set(int(x) for x in out[i2])
        inter = sorted(inter)
        i = k = 0
        while len(inter) < K:
            cand = int(ref[fam[i], k])
            if cand not in inter:
                inter.append(cand)
                inter.sort()
            if i + 1 < len(fam):
                i += 1
            else:
                i, k = 0, k + 1
        out[fam] = inter
    return out


class Plan:
    """Resident-data plan: inputs stay in HBM across runs / partitions (kgw_plan_*)."""

    def __init__(self, H, p, K, ref_local, b, mu, groups, *, win_start=None, win_end=None, unrelated=None,
                 ref_global=None, seed=2020, injected_u=None, families=None, options: Options | None = None):
        self._L = load_library()
        self._held = _Held(H, p, K, ref_local, b, mu, groups, win_start, win_end, unrelated, ref_global, seed,
                           injected_u, families)
        P = self._held.c()
        o = (options or Options())._c()
        self._h = C.c_void_p()
        _check(self._L.kgw_plan_create(C.byref(P), C.byref(o), C.byref(self._h)))

    def set_groups(self, groups):
        g = np.ascontiguousarray(groups, dtype=np.int32)
        assert g.shape == (self._held.p,)
        self._held.groups = g
        _check(self._L.kgw_plan_set_groups(self._h, g.ctypes.data))

    def run(self):
        _check(self._L.kgw_plan_run(self._h))

    def run_timed(self, iters=1) -> float:
        ms = C.c_float()
        _check(self._L.kgw_plan_run_timed(self._h, int(iters), C.byref(ms)))
        return float(ms.value)

    def sync(self):
        _check(self._L.kgw_plan_sync(self._h))

    def download(self, debug_paths=False, debug_beta=False, out=None):
        hk = np.zeros_like(self._held.H) if out is None else out
        d, dbg = _make_debug(self._held, debug_paths, debug_beta)
        _check(self._L.kgw_plan_download(self._h, hk.ctypes.data, C.byref(d) if d is not None else None))
        return hk, _finish_debug(d, dbg, self._held)

    def encode_bed(self, swap=None, include_genotypes=True, out=None):
        """Body of `<out>_res<r>.bed` (writeBED, plink_interface.cpp:67-103) from the matrices resident in the
        plan; returns (uint8 [columns][ceil(N/2/4)], device milliseconds of the transpose)."""
        N, p = self._held.N, self._held.p
        nb = (N // 2 + 3) // 4
        ncols = 2 * p if include_genotypes else p
        bed = np.empty((ncols, nb), dtype=np.uint8) if out is None else out
        assert bed.shape == (ncols, nb) and bed.dtype == np.uint8 and bed.flags.c_contiguous
        sw = None
        if include_genotypes:
            sw = np.ascontiguousarray(swap_flags(p) if swap is None else swap, dtype=np.uint8)
            assert sw.shape == (p,)
        ms = C.c_float()
        _check(self._L.kgw_plan_encode_bed(self._h, None if sw is None else sw.ctypes.data, int(bool(include_genotypes)),
                                           bed.ctypes.data, C.byref(ms)))
        return bed, float(ms.value)

    def device_hk_ptr(self) -> int:
        return int(self._L.kgw_plan_device_hk(self._h) or 0)

    def info(self) -> dict:
        pi = _PlanInfo()
        _check(self._L.kgw_plan_info(self._h, C.byref(pi)))
        return {name: int(getattr(pi, name)) for name, _ in _PlanInfo._fields_}

    def close(self):
        if self._h:
            self._L.kgw_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Knockoffs:
    """Host-side mirror of the reference's `Knockoffs` object for the hot path
    (snpknock2/src/knockoffs.h:22-68): same method names and argument meaning; `generate()`
    is the call that runs on the B200.  Everything else stays cheap host logic."""

    def __init__(self, H_packed, num_snps, K, seed=2020, genetic_dist=None, partitions=None,
                 win_start=None, win_end=None, unrelated=None, options: Options | None = None):
        self.H = np.ascontiguousarray(H_packed, dtype=np.uint8)
        self.num_haps, self.num_snps, self.K, self.seed = self.H.shape[0], int(num_snps), int(K), int(seed)
        self.genetic_dist = None if genetic_dist is None else np.asarray(genetic_dist, dtype=np.float64)
        self.partitions = partitions
        self.win_start = [0] if win_start is None else list(win_start)
        self.win_end = [self.num_snps] if win_end is None else list(win_end)
        self.num_windows = len(self.win_start)
        self.unrelated = unrelated
        self.families = None          # [(members, tidied segments)] as generate_related_cluster sees them
        self.options = options
        self.references_local = None
        self.references_global = None
        self.b = self.mut_rates = self.rec_rates = None
        self.groups = np.arange(self.num_snps, dtype=np.int32)
        self.Hk = None
        if partitions is not None:
            self.set_partition(0)
        if self.genetic_dist is not None:
            self.init_hmm(self.K, 1.0, 1.0e-4)  # knockoffs.cpp:60 (constructor default)

    def set_references(self, references_local):
        """knockoffs.cpp:109-126: each K-list is sorted ascending."""
        r = np.asarray(references_local, dtype=np.int32).reshape(self.num_haps, self.num_windows, self.K)
        self.references_local = np.sort(r, axis=2)

    def set_references_global(self, references_global):
        r = np.asarray(references_global, dtype=np.int32).reshape(self.num_haps, self.K)
        self.references_global = np.sort(r, axis=1)

    def init_hmm(self, K, rho, lam):
        """knockoffs.cpp:1508-1525: b_j = exp(-rho d_j), b_0 = 0, mu_j = lambda."""
        assert self.genetic_dist is not None
        self.rec_rates = np.ones(self.num_snps)
        self.rec_rates[1:] = rho * self.genetic_dist[1:]
        self.b = np.zeros(self.num_snps)
        self.b[1:] = np.exp(-self.rec_rates[1:])
        self.mut_rates = np.full(self.num_snps, float(lam))

    def load_hmm(self, rec_rates, mut_rates):
        """knockoffs.cpp:1593-1610: b_j = exp(-rec_j) for every j; mu clamped to [1e-6, 1e-3]."""
        self.rec_rates = np.asarray(rec_rates, dtype=np.float64).copy()
        self.b = np.exp(-self.rec_rates)
        self.mut_rates = np.maximum(1e-6, np.minimum(np.asarray(mut_rates, dtype=np.float64), 1e-3))

    def EM(self, num_threads=1):
        """Knockoffs::EM (knockoffs.cpp:706-748): re-estimates the mutation rates on the device and installs them
        (update_hmm, :1550-1561).  Returns (steps run, per-step deltas, converged)."""
        assert self.references_local is not None and self.b is not None
        mu, steps, deltas, conv = estimate_hmm(self.H, self.num_snps, self.K, self.references_local, self.b,
                                               mu_init=self.mut_rates, win_start=self.win_start, win_end=self.win_end,
                                               options=self.options)
        self.mut_rates = mu
        return steps, deltas, conv

    def set_partition(self, r):
        self.set_groups(self.partitions[r])

    def set_groups(self, groups):
        self.groups = np.ascontiguousarray(groups, dtype=np.int32)

    def num_groups(self):
        return int(self.groups[-1]) + 1

    def generate(self, num_threads=1):
        """Knockoffs::generate (knockoffs.cpp:311).  num_threads is accepted for signature parity
        and ignored: the result no longer depends on it."""
        assert self.references_local is not None and self.b is not None
        self.Hk, _ = generate(self.H, self.num_snps, self.K, self.references_local, self.b, self.mut_rates,
                              self.groups, win_start=self.win_start, win_end=self.win_end, unrelated=self.unrelated,
                              ref_global=self.references_global, seed=self.seed, families=self.families,
                              options=self.options)
        return self.Hk

    def audit(self, n_haps=64):
        """Samples `n_haps` evenly spaced unrelated haplotypes again in the reference's own summation order
        (kgw_generate_strict) and returns the ones whose knockoff row differs from the last generate()'s -- an empty
        array when the fast path drew exactly what a strict-IEEE build of the reference would."""
        assert self.Hk is not None, "generate() first"
        unrel = self.unrelated
        if unrel is None:
            unrel = np.arange(self.num_haps, dtype=np.int32)
            if self.families:
                unrel = np.setdiff1d(unrel, np.concatenate([np.asarray(m) for m, _ in self.families])).astype(np.int32)
        unrel = np.asarray(unrel, dtype=np.int32)
        n = min(int(n_haps), len(unrel))
        if n <= 0:
            return np.zeros(0, dtype=np.int32)
        pick = unrel[(np.arange(n, dtype=np.int64) * len(unrel)) // n]
        dev = self.options.device if self.options is not None else 0
        hs, _ = generate_strict(self.H, self.num_snps, self.K, self.references_local, self.b, self.mut_rates, self.groups,
                                win_start=self.win_start, win_end=self.win_end, unrelated=pick, seed=self.seed, device=dev)
        return pick[(hs[pick] != self.Hk[pick]).any(axis=1)]
