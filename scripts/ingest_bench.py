#!/usr/bin/env python
"""Development helper: device time of the ingest kernels (kgw_bgen_read, kgw_haps_read) and of the reference-selection
kernels on a population larger than bench.py's (whose ingest legs run at C2 size, where launches take a fraction of a
millisecond).
usage: python scripts/ingest_bench.py [--haps N] [--snps p]"""
import argparse, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--haps", type=int, default=40000)
ap.add_argument("--snps", type=int, default=30000)
ap.add_argument("--zlib-snps", type=int, default=2000, help="variants of the zlib-compressed image (0 skips that leg)")
a = ap.parse_args()
N, p = a.haps, a.snps
rng = np.random.default_rng(7)
rb = (p + 7) // 8
H = rng.integers(0, 256, size=(N, rb), dtype=np.uint8)
if p % 8:
    H[:, -1] &= (1 << (p % 8)) - 1
t0 = time.time()
image = synth.write_bgen_image(H, p, compression=0)
sidx, vcol = np.arange(N // 2, dtype=np.int32), np.arange(p, dtype=np.int32)
kg.read_bgen_image(image, sidx, vcol)
Hb, ms = kg.read_bgen_image(image, sidx, vcol, return_ms=True)
blk = p * (10 + N // 2 + N)
print(f"bgen: {N} x {p}: image {image.nbytes / 1e9:.2f} GB (built in {time.time() - t0:.0f} s), kernels {ms:.3f} ms = "
      f"{(blk + Hb.nbytes) / (ms * 1e-3) / 1e12:.2f} TB/s, identical {bool(np.array_equal(Hb, H))}", flush=True)
del image
if a.zlib_snps:
    import os
    pz = min(p, a.zlib_snps)
    Hz = np.ascontiguousarray(H[:, :(pz + 7) // 8])
    if pz % 8:
        Hz[:, -1] &= (1 << (pz % 8)) - 1
    image = synth.write_bgen_image(Hz, pz, compression=1)
    vz = np.arange(pz, dtype=np.int32)
    for mode in ("device", "host"):
        os.environ["KGW_BGEN_HOST_INFLATE"] = "1" if mode == "host" else "0"
        kg.read_bgen_image(image, sidx, vz)
        t0 = time.perf_counter()
        Hb, ms = kg.read_bgen_image(image, sidx, vz, return_ms=True)
        dt = time.perf_counter() - t0
        print(f"bgen zlib ({mode} inflate): {N} x {pz}: image {image.nbytes / 1e6:.1f} MB -> {pz * (10 + N // 2 + N) / 1e6:.1f} MB of blocks, "
              f"whole call {dt * 1e3:.1f} ms, kernels {ms:.2f} ms, identical {bool(np.array_equal(Hb, Hz))}", flush=True)
    del image
bits = np.unpackbits(H, axis=1, bitorder="little")[:, :p]
txt = np.full((p, 2 * N), ord(" "), dtype=np.uint8)
txt[:, 0::2] = bits.T + ord("0")
txt[:, -1] = ord("\n")
txt = txt.reshape(-1)
del bits
kg.read_haps_text(txt, sidx, vcol)
Ht, ms = kg.read_haps_text(txt, sidx, vcol, return_ms=True)
print(f"haps: text {txt.nbytes / 1e9:.2f} GB, kernels {ms:.3f} ms = {(txt.nbytes + Ht.nbytes) / (ms * 1e-3) / 1e12:.2f} TB/s, "
      f"identical {bool(np.array_equal(Ht, H))}", flush=True)
