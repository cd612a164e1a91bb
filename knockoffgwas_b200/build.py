"""In-tree build of the CUDA library (sm_100a only).  `python -m knockoffgwas_b200.build`."""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libkgw.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # host-side tables must follow the reference's evaluation order: no FMA contraction, no fast-math
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
]
OBJDIR = PKG / "build"


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "kgw.h"]
    return any(d.stat().st_mtime > t for d in deps)


def _compile(args):
    nvcc, src, obj, verbose = args
    import os
    fast = ["-DKGW_FAST"] if os.environ.get("KGW_FAST") == "1" else []   # development: benchmark variants only
    if os.environ.get("KGW_PHASES") == "1":
        fast.append("-DKGW_PHASES")
    if os.environ.get("KGW_CFLAGS"):
        fast += os.environ["KGW_CFLAGS"].split()
    if os.environ.get("KGW_FAM_MINB"):
        fast.append("-DKGW_FAM_MINB=" + os.environ["KGW_FAM_MINB"])
    if os.environ.get("KGW_M_OVERRIDE") and src.stem.endswith("_L" + os.environ.get("KGW_M_OVERRIDE_L", "4")):
        fast.append("-DKGW_M_OVERRIDE=" + os.environ["KGW_M_OVERRIDE"])
    cmd = [nvcc] + NVCC_FLAGS + fast + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", str(obj), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per translation unit, compiled in parallel (the kernel instantiations are split by
    lanes-per-haplotype so that the build takes about as long as its slowest unit), then one link.
    Development: KGW_BUILD_TAG=<tag> builds lib/libkgw_<tag>.so from its own object directory (e.g. with KGW_PHASES=1);
    KGW_LIB=<path> makes api.load_library() pick it up."""
    import os
    global LIB, OBJDIR
    tag = os.environ.get("KGW_BUILD_TAG")
    if tag:
        LIB = LIBDIR / f"libkgw_{tag}.so"
        OBJDIR = PKG / f"build_{tag}"
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: cannot build knockoffgwas_b200/lib/libkgw.so")
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    jobs = [(nvcc, s, OBJDIR / (s.stem + ".o"), verbose) for s in sources()]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        results = list(ex.map(_compile, jobs))
    for src, r in results:
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n" + r.stdout + r.stderr)
    cmd = [nvcc, "-shared", "-o", str(LIB)] + [str(j[2]) for j in jobs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
