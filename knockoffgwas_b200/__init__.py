"""knockoffgwas_b200 -- B200-native (sm_100a) replacement for the hot path of snpknock2's
haplotype knockoff sampler (msesia/knockoffgwas, snpknock2/src/knockoffs.cpp:311 `Knockoffs::generate`).

The product is the C-ABI shared library built from csrc/ (include/kgw.h).  This Python package is
a thin ctypes binding plus a host-side mirror of the reference's `Knockoffs` interface used by the
tests and the benchmark.  There is no CPU fallback: without the CUDA library / a B200 every compute
call raises.
"""
from .api import (BgenReader, HapsReader, KinshipMatrix, read_bgen_image, read_haps_text, KgwError, Knockoffs, Options, Plan, abi_version, device_count, partition_work, estimate_hmm, assign_references, bed_encode, combine_references_families, generate, generate_into, generate_strict, lib_path, load_library,
                  swap_flags,
                  release_workspace, uniform)

__all__ = ["BgenReader", "HapsReader", "KinshipMatrix", "read_bgen_image", "read_haps_text", "KgwError", "Knockoffs", "Options", "Plan", "abi_version", "device_count", "partition_work", "estimate_hmm", "assign_references", "combine_references_families", "bed_encode", "swap_flags", "generate", "generate_strict", "generate_into", "lib_path",
           "load_library", "release_workspace", "uniform"]
