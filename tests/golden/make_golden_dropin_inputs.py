#!/usr/bin/env python
"""Inputs of the drop-in test (tests/test_dropin_shim.py): the small text files of the reference's toy chr22 example
(data/haplotypes/example_chr22.sample|.bim, data/qc, data/maps, data/partitions, data/ibd -- the arguments of
examples/example_1_generate_knockoffs.sh) as raw bytes, so that the test can lay them out again on the GPU box,
where /root/reference does not exist.  The .bgen image itself is already in tests/golden/bgen_toy_chr22.npz.
Needs /root/reference; the .npz output is committed."""
from pathlib import Path

import numpy as np

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent
files = {
    "sample": REF / "haplotypes/example_chr22.sample",
    "bim": REF / "haplotypes/example_chr22.bim",
    "keep": REF / "qc/samples_qc.txt",
    "extract": REF / "qc/qc_chr22.txt",
    "map": REF / "maps/genetic_map_chr22.txt",
    "part": REF / "partitions/example_chr22.txt",
    "ibd": REF / "ibd/example_chr22.txt",
}
np.savez_compressed(OUT / "dropin_chr22_inputs.npz", **{k: np.frombuffer(p.read_bytes(), np.uint8) for k, p in files.items()})
print("wrote", OUT / "dropin_chr22_inputs.npz", {k: p.stat().st_size for k, p in files.items()})
