#!/usr/bin/env python
"""Golden fixture for the BGEN ingest.

`H` in tests/golden/toy_chr22_K10_res0.npz is Haplotypes::H exactly as the UNMODIFIED reference binary read it
from data/haplotypes/example_chr22.bgen with --keep data/qc/samples_qc.txt --extract data/qc/qc_chr22.txt
(BgenReader::read, bgen_reader.cpp:66-261; dumped by the probe of oracle/ref_shims).  This script stores the
inputs of that read next to it: the .bgen image (the reference's own toy data), the kept sample rows and the
kept variant positions, following Metadata::apply_filters (metadata.cpp:358-392: file order is kept).
Needs /root/reference; the .npz output is committed."""
from pathlib import Path

import numpy as np

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent
image = np.fromfile(REF / "haplotypes/example_chr22.bgen", dtype=np.uint8)
sample_rows = [ln.split()[0] for ln in (REF / "haplotypes/example_chr22.sample").read_text().splitlines()[2:]]
keep = {ln.split()[0] for ln in (REF / "qc/samples_qc.txt").read_text().splitlines() if ln.strip()}
extract = {ln.strip() for ln in (REF / "qc/qc_chr22.txt").read_text().splitlines() if ln.strip()}
import sys
sys.path.insert(0, str(OUT.parents[1] / "oracle"))
import kgw_oracle as ko  # noqa: E402
variants, n = ko.bgen_variants(image.tobytes())
sample_idx = np.array([i for i, s in enumerate(sample_rows) if s in keep], dtype=np.int32)
variant_of_col = np.array([v for v, (rsid, _) in enumerate(variants) if rsid in extract], dtype=np.int32)
H = np.load(OUT / "toy_chr22_K10_res0.npz")["H"]
assert H.shape == (2 * len(sample_idx), (len(variant_of_col) + 7) // 8), (H.shape, len(sample_idx), len(variant_of_col))
assert np.array_equal(ko.bgen_read(image, sample_idx, variant_of_col), H), "oracle restatement != reference reader"
np.savez_compressed(OUT / "bgen_toy_chr22.npz", image=image, sample_idx=sample_idx, variant_of_col=variant_of_col, H=H,
                    sample_ids=np.array(sample_rows), rsids=np.array([r for r, _ in variants]))
print("wrote", OUT / "bgen_toy_chr22.npz", image.nbytes, "bytes of BGEN,", H.shape)
