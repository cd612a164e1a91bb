#!/usr/bin/env python
"""Golden vector for the .bed body: runs the UNMODIFIED reference (oracle/_ref/snpknock2_probe) on the toy
chr22 data at resolution 2 and stores H, the reference's Hk (probe dump), the bytes of the
`<out>_res2.bed` it wrote (plink_interface.cpp:67-103) and the column order recorded in its `.bim`
(`.k` suffix first = swapped, :117-146).  Output: tests/golden/bed_toy_chr22_K10_res2.npz (committed).
Needs /root/reference and oracle/_ref: build container only."""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1] / "oracle"))
import kgw_oracle as ko  # noqa: E402
from make_golden import run_ref  # noqa: E402


def main():
    tmp = Path(tempfile.mkdtemp(prefix="kgw_golden_bed_"))
    try:
        out = run_ref(tmp, 22, 10, ["--hmm-rho", "1", "--hmm-lambda", "1e-3", "--windows", "0", "--resolution", "2"],
                      "chr22_K10_bed", None)
        prob, Hk, _ = ko.load_probe_dump(out / "dump" / "call0")
        bed = np.fromfile(out / "ko_chr22_res2.bed", dtype=np.uint8)
        assert bytes(bed[:3]) == b"\x6c\x1b\x01"
        names = [ln.split("\t")[1] for ln in open(out / "ko_chr22_res2.bim")]
        assert len(names) == 2 * prob.p
        swap = np.array([1 if names[2 * j].endswith(".k") else 0 for j in range(prob.p)], dtype=np.uint8)
        nb = (prob.N // 2 + 3) // 4
        body = bed[3:].reshape(2 * prob.p, nb)
        np.savez_compressed(HERE / "bed_toy_chr22_K10_res2.npz", p=prob.p, H=prob.H, Hk=Hk, swap=swap, bed_body=body,
                            note=np.array("toy chr22, K=10, resolution 2, --n_threads 1; bytes of ko_chr22_res2.bed after the magic"))
        print("bed fixture:", prob.N, prob.p, body.shape, "swapped:", int(swap.sum()))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
