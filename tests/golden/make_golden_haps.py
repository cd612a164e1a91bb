#!/usr/bin/env python
"""Golden fixture for the `.haps` ingest: a small synthetic data set written in the reference's
input formats (oracle/refio.py), its .haps file rewritten with the whitespace the reference's tokenizer accepts
(runs of blanks, tabs, leading / trailing blanks, CR LF line ends) and with tokens that are not "0" / "1", then read
by the UNMODIFIED reference binary (HapsReader::read, haps_reader.cpp:37-132); the probe of oracle/ref_shims dumps
Haplotypes::H as Knockoffs::generate() sees it.  Needs /root/reference and oracle/_ref; the .npz output is committed."""
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(REPO), str(REPO / "oracle")]
import kgw_oracle as ko  # noqa: E402
import refio  # noqa: E402
from knockoffgwas_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent
sp = synth.make_problem(60, 97, 6, seed=31, n_founders=12, mean_dist_cm=0.05, block=60)
rng = np.random.default_rng(5)
with tempfile.TemporaryDirectory(prefix="kgw_haps_") as tmp:
    tmp = Path(tmp)
    files = refio.write_inputs(tmp / "in", sp.H, sp.p, sp.ref, sp.groups, sp.cm, sp.bp)
    rows = (tmp / "in" / "syn.haps").read_bytes().split(b"\n")[:-1]
    quirky = []
    for r, row in enumerate(rows):
        toks = row.split(b" ")
        for t in range(len(toks)):
            u = rng.random()
            if u < 0.03:
                toks[t] = [b"10", b"?", b"01", b"1.0", b"2", b"11"][int(rng.integers(6))]     # anything but "1" reads as 0
        seps = [[b" ", b"  ", b"\t", b" \t "][int(x)] for x in rng.choice(4, size=len(toks) - 1, p=[0.85, 0.05, 0.05, 0.05])]
        line = b"".join(a + b for a, b in zip(toks, seps + [b""]))
        if r % 5 == 1:
            line = b"  " + line
        if r % 7 == 2:
            line = line + b" \t"
        if r % 3 == 0:
            line = line + b"\r"
        quirky.append(line)
    text = b"\n".join(quirky) + b"\n"
    (tmp / "in" / "syn.haps").write_bytes(text)
    refio.run_reference(files, tmp / "out", rho=1.0, lam=1e-3, n_threads=1, dump_dir=tmp / "dump")
    dumps = sorted((tmp / "dump").iterdir())
    prob, _, _ = ko.load_probe_dump(dumps[0] if dumps[0].is_dir() else tmp / "dump")
H = prob.H
sidx, rows_idx = np.arange(sp.N // 2, dtype=np.int32), np.arange(sp.p, dtype=np.int32)
assert H.shape == (sp.N, (sp.p + 7) // 8)
mine = ko.haps_read(text, sidx, rows_idx)
assert np.array_equal(mine, H), "oracle restatement != reference reader"
np.savez_compressed(OUT / "haps_quirky.npz", text=np.frombuffer(text, np.uint8), H=H, n_samples=sp.N // 2, p=sp.p)
print("wrote", OUT / "haps_quirky.npz", len(text), "bytes of text,", H.shape, "differs from the clean matrix in",
      int(np.unpackbits(H ^ sp.H).sum()), "cells")
