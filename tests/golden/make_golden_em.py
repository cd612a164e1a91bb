#!/usr/bin/env python
"""Golden vector for --estimate-hmm: the UNMODIFIED reference binary run on its toy chr22 example
with `--estimate-hmm` instead of `--hmm-lambda` (Knockoffs::EM, knockoffs.cpp:706-960, --n_threads 1); the probe dumps
the HMM tables as Knockoffs::generate() then sees them (mutation rates after EM at full binary precision), and the
iteration table the reference prints is parsed from its stdout.  Needs /root/reference and oracle/_ref."""
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "oracle"))
import kgw_oracle as ko  # noqa: E402

DATA = Path("/root/reference/data")
PROBE = REPO / "oracle" / "_ref" / "snpknock2_probe"
OUT = Path(__file__).resolve().parent
with tempfile.TemporaryDirectory(prefix="kgw_em_") as tmp:
    tmp = Path(tmp)
    args = [str(PROBE), "--bgen", str(DATA / "haplotypes/example_chr22"), "--keep", str(DATA / "qc/samples_qc.txt"),
            "--extract", str(DATA / "qc/qc_chr22.txt"), "--map", str(DATA / "maps/genetic_map_chr22.txt"),
            "--part", str(DATA / "partitions/example_chr22.txt"), "--ibd", str(DATA / "ibd/example_chr22.txt"),
            "--K", "10", "--cluster_size_min", "1000", "--cluster_size_max", "10000", "--n_threads", "1", "--seed", "2020",
            "--hmm-rho", "1", "--estimate-hmm", "--windows", "0", "--resolution", "2", "--compute-references",
            "--generate-knockoffs", "--out", str(tmp / "ko")]
    env = dict(os.environ, KGW_PROBE_DIR=str(tmp / "dump"))
    r = subprocess.run(args, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    prob, _, _ = ko.load_probe_dump(tmp / "dump" / "call0")
    rows = re.findall(r"^\s*(\d+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s*$", r.stdout, flags=re.M)
    its = np.array([int(a) for a, _, _ in rows], dtype=np.int32)
    deltas = np.array([float(c) for _, _, c in rows], dtype=np.float64)
    conv = re.search(r"EM converged after (\d+) iterations", r.stdout)
    print(r.stdout[r.stdout.find("Estimating HMM"):][:900])
ref0, _ = None, None
base = np.load(OUT / "toy_chr22_K10_res2.npz")
assert np.array_equal(base["H"], prob.H) and np.array_equal(base["ref_local"], prob.ref_local) and np.array_equal(base["b"], prob.b)
np.savez_compressed(OUT / "em_toy_chr22_K10.npz", mu=prob.mu, b=prob.b, iterations=its, deltas=deltas,
                    converged_after=int(conv.group(1)) if conv else -1,
                    note=np.array("toy chr22, K=10, --hmm-rho 1 --estimate-hmm, --n_threads 1; mu as Knockoffs::generate() sees it after EM; "
                                  "H / references / b are those of toy_chr22_K10_res2.npz"))
print("wrote em_toy_chr22_K10.npz: iterations", its.tolist(), "deltas", deltas.tolist(), "converged after", conv.group(1) if conv else None,
      "mu range", float(prob.mu.min()), float(prob.mu.max()))
