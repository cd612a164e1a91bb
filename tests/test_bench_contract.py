"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the keys
the driver reads (the unmodified reference binary on the host cores, oracle/_ref; skipped where it was not built), and
the B200 arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def _cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(not (REPO / "oracle" / "_ref" / "snpknock2_probe").exists(), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-threads", "2", "--snps", "600", "--haps", "600"], capture_output=True, text=True, timeout=600, cwd=str(REPO))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "knockoff_haplotype_snps_per_sec_K50" and d["unit"] == "hap-SNPs/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.skipif(_cuda(), reason="a CUDA device is present")
def test_b200_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=str(REPO))
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert r.stdout.strip() == ""
