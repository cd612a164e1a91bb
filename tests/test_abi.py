"""CPU-side checks of the drop-in boundary: libkgw.so loads, exports every symbol include/kgw.h
declares, the ctypes mirrors have the C layout, and -- with no CUDA device -- every compute entry
point fails loudly (there is no CPU fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import knockoffgwas_b200 as kg
from knockoffgwas_b200 import api
from kgw_testutil import synthetic_problem

REPO = Path(__file__).resolve().parents[1]
HEADER = REPO / "include" / "kgw.h"


@pytest.fixture(scope="module")
def lib():
    from knockoffgwas_b200 import build
    build.build()
    return kg.load_library()


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(kgw_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert {"kgw_generate", "kgw_plan_create", "kgw_plan_run", "kgw_plan_download", "kgw_uniform",
            "kgw_last_error", "kgw_abi_version", "kgw_plan_set_groups", "kgw_plan_info"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/kgw.h but not exported by libkgw.so"
    out = subprocess.run(["nm", "-D", "--defined-only", str(kg.lib_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (kgw_[a-z_0-9]+)", out))
    assert set(names) <= exported
    assert lib.kgw_abi_version() == api.KGW_ABI_VERSION


def test_ctypes_layout_matches_c_header(tmp_path):
    src = tmp_path / "layout.c"
    fields = {
        "kgw_problem": [f for f, _ in api._Problem._fields_],
        "kgw_options": [f for f, _ in api._Options._fields_],
        "kgw_debug": [f for f, _ in api._Debug._fields_],
        "kgw_plan_info_t": [f for f, _ in api._PlanInfo._fields_],
    }
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s, fl in fields.items():
        body.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fl:
            body.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    body.append("return 0;}")
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirrors = {"kgw_problem": api._Problem, "kgw_options": api._Options, "kgw_debug": api._Debug,
               "kgw_plan_info_t": api._PlanInfo}
    for s, cls in mirrors.items():
        assert int(got[s]) == C.sizeof(cls), s
        for f, _ in cls._fields_:
            assert int(got[f"{s}.{f}"]) == getattr(cls, f).offset, f"{s}.{f}"


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a host without a CUDA device")
def test_no_cpu_fallback(lib):
    prob = synthetic_problem(N=16, p=40, K=4, seed=2)
    with pytest.raises(kg.KgwError) as ei:
        kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)
    with pytest.raises(kg.KgwError):
        kg.Plan(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups)
    with pytest.raises(kg.KgwError) as ei:      # the audit mode runs on the device as well
        kg.generate_strict(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_argument_validation_precedes_device_probe(lib):
    prob = synthetic_problem(N=16, p=40, K=4, seed=2)
    with pytest.raises(kg.KgwError) as ei:  # row_bytes < ceil(p/8)
        kg.generate(prob.H[:, :3], prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups)
    assert ei.value.code == -1
    big = synthetic_problem(N=300, p=16, K=257, seed=2)
    with pytest.raises(kg.KgwError) as ei:  # the stated bound on K (include/kgw.h)
        kg.generate(big.H, big.p, big.K, big.ref_local, big.b, big.mu, big.groups)
    assert ei.value.code == -2 and "K > 256" in str(ei.value)
    bad_ref = prob.ref_local.copy()
    bad_ref[3, 0, 1] = 99
    with pytest.raises(kg.KgwError) as ei:
        kg.generate(prob.H, prob.p, prob.K, bad_ref, prob.b, prob.mu, prob.groups)
    assert ei.value.code == -1 and "reference index" in str(ei.value)


def test_host_replay_of_generator_matches_oracle_without_gpu(lib):
    import kgw_oracle as ko
    for (hap, s, j) in [(0, 0, 0), (5, 1, 3), (1999, 2, 999), (699999, 1, 59999), (12, 2 | (3 << 2), 77)]:
        assert kg.uniform(2020, hap, s, j) == ko.philox_uniform(2020, hap, s, j)
    u = np.array([kg.uniform(7, 1, 0, j) for j in range(4096)])
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02


def test_host_mirror_of_knockoffs_interface():
    """Knockoffs mirror: init_hmm / load_hmm / set_references follow knockoffs.cpp:1508-1525, :1593-1610, :109-148."""
    prob = synthetic_problem(N=12, p=30, K=3, seed=4)
    d = np.full(30, 0.01)
    k = kg.Knockoffs(prob.H, 30, 3, genetic_dist=d)
    assert k.b[0] == 0.0 and np.allclose(k.b[1:], np.exp(-0.01)) and (k.mut_rates == 1e-4).all()
    k.init_hmm(3, 2.0, 1e-3)
    assert np.allclose(k.b[1:], np.exp(-0.02)) and (k.mut_rates == 1e-3).all()
    k.load_hmm(np.full(30, 0.5), np.array([1e-9] * 15 + [1.0] * 15))
    assert np.allclose(k.b, np.exp(-0.5)) and k.mut_rates.min() == 1e-6 and k.mut_rates.max() == 1e-3
    k.set_references(prob.ref_local[:, :, ::-1])
    assert (np.diff(k.references_local, axis=2) > 0).all()
    assert k.num_groups() == 30
