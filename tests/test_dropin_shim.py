"""The drop-in, end to end: oracle/_ref/snpknock2_kgw is the UNMODIFIED reference (CLI, BGEN reader, metadata, kinship
clusters and reference selection, IBD tidying, partitions, PLINK writer) linked with integration/snpknock2_kgw_shim.cpp
in front of Knockoffs::generate (GNU ld --wrap) and with libkgw.so (integration/build_dropin.sh; the binary exists
only where /root/reference was available to build it and travels to the GPU box with oracle/_ref/).

The shim also replaces BgenReader::read (-> kgw_bgen_read) and plink::write_binary (-> kgw_bed_encode) unless KGW_SHIM_IO=0.

CPU: the kgw_problem the shim assembles from the reference's objects equals, array for array, what the probe
recorded from the reference itself (tests/golden/toy_chr22_K10_res2.npz), and without a GPU the binary stops with
the library's "no CPU fallback" error instead of producing output.
GPU: the `.bed` the binary writes is the reference's own writer applied to (H, Hk) with Hk = this library's knockoffs
for that problem -- byte for byte."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from kgw_testutil import load_golden

REPO = Path(__file__).resolve().parents[1]
BINARY = REPO / "oracle" / "_ref" / "snpknock2_kgw"
GOLD = Path(__file__).resolve().parent / "golden"

needs_binary = pytest.mark.skipif(not BINARY.exists(), reason="oracle/_ref/snpknock2_kgw not built (needs /root/reference)")


def _lay_out_inputs(d: Path) -> dict:
    f = np.load(GOLD / "dropin_chr22_inputs.npz")
    (d / "example_chr22.bgen").write_bytes(np.load(GOLD / "bgen_toy_chr22.npz")["image"].tobytes())
    (d / "example_chr22.sample").write_bytes(f["sample"].tobytes())
    (d / "example_chr22.bim").write_bytes(f["bim"].tobytes())
    for k in ("keep", "extract", "map", "part", "ibd"):
        (d / f"{k}.txt").write_bytes(f[k].tobytes())
    return {k: str(d / f"{k}.txt") for k in ("keep", "extract", "map", "part", "ibd")} | {"bgen": str(d / "example_chr22")}


def _run(tmp_path: Path, dump: bool, shim_io: bool = True, resolution="2", devices=None, cluster_sizes=("1000", "10000"), audit=0):
    (tmp_path / "in").mkdir(exist_ok=True)
    (tmp_path / "out").mkdir(exist_ok=True)
    fl = _lay_out_inputs(tmp_path / "in")
    args = [str(BINARY), "--bgen", fl["bgen"], "--keep", fl["keep"], "--extract", fl["extract"], "--map", fl["map"],
            "--part", fl["part"], "--ibd", fl["ibd"], "--K", "10", "--cluster_size_min", cluster_sizes[0], "--cluster_size_max", cluster_sizes[1],
            "--n_threads", "1", "--seed", "2020", "--hmm-rho", "1", "--hmm-lambda", "1e-3", "--windows", "0",
            "--compute-references", "--generate-knockoffs", "--out", str(tmp_path / "out" / "ko")]
    if resolution is not None:
        args += ["--resolution", resolution]
    env = dict(os.environ)
    if devices is not None:
        env["KGW_DEVICES"] = devices
    env["KGW_SHIM_IO"] = "1" if shim_io else "0"          # 1: BgenReader::read and plink::write_binary also go through libkgw
    if dump:
        env["KGW_SHIM_DUMP"] = str(tmp_path / "dump")
    if audit:
        env["KGW_AUDIT_HAPS"] = str(audit)
    return subprocess.run(args, env=env, capture_output=True, text=True, cwd=str(tmp_path), timeout=900)


def _cuda_device_present() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@needs_binary
def test_shim_assembles_the_problem_the_reference_hands_to_generate(tmp_path):
    r = _run(tmp_path, dump=True, shim_io=False)         # the reference reads the BGEN itself: possible without a GPU
    prob, _ = load_golden("toy_chr22_K10_res2")
    d = tmp_path / "dump" / "call0"
    rd = lambda name, t: np.fromfile(d / name, dtype=t)
    N, p, K, W, rb, seed = (int(x) for x in rd("dims.i64", np.int64))
    assert (N, p, K, W, seed) == (prob.N, prob.p, prob.K, 1, prob.seed)
    assert np.array_equal(rd("H.u8", np.uint8).reshape(N, rb), prob.H)
    assert np.array_equal(rd("ref_local.i32", np.int32).reshape(N, W, K), prob.ref_local)
    assert np.array_equal(rd("ref_global.i32", np.int32).reshape(N, K), prob.ref_global)
    assert np.array_equal(rd("b.f64", np.float64), prob.b) and np.array_equal(rd("mu.f64", np.float64), prob.mu)
    assert np.array_equal(rd("groups.i32", np.int32), prob.groups)
    assert np.array_equal(rd("unrelated.i32", np.int32), prob.unrelated)
    fam_off, mem = rd("fam_off.i32", np.int32), rd("fam_members.i32", np.int32)
    seg_off, jmin, jmax = rd("seg_off.i32", np.int32), rd("seg_jmin.i32", np.int32), rd("seg_jmax.i32", np.int32)
    ids_off, ids = rd("seg_ids_off.i32", np.int32), rd("seg_ids.i32", np.int32)
    got = {}
    for fi in range(len(fam_off) - 1):
        m = tuple(mem[fam_off[fi]:fam_off[fi + 1]].tolist())
        got[m] = [(int(jmin[s]), int(jmax[s]), tuple(ids[ids_off[s]:ids_off[s + 1]].tolist())) for s in range(seg_off[fi], seg_off[fi + 1])]
    want = {tuple(int(x) for x in m): [(int(s[0]), int(s[1]), tuple(int(x) for x in s[4])) for s in segs] for m, segs in prob.fam_clusters}
    assert got == want                                   # families and their tidied IBD segments, in the reference's order
    if not _cuda_device_present():                        # no silent fallback: the reference's error convention, exit(-1)
        assert r.returncode != 0
        assert "no CPU fallback" in r.stderr
        assert not list((tmp_path / "out").glob("*.bed"))
        r2 = _run(tmp_path, dump=False, shim_io=True)    # with the reader replaced too the stop comes at the BGEN ingest
        assert r2.returncode != 0 and "no CPU fallback" in r2.stderr


@pytest.mark.gpu
@needs_binary
def test_dropin_binary_writes_the_librarys_knockoffs(tmp_path):
    kg = pytest.importorskip("knockoffgwas_b200")
    r = _run(tmp_path, dump=False, shim_io=True)          # BGEN ingest, sampler and .bed body all on the device
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    beds = sorted((tmp_path / "out").glob("*_res2.bed"))
    assert len(beds) == 1, [b.name for b in (tmp_path / "out").iterdir()]
    bed = np.fromfile(beds[0], dtype=np.uint8)
    assert bytes(bed[:3]) == b"\x6c\x1b\x01"
    bim = (tmp_path / "out" / beds[0].name.replace(".bed", ".bim")).read_bytes()
    clust = {f.name: f.read_bytes() for f in sorted((tmp_path / "out").glob("*_clust.txt"))}   # k-means with its data passes on the device
    assert clust and "GPU" in r.stderr, "the kinship clusters were not computed through the library"
    r0 = _run(tmp_path, dump=False, shim_io=False)        # the reference's own reader, k-means and writer around the device sampler
    assert r0.returncode == 0, r0.stdout[-2000:] + r0.stderr[-2000:]
    assert {f.name: f.read_bytes() for f in sorted((tmp_path / "out").glob("*_clust.txt"))} == clust, \
        "the kinship clusters of the device k-means differ from the reference's"
    assert np.array_equal(np.fromfile(beds[0], dtype=np.uint8), bed), "the two forms of the drop-in write different .bed files"
    assert (tmp_path / "out" / beds[0].name.replace(".bed", ".bim")).read_bytes() == bim
    prob, _ = load_golden("toy_chr22_K10_res2")
    fams = [(m, [(s[0], s[1], s[4]) for s in segs]) for m, segs in prob.fam_clusters]
    hk, _ = kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, unrelated=prob.unrelated,
                        ref_global=prob.ref_global, seed=prob.seed, families=fams)
    want = kg.bed_encode(prob.H, hk, prob.p)              # == the reference's writeBED on (H, Hk), tests/test_bed.py
    assert np.array_equal(bed[3:].reshape(want.shape), want), "the .bed of the drop-in binary is not (H, this library's Hk)"


@pytest.mark.gpu
@needs_binary
def test_dropin_binary_audits_its_knockoffs_in_reference_order(tmp_path):
    """KGW_AUDIT_HAPS=n: after every generate() the drop-in samples n unrelated haplotypes again with
    kgw_generate_strict (the reference's summation order on the device) and stops if a row differs."""
    r = _run(tmp_path, dump=False, resolution=None, audit=48)      # all seven resolutions
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("[kgw] audit:")]
    assert len(lines) == 7 and all(l.endswith("48 haplotypes sampled again in the reference's summation order, 0 differ") for l in lines), lines


@pytest.mark.gpu
@needs_binary
def test_dropin_binary_all_resolutions_one_plan_sharded(tmp_path):
    """The reference's loop over every partition of the --part file (main.cpp:159-181) through the drop-in: the shim keeps
    ONE plan for the Knockoffs object and only replaces the partition between resolutions (kgw_plan_set_groups), and it
    shards the work list over the devices of KGW_DEVICES (here the same GPU twice).  Every `_res<r>.bed` whose inputs
    exist as a fixture must be (H, this library's Hk for that partition)."""
    kg = pytest.importorskip("knockoffgwas_b200")
    r = _run(tmp_path, dump=False, shim_io=True, resolution=None, devices="0,0")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    checked = 0
    for res in (0, 2, 4, 6):
        prob, _ = load_golden(f"toy_chr22_K10_res{res}")
        beds = sorted((tmp_path / "out").glob(f"*_res{res}.bed"))
        assert len(beds) == 1, [b.name for b in (tmp_path / "out").iterdir()]
        bed = np.fromfile(beds[0], dtype=np.uint8)
        fams = [(m, [(s[0], s[1], s[4]) for s in segs]) for m, segs in prob.fam_clusters]
        hk, _ = kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups, unrelated=prob.unrelated,
                            ref_global=prob.ref_global, seed=prob.seed, families=fams)
        want = kg.bed_encode(prob.H, hk, prob.p)
        assert np.array_equal(bed[3:].reshape(want.shape), want), f"resolution {res}"
        checked += 1
    assert checked == 4


@pytest.mark.gpu
@needs_binary
def test_dropin_binary_haps_input(tmp_path):
    """The same drop-in fed `--haps` text input (HapsReader::read replaced by kgw_haps_read in the shim) and preloaded
    references: the `.bed` must be (H, this library's Hk) for the synthetic population that was written to the text files."""
    kg = pytest.importorskip("knockoffgwas_b200")
    import sys
    sys.path.insert(0, str(REPO / "oracle"))
    import refio
    from knockoffgwas_b200 import synth
    sp = synth.make_problem_fast(240, 500, 10, seed=33, n_founders=30, mean_group=3.0, mean_dist_cm=0.02)
    files = refio.write_inputs(tmp_path / "in", sp.H, sp.p, sp.ref, sp.groups, sp.cm, sp.bp)
    out = tmp_path / "out"
    out.mkdir()
    args = [str(BINARY), "--haps", files["prefix"], "--map", files["map"], "--part", files["part"], "--ref", files["ref"],
            "--lref", files["lref"], "--hmm-rho", "1.0", "--hmm-lambda", "0.001", "--windows", "0", "--resolution", "0",
            "--seed", str(sp.seed), "--n_threads", "1", "--generate-knockoffs", "--out", str(out / "ko")]
    got = {}
    for io in ("1", "0"):                                 # shim reader / writer, then the reference's own around the device sampler
        env = dict(os.environ, KGW_SHIM_IO=io)
        r = subprocess.run(args, env=env, capture_output=True, text=True, cwd=str(out), timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        beds = sorted(out.glob("*_res0.bed"))
        assert len(beds) == 1
        got[io] = np.fromfile(beds[0], dtype=np.uint8)
    assert np.array_equal(got["1"], got["0"]), "the shim's HAPS reader / .bed writer and the reference's give different files"
    # the reference's init_hmm(rho, lambda) on the spline-interpolated map: b agrees with the synthetic table to ~1e-15, which
    # cannot move a draw here; the knockoffs must be the library's for (H, references, groups) as they were written
    hk, _ = kg.generate(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, seed=sp.seed)
    want = kg.bed_encode(sp.H, hk, sp.p)
    assert np.array_equal(got["1"][3:].reshape(want.shape), want)


@pytest.mark.gpu
@needs_binary
def test_dropin_binary_kmeans_on_device(tmp_path):
    """`--compute-references` with cluster sizes that make the bifurcating k-means split the 2000 toy haplotypes several
    times (kinship_computer.cpp:511-691): with the data passes of the k-means on the device (kgw_kinship_*) the drop-in
    writes the clusters, the references and the knockoffs the reference's own k-means leads to."""
    pytest.importorskip("knockoffgwas_b200")
    outs = {}
    for io in (True, False):
        r = _run(tmp_path, dump=False, shim_io=io, cluster_sizes=("100", "300"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        if io:
            assert "Bifurcating K-means on the GPU" in r.stderr
        files = {}
        for pat in ("*_clust.txt", "*_lref.txt", "*_ref.txt", "*_res2.bed"):
            got = sorted((tmp_path / "out").glob(pat))
            assert got, pat
            files.update({f.name: f.read_bytes() for f in got})
        outs[io] = files
    clust = next(v for k, v in outs[True].items() if k.endswith("_clust.txt"))
    assert len(set(clust.split())) >= 4, "the k-means did not split anything: the test exercises nothing"
    assert outs[True].keys() == outs[False].keys()
    for k in outs[True]:
        assert outs[True][k] == outs[False][k], k
