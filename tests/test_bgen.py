"""BGEN ingest: kgw_bgen_read replaces the data loop of BgenReader::read_worker
(snpknock2/src/bgen_reader.cpp:139-261).  The numpy restatement (oracle/kgw_oracle.py: bgen_read) is pinned to the
matrix the unmodified reference binary read from its own toy .bgen (tests/golden/bgen_toy_chr22.npz, made by
tests/golden/make_golden_bgen.py); the CUDA path must return the same bytes."""
from pathlib import Path

import numpy as np
import pytest

import kgw_oracle as ko

GOLD = Path(__file__).resolve().parent / "golden" / "bgen_toy_chr22.npz"


def _synthetic(n=37, p=301, B=8, compression=1, seed=0, **kw):
    rng = np.random.default_rng(seed)
    bits = (rng.random((2 * n, p)) < 0.3).astype(np.uint8)
    rsids = [f"rs{j}" for j in range(p)]
    return bits, rsids, ko.bgen_write(bits, rsids, B=B, compression=compression, **kw)


def test_oracle_matches_reference_reader():
    f = np.load(GOLD)
    assert np.array_equal(ko.bgen_read(f["image"], f["sample_idx"], f["variant_of_col"]), f["H"])


@pytest.mark.parametrize("B,compression", [(8, 1), (16, 0), (5, 1), (1, 0), (32, 1)])
def test_oracle_roundtrip_synthetic(B, compression):
    bits, rsids, img = _synthetic(B=B, compression=compression)
    rng = np.random.default_rng(1)
    sidx = rng.permutation(bits.shape[0] // 2)[:20]
    vcol = rng.permutation(bits.shape[1])[:150]
    H = ko.bgen_read(np.frombuffer(img, np.uint8), sidx, vcol)
    rows = np.stack([2 * sidx, 2 * sidx + 1], axis=1).reshape(-1)
    assert np.array_equal(ko.unpack_bits(H, len(vcol)), bits[rows][:, vcol])


def test_oracle_rejects_what_the_reference_rejects():
    _, _, img = _synthetic(bad_prob=(3, 7))
    with pytest.raises(ValueError):
        ko.bgen_read(np.frombuffer(img, np.uint8), np.arange(37), np.arange(301))
    _, _, img = _synthetic(missing=(5, 11))
    with pytest.raises(ValueError):
        ko.bgen_read(np.frombuffer(img, np.uint8), np.arange(37), np.arange(301))
    # ... but not when the offending haplotype / variant is not requested
    ko.bgen_read(np.frombuffer(img, np.uint8), np.arange(5), np.arange(301))
    ko.bgen_read(np.frombuffer(img, np.uint8), np.arange(37), np.arange(11))


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_reads_reference_toy_bgen_like_the_reference():
    kg = pytest.importorskip("knockoffgwas_b200")
    f = np.load(GOLD)
    H, ms = kg.read_bgen_image(f["image"], f["sample_idx"], f["variant_of_col"], return_ms=True)
    assert np.array_equal(H, f["H"]), "kgw_bgen_read differs from the matrix the reference binary read"
    assert ms > 0


@pytest.mark.gpu
@pytest.mark.parametrize("B,compression,n,p", [(8, 1, 37, 301), (8, 0, 700, 130), (16, 0, 130, 129), (5, 1, 300, 31), (1, 0, 64, 1024), (32, 1, 9, 200)])
def test_gpu_matches_oracle_on_synthetic_images(B, compression, n, p):
    kg = pytest.importorskip("knockoffgwas_b200")
    bits, rsids, img = _synthetic(n=n, p=p, B=B, compression=compression, seed=B)
    image = np.frombuffer(img, np.uint8)
    rng = np.random.default_rng(2)
    for sidx, vcol in ((np.arange(n), np.arange(p)),
                       (rng.permutation(n)[:max(1, n // 2)], rng.permutation(p)[:max(1, 2 * p // 3)]),
                       (np.arange(n // 3, n - 1), rng.permutation(p))):      # consecutive samples from s0 > 0: staged fast path at B = 8
        if len(sidx) == 0:
            continue
        assert np.array_equal(kg.read_bgen_image(image, sidx, vcol), ko.bgen_read(image, sidx, vcol))


@pytest.mark.gpu
@pytest.mark.parametrize("host_inflate", [False, True])
@pytest.mark.parametrize("zlevel,n,p,B", [(0, 50, 40, 8), (1, 3, 70, 8), (9, 300, 33, 8), (6, 40000, 5, 8), (0, 40000, 3, 16), (6, 9000, 130, 1)])
def test_gpu_inflates_every_kind_of_deflate_block(zlevel, n, p, B, host_inflate, monkeypatch):
    """zlib blocks are inflated by bgen_inflate_kernel (one thread per variant): stored blocks (level 0; several of them
    when a block is longer than 65 535 bytes), fixed-Huffman blocks (tiny inputs), dynamic ones, matches further back than
    a few KB (n = 40 000 makes a 80 KB block).  KGW_BGEN_HOST_INFLATE=1 keeps libz on the host threads: same bytes."""
    kg = pytest.importorskip("knockoffgwas_b200")
    if host_inflate:
        monkeypatch.setenv("KGW_BGEN_HOST_INFLATE", "1")
    rng = np.random.default_rng(zlevel + n)
    bits = (rng.random((2 * n, p)) < rng.random(p)[None, :] * 0.5).astype(np.uint8)     # very compressible to noisy columns
    img = ko.bgen_write(bits, [f"rs{j}" for j in range(p)], B=B, compression=1, zlevel=zlevel)
    image = np.frombuffer(img, np.uint8)
    sidx = np.sort(rng.permutation(n)[:max(1, min(n, 500))])
    rows = np.stack([2 * sidx, 2 * sidx + 1], axis=1).reshape(-1)
    for vcol in (rng.permutation(p), rng.permutation(p)[:max(1, p // 4)]):      # the file's own span / a packed copy of the blocks
        H = kg.read_bgen_image(image, sidx, vcol)
        assert np.array_equal(ko.unpack_bits(H, len(vcol)), bits[rows][:, vcol])


@pytest.mark.gpu
@pytest.mark.parametrize("host_inflate", [False, True])
def test_gpu_rejects_corrupt_zlib_block(host_inflate, monkeypatch):
    """A damaged deflate stream or a wrong Adler-32 is an error (genfile::bgen throws on uncompress() != Z_OK)."""
    kg = pytest.importorskip("knockoffgwas_b200")
    if host_inflate:
        monkeypatch.setenv("KGW_BGEN_HOST_INFLATE", "1")
    bits, rsids, img = _synthetic(n=400, p=6, B=8, compression=1)
    for where in (-1, -40, -100):        # the last variant's block: checksum, then the stream itself
        bad = np.frombuffer(img, np.uint8).copy()
        bad[where] ^= 0x5a
        with pytest.raises(kg.KgwError):
            kg.read_bgen_image(bad, np.arange(400), np.arange(6))


@pytest.mark.gpu
def test_gpu_reader_mirror_and_errors(tmp_path):
    kg = pytest.importorskip("knockoffgwas_b200")
    bits, rsids, img = _synthetic(n=20, p=100)
    (tmp_path / "x.bgen").write_bytes(img)
    (tmp_path / "x.sample").write_text("ID_1 ID_2 missing sex\n0 0 0 D\n" + "".join(f"s{i} s{i} 0 1\n" for i in range(20)))
    rd = kg.BgenReader(tmp_path / "x")
    assert rd.n_samples == 20 and rd.variant_ids == rsids
    H = rd.read(sample_ids=["s3", "s1"], snp_ids=["rs5", "rs99", "rs0"])
    assert np.array_equal(ko.unpack_bits(H, 3), bits[[6, 7, 2, 3]][:, [5, 99, 0]])
    with pytest.raises(kg.KgwError):
        rd.read(snp_ids=["rs5", "nope"])
    _, _, bad = _synthetic(n=20, p=100, bad_prob=(3, 7))
    with pytest.raises(kg.KgwError, match="Incorrect input"):
        kg.read_bgen_image(np.frombuffer(bad, np.uint8), np.arange(20), np.arange(100))
    _, _, mis = _synthetic(n=20, p=100, missing=(5, 11))
    with pytest.raises(kg.KgwError, match="Incorrect input"):
        kg.read_bgen_image(np.frombuffer(mis, np.uint8), np.arange(20), np.arange(100))
    assert kg.read_bgen_image(np.frombuffer(mis, np.uint8), np.arange(5), np.arange(100)).shape == (10, 13)
    with pytest.raises(kg.KgwError):
        kg.read_bgen_image(np.frombuffer(img, np.uint8), np.array([20]), np.arange(100))   # sample outside the file


@pytest.mark.gpu
def test_gpu_rejects_truncated_probability_block():
    """A block whose stated length cannot hold 2 N values of B bits is refused (no out-of-bounds read on the device)."""
    import struct
    kg = pytest.importorskip("knockoffgwas_b200")
    bits, rsids, img = _synthetic(n=16, p=3, B=8, compression=0)
    b = bytearray(img)
    # first variant block: identifying data, then u32 length C, then C bytes; claim 16-bit probabilities instead of 8
    pos = 4 + 20
    for _ in range(3):
        (l,) = struct.unpack_from("<H", b, pos); pos += 2 + l
    pos += 6 + 5 + 5
    (c,) = struct.unpack_from("<I", b, pos)
    assert c == 10 + 16 + 32
    b[pos + 4 + 9 + 16] = 16
    with pytest.raises(kg.KgwError):
        kg.read_bgen_image(np.frombuffer(bytes(b), np.uint8), np.arange(16), np.arange(3))


def test_reader_mirror_id_bookkeeping_without_gpu(tmp_path):
    """The host part of api.BgenReader (variant ids from the container, sample ids from the .sample file, the reference's
    error for a variant that is not in the file) needs no device."""
    kg = pytest.importorskip("knockoffgwas_b200")
    bits, rsids, img = _synthetic(n=6, p=9)
    (tmp_path / "x.bgen").write_bytes(img)
    (tmp_path / "x.sample").write_text("ID_1 ID_2 missing sex\n0 0 0 D\n" + "".join(f"s{i} s{i} 0 1\n" for i in range(6)))
    rd = kg.BgenReader(tmp_path / "x")
    assert (rd.n_samples, rd.n_variants) == (6, 9) and rd.variant_ids == rsids and rd.sample_ids == [f"s{i}" for i in range(6)]
    with pytest.raises(kg.KgwError, match="requested variants were not found"):
        rd.read(snp_ids=["rs1", "nope"])
    with pytest.raises(kg.KgwError, match="not in the BGEN sample file"):
        rd.read(sample_ids=["s1", "nobody"])
    (tmp_path / "y.bgen").write_bytes(img)
    (tmp_path / "y.sample").write_text("ID_1 ID_2 missing sex\n0 0 0 D\ns0 s0 0 1\n")
    with pytest.raises(kg.KgwError, match="number of samples does not match"):
        kg.BgenReader(tmp_path / "y")
