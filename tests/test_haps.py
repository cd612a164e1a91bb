"""`.haps` ingest (text input): kgw_haps_read replaces the data loop of HapsReader::read
(snpknock2/src/haps_reader.cpp:93-124).  The restatement (oracle/kgw_oracle.py: haps_read / haps_tokenize) is pinned to
the matrix the unmodified reference binary read from a file with every kind of whitespace its tokenizer accepts
(tests/golden/haps_quirky.npz, made by tests/golden/make_golden_haps.py); the CUDA path must return the same bytes."""
from pathlib import Path

import numpy as np
import pytest

import kgw_oracle as ko

GOLD = Path(__file__).resolve().parent / "golden" / "haps_quirky.npz"


def _text(bits, rng, quirks=True, final_newline=True):
    rows = []
    for j in range(bits.shape[1]):
        toks = [b"1" if b else b"0" for b in bits[:, j]]
        if quirks:
            for t in rng.choice(len(toks), size=max(1, len(toks) // 40), replace=False):
                toks[t] = [b"10", b"x", b"1\r", b"01"][int(rng.integers(4))]
            seps = [[b" ", b"\t", b"   "][int(x)] for x in rng.choice(3, size=len(toks) - 1, p=[0.9, 0.05, 0.05])]
        else:
            seps = [b" "] * (len(toks) - 1)
        line = b"".join(a + b for a, b in zip(toks, seps + [b""]))
        rows.append(line + (b"\r" if quirks and j % 2 else b""))
    return b"\n".join(rows) + (b"\n" if final_newline else b"")


def test_oracle_matches_reference_reader():
    f = np.load(GOLD)
    H = ko.haps_read(f["text"].tobytes(), np.arange(int(f["n_samples"])), np.arange(int(f["p"])))
    assert np.array_equal(H, f["H"])


def test_tokenizer_rules():
    assert ko.haps_tokenize(b"  0\t1  10 1\r") == [b"0", b"1", b"10", b"1"]       # CR stripped from the last token only
    assert ko.haps_tokenize(b"1\r 0") == [b"1\r", b"0"]
    assert ko.haps_tokenize(b"0 1\r \t") == [b"0", b"1"]
    assert ko.haps_tokenize(b"") == []


@pytest.mark.gpu
def test_gpu_reads_quirky_file_like_the_reference():
    kg = pytest.importorskip("knockoffgwas_b200")
    f = np.load(GOLD)
    H, ms = kg.read_haps_text(f["text"], np.arange(int(f["n_samples"])), np.arange(int(f["p"])), return_ms=True)
    assert np.array_equal(H, f["H"]), "kgw_haps_read differs from the matrix the reference binary read"
    assert ms > 0


@pytest.mark.gpu
@pytest.mark.parametrize("n_file,p,quirks,final_newline", [(40, 33, True, True), (3000, 70, True, False), (5000, 130, False, True),
                                                           (2, 1, True, True)])
def test_gpu_matches_oracle_on_synthetic_text(n_file, p, quirks, final_newline):
    """Rows longer than one tile of the tokenizer (4096 bytes), sample subsets in permuted order, rows in permuted
    order, a file whose last row has no newline."""
    kg = pytest.importorskip("knockoffgwas_b200")
    rng = np.random.default_rng(n_file + p)
    bits = (rng.random((2 * n_file, p)) < 0.4).astype(np.uint8)
    text = _text(bits, rng, quirks, final_newline)
    for sidx, rows in ((np.arange(n_file), np.arange(p)),
                       (rng.permutation(n_file)[:max(1, n_file // 3)], rng.permutation(p)[:max(1, p // 2)])):
        assert np.array_equal(kg.read_haps_text(text, sidx, rows), ko.haps_read(text, sidx, rows))


@pytest.mark.gpu
def test_gpu_reader_mirror_and_errors(tmp_path):
    kg = pytest.importorskip("knockoffgwas_b200")
    rng = np.random.default_rng(3)
    bits = (rng.random((24, 50)) < 0.5).astype(np.uint8)
    (tmp_path / "x.haps").write_bytes(_text(bits, rng, quirks=False))
    (tmp_path / "x.legend").write_text("id position a0 a1\n" + "".join(f"rs{j} {100 + j} A G\n" for j in range(50)))
    (tmp_path / "x.sample").write_text("ID_1 ID_2 missing sex\n0 0 0 D\n" + "".join(f"s{i} s{i} 0 1\n" for i in range(12)))
    rd = kg.HapsReader(tmp_path / "x")
    H = rd.read(sample_ids=["s7", "s0"], snp_ids=["rs49", "rs3"])
    assert np.array_equal(ko.unpack_bits(H, 2), bits[[14, 15, 0, 1]][:, [49, 3]])
    assert np.array_equal(ko.unpack_bits(rd.read(), 50), bits)
    with pytest.raises(kg.KgwError):
        rd.read(sample_ids=["nobody"])
    short = b"0 1 0 1\n0 1\n"
    with pytest.raises(kg.KgwError):                        # second row has no token 2, 3
        kg.read_haps_text(short, np.arange(2), np.arange(2))
    assert kg.read_haps_text(short, np.arange(1), np.arange(2)).tolist() == [[0], [3]]
    with pytest.raises(kg.KgwError):                        # row outside the file
        kg.read_haps_text(short, np.arange(1), np.array([2]))
