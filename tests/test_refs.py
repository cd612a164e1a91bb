"""Reference selection: KinshipComputer::assign_references
(kinship_computer.cpp:174-233, 765-931) given the kinship clusters.

CPU: the numpy restatement in oracle/ is pinned to the `_lref.txt` / `_ref.txt` the unmodified reference wrote
(tests/golden/refs_*.npz, made by make_golden_refs.py).  GPU: the device K-NN reproduces the restatement's
lists exactly, order included."""
import numpy as np
import pytest

import kgw_oracle as ko
import knockoffgwas_b200 as kg
from kgw_testutil import GOLDEN

NAMES = ["refs_toy_chr22_K10_clusters"]


def _load(name):
    f = np.load(GOLDEN / f"{name}.npz")
    clusters, t = [], 0
    for n in f["clust_sizes"]:
        clusters.append(f["clust_members"][t:t + n])
        t += n
    fams, t = [], 0
    for n in f["fam_sizes"]:
        fams.append(f["fam_members"][t:t + n])
        t += n
    return f, clusters, fams


@pytest.mark.parametrize("name", NAMES)
def test_oracle_references_match_reference_files(name):
    f, clusters, fams = _load(name)
    p, K = int(f["p"]), int(f["K"])
    nn = ko.find_neighbors(f["H"], p, K, clusters, ko.family_labels(fams))
    final = ko.combine_references_families(nn, fams)
    assert np.array_equal(final, f["lref"][:, 0, :])      # <out>_lref.txt
    assert np.array_equal(final, f["gref"])               # <out>_ref.txt (one window: same lists, kinship.cpp:264-267)


def test_host_combine_matches_oracle():
    """api.combine_references_families (the host step after the device K-NN) == the oracle's restatement."""
    f, clusters, fams = _load(NAMES[0])
    nn = ko.find_neighbors(f["H"], int(f["p"]), int(f["K"]), clusters, ko.family_labels(fams))
    assert np.array_equal(kg.combine_references_families(nn, fams), f["gref"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_references_match_reference_files(name):
    f, clusters, fams = _load(name)
    p, K = int(f["p"]), int(f["K"])
    lab = ko.family_labels(fams)
    nn, ms = kg.assign_references(f["H"], p, K, clusters, lab)
    assert ms >= 0.0
    assert np.array_equal(nn, ko.find_neighbors(f["H"], p, K, clusters, lab))          # order included
    assert np.array_equal(kg.combine_references_families(nn, fams), f["gref"])         # == <out>_ref.txt


@pytest.mark.gpu
@pytest.mark.parametrize("N,p,K,sizes,comp", [(300, 64, 10, [300], 1), (700, 333, 50, [130, 257, 313], 10),
                                              (64, 20, 40, [31, 33], 3), (1500, 2000, 100, [700, 800], 10),
                                              (40, 90, 50, [9, 31], 10)])
def test_device_references_ties_and_ragged_clusters(N, p, K, sizes, comp):
    """Short thinned rows make many equal distances: ties must come out in cluster order; clusters smaller than
    K+1 give -1 padding; IBD relatives are pushed behind everyone else."""
    rng = np.random.default_rng(N + p + K)
    H = ko.pack_bits((rng.random((N, p)) < 0.3).astype(np.uint8))
    perm = rng.permutation(N)
    clusters, t = [], 0
    for n in sizes:
        clusters.append(perm[t:t + n].astype(np.int32))
        t += n
    lab = np.full(N, -1, dtype=np.int32)
    for a in range(0, min(N, 40), 4):                      # a few families of two haplotypes
        lab[perm[a]] = lab[perm[a + 1]] = int(perm[a])
    nn, _ = kg.assign_references(H, p, K, clusters, lab, compression=comp)
    assert np.array_equal(nn, ko.find_neighbors(H, p, K, clusters, lab, compression=comp))
