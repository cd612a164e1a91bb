"""The data passes of the bifurcating k-means of `--compute-references` (SURVEY.md section 8f rank 2;
kinship_computer.cpp:239-250, 314-336, 368-405) on the device: kgw_kinship_distances / kgw_kinship_centroids against the
restatement in oracle/ -- bit for bit (the distances are sequential fp64 sums in locus order without fused multiply-adds,
the centroids exact counts divided once)."""
import numpy as np
import pytest

import kgw_oracle as ko
import knockoffgwas_b200 as kg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,p,n,seed", [(40, 33, 40, 1), (200, 257, 77, 2), (1000, 1500, 1000, 3), (64, 1, 10, 4)])
def test_kinship_passes_match_restatement(N, p, n, seed):
    rng = np.random.default_rng(seed)
    bits = (rng.random((N, p)) < rng.uniform(0.05, 0.6, size=p)).astype(np.uint8)
    H = ko.pack_bits(bits)
    members = rng.choice(N, size=n, replace=(n > N)).astype(np.int32)
    assign = (rng.random(n) < 0.4).astype(np.int32)
    assign[:2] = [0, 1]
    use = np.ones(p, dtype=np.uint8)
    use[p // 3: p // 2] = 0                      # a chromosome left out
    km = kg.KinshipMatrix(H, p)
    mu0, mu1 = km.centroids(members, assign)
    r0, r1 = ko.kinship_centroids(H, p, members, assign)
    assert np.array_equal(mu0, r0) and np.array_equal(mu1, r1)
    for u in (None, use):
        d0, d1 = km.distances(members, mu0, mu1, use=u)
        e0, e1 = ko.kinship_distances(H, p, members, mu0, mu1, use=u)
        assert np.array_equal(d0, e0) and np.array_equal(d1, e1)
    km.close()


def test_kinship_errors():
    H = np.zeros((4, 2), dtype=np.uint8)
    with pytest.raises(kg.KgwError):
        kg.KinshipMatrix(np.zeros((3, 2), dtype=np.uint8), 10)      # odd number of haplotypes
    km = kg.KinshipMatrix(H, 10)
    with pytest.raises(kg.KgwError):
        km.distances(np.array([7], dtype=np.int32), np.zeros(10), np.zeros(10))   # member out of range
    km.close()
