"""Goodness-of-fit diagnostics of module_2b (knockoffgwas/module_2b_knockoffs_gof.sh:57-72 runs plink
--freq / --r / --r2 on the <out>_res<r>.bed written after Knockoffs::generate(); utils/knockoffs_gof.R:17-113
turns them into four panels) restated in numpy on the toy fixtures:

  1. MAF(X_j) against MAF(Xk_j)                                   knockoffs_gof.R:20-38
  2. |corr(X_j, X_k)| against |corr(Xk_j, Xk_k)|, groups <= 10 apart  :58-86
  3. |corr(X_j, X_k)| against |corr(X_j, Xk_k)|, different groups     :66-98
  4. corr(X_j, Xk_j)^2  (self-similarity)                             :99-103

Genotypes are sums of the two haplotypes of a sample, as in the BED (plink_interface.cpp:10-35).  The
production generator (counter-based Philox) draws a different stream than the reference's taus88, so the
two knockoff matrices are different samples of the same distribution: the test checks that the summary of
every panel agrees between the device's knockoffs and the reference binary's (committed in the fixture)
within the sampling noise, and that both satisfy the exchangeability identities the panels visualise.
(Under injected uniforms the matrices are bit-identical, tests/test_gpu_parity.py, and so are these.)
"""
import numpy as np
import pytest

import kgw_oracle as ko
import knockoffgwas_b200 as kg
from kgw_testutil import GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu

MAX_SNP_DIST = 12   # SNP pairs examined per SNP (the R script keeps pairs at most 10 groups apart)


def _geno(Hpacked, p, rows):
    bits = ko.unpack_bits(Hpacked, p)[rows].astype(np.float64)
    assert len(rows) % 2 == 0
    return bits[0::2] + bits[1::2]   # samples x SNPs, values 0/1/2


def _corr_cols(A, B):
    """column-wise Pearson correlation of two (n x m) matrices; 0 where a column is constant."""
    A = A - A.mean(0)
    B = B - B.mean(0)
    den = np.sqrt((A * A).sum(0) * (B * B).sum(0))
    num = (A * B).sum(0)
    return np.where(den > 0, num / np.where(den > 0, den, 1.0), 0.0)


def gof_panels(X, Xk, groups):
    """The four panels as arrays (see the module docstring)."""
    p = X.shape[1]
    maf = lambda G: np.minimum(G.mean(0) / 2.0, 1.0 - G.mean(0) / 2.0)
    xx, kk, xk = [], [], []
    for d in range(1, MAX_SNP_DIST + 1):
        a, b = np.arange(0, p - d), np.arange(d, p)
        gd = np.abs(groups[b] - groups[a])
        keep = gd <= 10
        rxx = _corr_cols(X[:, a], X[:, b])
        rkk = _corr_cols(Xk[:, a], Xk[:, b])
        rxk = _corr_cols(X[:, a], Xk[:, b])
        xx.append(np.stack([np.abs(rxx[keep]), gd[keep]]))
        kk.append(np.abs(rkk[keep]))
        xk.append(np.abs(rxk[keep]))
    xx = np.concatenate(xx, axis=1)
    return dict(maf_x=maf(X), maf_xk=maf(Xk), r_xx=xx[0], gdist=xx[1], r_kk=np.concatenate(kk), r_xk=np.concatenate(xk),
                self_r2=_corr_cols(X, Xk) ** 2)


def summarise(P):
    diff = P["gdist"] > 0   # panel 3 keeps pairs from different groups only
    strong = P["r_xx"] > 0.2
    return dict(
        maf_bias=float(np.mean(P["maf_xk"] - P["maf_x"])),
        maf_rmse=float(np.sqrt(np.mean((P["maf_xk"] - P["maf_x"]) ** 2))),
        kk_slope=float(np.sum(P["r_kk"][strong] * P["r_xx"][strong]) / np.sum(P["r_xx"][strong] ** 2)),
        kk_rmse=float(np.sqrt(np.mean((P["r_kk"] - P["r_xx"]) ** 2))),
        xk_slope=float(np.sum(P["r_xk"][diff & strong] * P["r_xx"][diff & strong]) / np.sum(P["r_xx"][diff & strong] ** 2)),
        xk_rmse=float(np.sqrt(np.mean((P["r_xk"][diff] - P["r_xx"][diff]) ** 2))),
        self_mean=float(np.mean(P["self_r2"])),
        self_q90=float(np.quantile(P["self_r2"], 0.9)),
    )


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if "windows" not in n and "hmmfile" not in n])
def test_gof_statistics_match_reference(name):
    prob, f = load_golden(name)
    rows = np.asarray(prob.unrelated, dtype=np.int64)
    # whole samples only (both haplotypes unrelated), as the BED holds genotypes
    s = set(rows.tolist())
    rows = np.array([r for r in rows if (r ^ 1) in s], dtype=np.int64)
    hk_dev, _ = kg.generate(prob.H, prob.p, prob.K, prob.ref_local, prob.b, prob.mu, prob.groups,
                            unrelated=prob.unrelated, seed=2024)
    X = _geno(prob.H, prob.p, rows)
    ours = summarise(gof_panels(X, _geno(hk_dev, prob.p, rows), prob.groups))
    ref = summarise(gof_panels(X, _geno(f["Hk"], prob.p, rows), prob.groups))
    n = X.shape[0]
    se = 1.0 / np.sqrt(n)   # order of the sampling error of a correlation on n samples
    # panel 1: knockoff allele frequencies are unbiased copies of the originals
    assert abs(ours["maf_bias"]) < 0.5 * se and abs(ours["maf_bias"] - ref["maf_bias"]) < 0.5 * se
    assert ours["maf_rmse"] < 2.0 * se and abs(ours["maf_rmse"] - ref["maf_rmse"]) < 0.5 * se
    # panels 2 and 3: correlations among knockoffs, and between originals and knockoffs of other groups,
    # reproduce the correlations among originals; same fit as the reference's knockoffs
    # (at coarse resolutions few strongly correlated pairs sit in different groups: wider band for panel 3)
    for k, tol, lo in (("kk_slope", 0.03, 0.9), ("xk_slope", 0.12, 0.6)):
        assert abs(ours[k] - ref[k]) < tol, (k, ours[k], ref[k])
        assert lo < ours[k] < 1.15, (k, ours[k])
    for k in ("kk_rmse", "xk_rmse"):
        assert abs(ours[k] - ref[k]) < 1.0 * se, (k, ours[k], ref[k])
    # panel 4: self-similarity distribution
    assert abs(ours["self_mean"] - ref["self_mean"]) < 0.03, (ours["self_mean"], ref["self_mean"])
    assert abs(ours["self_q90"] - ref["self_q90"]) < 0.08, (ours["self_q90"], ref["self_q90"])
