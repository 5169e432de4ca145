"""Median-of-ratios size factors on the device (csrc/size_factors.cu) against the numpy
restatement of estimateSizeFactorsForMatrix (R/core.R:535-578), and the device-resident DESeq() started from raw
counts.  The same functions run on the emulated engine in tests/test_emulated_kernels.py (DEV = "cpu")."""
import numpy as np
import pytest

from helpers import DEV, rel_err

pytestmark = pytest.mark.gpu



def _poscounts_reference(counts):
    """R/core.R:544-549 + :560-563: zeros contribute 0 to the row mean of logs; all-zero rows are left out."""
    counts = np.asarray(counts, dtype=np.float64)
    with np.errstate(divide="ignore"):
        lc = np.log(counts)
    lc0 = np.where(np.isfinite(lc), lc, 0.0)
    lg = lc0.mean(axis=1)
    lg[counts.sum(axis=1) == 0] = -np.inf
    sf = np.empty(counts.shape[1])
    for j in range(counts.shape[1]):
        sel = np.isfinite(lg) & (counts[:, j] > 0)
        sf[j] = np.exp(np.median(lc[sel, j] - lg[sel])) if sel.any() else np.nan
    return sf, lg


@pytest.mark.parametrize("n,m,seed,as_double", [(4000, 12, 1, False), (2501, 37, 2, True), (600, 130, 3, False),
                                                (64, 5, 4, False)])
def test_size_factors_match_numpy(engine, n, m, seed, as_double):
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    d = synth.make_example_counts(n, m, seed=seed, interceptMean=6.0)
    counts = d["counts"].copy()
    counts[::11] = 0                                        # all-zero genes
    counts[1::13, 0] = 0                                    # genes with a single zero: out for "ratio", in for "poscounts"
    counts[5:40] = counts[5]                                # ties: many identical ratios around the median
    y = D.to_gene_major(counts.astype(np.float64) if as_double else counts, torch.device(DEV))
    got = DP.size_factors(y, m)
    ref = pipeline.estimateSizeFactorsForMatrix(counts)
    assert np.max(rel_err(got["sizeFactors"].cpu().numpy(), ref)) < 1e-12
    with np.errstate(divide="ignore"):
        lg = np.log(counts.astype(np.float64)).mean(axis=1)
    glg = got["loggeomeans"].cpu().numpy()
    assert np.array_equal(np.isfinite(glg), np.isfinite(lg))
    assert np.max(rel_err(glg[np.isfinite(lg)], lg[np.isfinite(lg)])) < 1e-13
    got2 = DP.size_factors(y, m, type="poscounts")
    ref2, lg2 = _poscounts_reference(counts)
    assert np.max(rel_err(got2["sizeFactors"].cpu().numpy(), ref2)) < 1e-12
    assert np.array_equal(np.isfinite(got2["loggeomeans"].cpu().numpy()), np.isfinite(lg2))


def test_size_factors_degenerate_inputs(engine):
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP
    dev = torch.device(DEV)
    counts = np.array([[0, 3, 5], [2, 0, 7], [4, 1, 0]], dtype=np.int32)        # every gene has a zero
    with pytest.raises(ValueError, match="every gene contains at least one zero"):
        DP.size_factors(D.to_gene_major(counts, dev), 3)
    sf = DP.size_factors(D.to_gene_major(counts, dev), 3, type="poscounts")["sizeFactors"].cpu().numpy()
    lg = np.array([np.log([3, 5]).sum(), np.log([2, 7]).sum(), np.log([4, 1]).sum()]) / 3
    want = [np.exp(np.median([np.log(2) - lg[1], np.log(4) - lg[2]])), np.exp(np.median([np.log(3) - lg[0], np.log(1) - lg[2]])),
            np.exp(np.median([np.log(5) - lg[0], np.log(7) - lg[1]]))]
    assert np.allclose(sf, want, rtol=1e-13)
    one = np.array([[10, 20, 40, 80]], dtype=np.int32)                          # a single gene, even sample count
    sf1 = DP.size_factors(D.to_gene_major(one, dev), 4)["sizeFactors"].cpu().numpy()
    assert np.allclose(sf1, one[0] / np.exp(np.log(one[0]).mean()), rtol=1e-13)
    col = np.array([[5, 0], [7, 0], [9, 0]], dtype=np.int32)                    # a sample with no usable count
    sfc = DP.size_factors(D.to_gene_major(col, dev), 2, type="poscounts")["sizeFactors"].cpu().numpy()
    assert np.isfinite(sfc[0]) and np.isnan(sfc[1])


def test_deseq_device_from_raw_counts(engine, n=3000):
    """DESeq_device(sizeFactors=None) == DESeq_device(sizeFactors = the numpy median of ratios)."""
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    m = 24
    x = synth.design_condition(m)
    d = synth.make_example_counts(n, m, x=x, seed=9)
    y = D.to_gene_major(d["counts"], torch.device(DEV))
    sf = pipeline.estimateSizeFactorsForMatrix(d["counts"])
    a = DP.DESeq_device(y, x)
    assert np.max(rel_err(a["sizeFactors"], sf)) < 1e-12
    b = DP.DESeq_device(y, x, sf)
    for k in ("dispGeneEst", "dispersion", "betaMatrix", "betaSE", "WaldPvalue", "maxCooks"):
        e = rel_err(a[k].cpu().numpy(), b[k].cpu().numpy(), floor=1e-9)
        assert np.mean(e < 1e-6) > 0.99, k


@pytest.mark.parametrize("design", ["condition", "mixed"])
def test_outlier_replacement_and_refit_device_vs_host(engine, design, n=2500):
    """replaceOutliers + refitWithoutOutliers (R/core.R:2069-2115, 2484-2565; DESeq's default for cells with >= 7
    replicates): the device-resident pipeline against the numpy restatement driving the same engine.  "mixed" has a
    cell of 3 samples (used for Cook's, not replaceable) next to cells of 9, so maxCooks is recomputed, not NA."""
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    if design == "condition":
        m = 20
        x = synth.design_condition(m)
    else:
        g = np.r_[np.zeros(9), np.ones(9), np.full(3, 2)].astype(int)
        m = len(g)
        x = np.zeros((m, 3))
        x[:, 0] = 1
        x[g == 1, 1] = 1
        x[g == 2, 2] = 1
    d = synth.make_example_counts(n, m, x=x, seed=23, interceptMean=5.0)
    counts = d["counts"].copy()
    rng = np.random.default_rng(3)
    planted = rng.choice(n, n // 25, replace=False)
    counts[planted, rng.integers(0, m, len(planted))] += 100 + 12 * counts[planted].max(axis=1)    # count outliers
    sf = d["sizeFactors"]
    host = pipeline.DESeq(counts, x, sizeFactors=sf, engine=engine, minReplicatesForReplace=7)
    dv = DP.DESeq_device(D.to_gene_major(counts, torch.device(DEV)), x, sf, minReplicatesForReplace=7)
    idx = dv["idx"].cpu().numpy()
    assert np.array_equal(idx, np.flatnonzero(~host["allZero"]))
    raw = pipeline.DESeq(counts, x, sizeFactors=sf, engine=engine, useOptim=False)
    # rows whose first-pass IRLS diverged go to the box-constrained maximiser on both sides (L-BFGS-B on the host, projected
    # Newton on the device: tests/test_device_pipeline_gpu.py::test_optim_fallback_device_vs_host); they stop at different
    # distances from the optimum along flat directions, so they are compared there, not here
    ok1 = raw["betaConv"][idx] == 1
    rep_h = host["replace"][idx]
    rep_d = dv["replace"].cpu().numpy()
    assert rep_h.sum() >= len(planted) * 0.5 and np.mean((rep_h == rep_d)[ok1]) > 0.995 and ok1.mean() > 0.97
    assert np.array_equal(dv["replaceable"].cpu().numpy(), host["replaceable"])
    both = rep_h & rep_d & ok1
    assert np.max(rel_err(dv["baseMean"].cpu().numpy()[both], host["baseMean"][idx][both])) < 1e-12
    for k in ("dispGeneEst", "dispersion"):
        e = rel_err(dv[k].cpu().numpy()[both], host[k][idx][both])
        assert np.mean(e < 1e-6) > 0.95, (k, np.mean(e < 1e-6))
    conv = both & (host["betaConv"][idx] == 1)
    e = rel_err(dv["betaMatrix"].cpu().numpy()[conv], host["betaMatrix"][idx][conv], floor=1e-6)
    assert np.mean(e < 1e-5) > 0.95
    mc_h, mc_d = host["maxCooks"][idx], dv["maxCooks"].cpu().numpy()
    if design == "condition":
        assert np.all(np.isnan(mc_h)) and np.all(np.isnan(mc_d))          # every sample replaceable: column set to NA
    else:
        # both sides use the unclamped fitted mean nf * exp(x beta) of R/fitNbinomGLMs.R:180 (an all-zero cell has a
        # fitted mean far below minmu)
        assert np.max(rel_err(mc_d[ok1], mc_h[ok1], floor=1e-12)) < 1e-8
    # the refit pulls the planted genes' fold changes back to what the data say without the outlier
    clean = pipeline.DESeq(d["counts"], x, sizeFactors=sf, engine=engine)
    pl = np.intersect1d(planted, np.flatnonzero(host["replace"]))
    err_refit = np.abs(host["betaMatrix"][pl, 1] - clean["betaMatrix"][pl, 1])
    err_raw = np.abs(raw["betaMatrix"][pl, 1] - clean["betaMatrix"][pl, 1])
    assert np.median(err_refit) < 0.25 * np.median(err_raw)


def test_get_contrast_device_matches_host(engine, n=1500):
    """getContrast (R/results.R:760-827; the one direct R caller of fitBeta besides the wrappers) on device tensors
    against the host glue: a level-vs-level contrast of a 3-level factor after a device-resident DESeq()."""
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    m = 18
    g = np.arange(m) % 3
    x = np.c_[np.ones(m), g == 1, g == 2].astype(float)
    d = synth.make_example_counts(n, m, x=x, seed=31, betaSD=0.8)
    sf = d["sizeFactors"]
    y = D.to_gene_major(d["counts"], torch.device(DEV))
    dv = DP.DESeq_device(y, x, sf)
    idx = dv["idx"]
    ynz = y[idx].contiguous()
    cnz = d["counts"][idx.cpu().numpy()]
    contrast = np.array([0.0, 1.0, -1.0])
    got = DP.getContrast_device(ynz, x, sf, dv["dispersion"], dv["betaMatrix"], contrast)
    ref = pipeline.getContrast(cnz, np.broadcast_to(sf[None, :], cnz.shape), x, dv["dispersion"].cpu().numpy(),
                               dv["betaMatrix"].cpu().numpy(), contrast, engine=engine)
    ok = dv["betaConv"].cpu().numpy()
    for k in ("log2FoldChange", "lfcSE", "stat"):
        assert np.max(rel_err(got[k].cpu().numpy()[ok], ref[k][ok], floor=1e-9)) < 1e-9, k
    assert np.allclose(got["log2FoldChange"].cpu().numpy()[ok],
                       (dv["betaMatrix"][:, 1] - dv["betaMatrix"][:, 2]).cpu().numpy()[ok], atol=1e-12)
    pv = got["pvalue"].cpu().numpy()[ok]
    assert np.all((pv >= 0) & (pv <= 1))
