"""Host glue (deseq2_b200/pipeline.py, the numpy restatement of the R callers) driven by the oracle engine:
BASELINE.json config 1 (makeExampleDESeqDataSet(n=1000, m=6), ~condition, Wald, CPU plumbing)."""
import numpy as np
import pytest

from deseq2_b200 import pipeline, synth


def test_c1_pipeline_recovers_truth(oracle):
    d = synth.make_example_counts(1000, 6, seed=20260923 + 1)
    r = pipeline.DESeq(d["counts"], d["x"], engine=oracle)
    nz = ~r["allZero"]
    assert np.all(np.isnan(r["dispersion"][~nz])) and np.all(np.isfinite(r["dispersion"][nz]))
    assert 0.02 < r["trendCoefs"][0] < 0.5 and 1.0 < r["trendCoefs"][1] < 12.0      # truth: 0.1 + 4/mean
    hi = nz & (r["baseMean"] > 50)
    assert np.corrcoef(r["betaMatrix"][hi, 1], d["trueBeta"][hi, 1])[0, 1] > 0.85
    assert 0.5 < np.nanmedian(r["dispersion"][hi] / d["trueDisp"][hi]) < 2.0
    assert np.nanmean(r["betaConv"]) > 0.99
    assert np.all((r["WaldPvalue"][nz] >= 0) & (r["WaldPvalue"][nz] <= 1))


def test_size_factors_and_trend_fit():
    d = synth.make_example_counts(4000, 12, seed=5)
    sf = pipeline.estimateSizeFactorsForMatrix(d["counts"])
    assert np.allclose(sf / np.exp(np.mean(np.log(sf))), d["sizeFactors"], rtol=0.05)
    rng = np.random.default_rng(0)
    means = 10 ** rng.uniform(0, 4, 5000)
    disps = (0.1 + 4 / means) * rng.gamma(20, 1 / 20, 5000)
    c = pipeline.parametricDispersionFit(means, disps)
    assert abs(c[0] - 0.1) < 0.01 and abs(c[1] - 4) < 0.3


def test_linear_mu_matches_group_means():
    d = synth.make_example_counts(50, 8, seed=9)
    norm = d["counts"] / d["sizeFactors"]
    mu = pipeline.linearModelMu(norm, d["x"])
    assert np.allclose(mu[:, :4], norm[:, :4].mean(axis=1, keepdims=True))
    assert np.allclose(mu[:, 4:], norm[:, 4:].mean(axis=1, keepdims=True))


def test_lrt_full_vs_reduced(oracle):
    """BASELINE.json config 5 shape (nbinomLRT ~batch+condition vs ~batch), small, oracle engine: the LRT statistic is
    2 (l_full - l_reduced) >= 0 and genes with a true condition effect get small p-values."""
    m = 24
    full = synth.design_batch_condition(m, 2)            # intercept, batch, condition  (p = 3)
    reduced = full[:, :2]                                 # ~batch
    d = synth.make_example_counts(600, m, x=full, seed=31, betaSD=1.0)
    counts = d["counts"][d["counts"].sum(axis=1) > 0]
    sf = d["sizeFactors"]
    nf = np.broadcast_to(sf[None, :], counts.shape)
    alpha = np.clip(0.1 + 4 / (counts / sf).mean(axis=1), 1e-8, m)
    r = pipeline.nbinomLRT(counts, nf, full, reduced, alpha, engine=oracle)
    # low-count genes are excluded: fitBeta clamps mu at minmu = 0.5 during the fit while the log-likelihood is
    # evaluated at the unclamped mu (R/fitNbinomGLMs.R:180-182), so nesting need not hold there -- in R either
    ok = r["fullBetaConv"] & r["reducedBetaConv"] & ((counts / sf).mean(axis=1) > 5)
    # both fits stop at a relative deviance change of 1e-8, so the statistic is only non-negative up to that
    assert r["df"] == 1 and np.all(r["LRTStatistic"][ok] > -1e-6 * np.abs(r["deviance"][ok]))
    truth = np.abs(d["trueBeta"][d["counts"].sum(axis=1) > 0][:, 2])
    hi = ok & ((counts / sf).mean(axis=1) > 50)
    assert np.median(r["LRTPvalue"][hi & (truth > 1.5)]) < 1e-3 < np.median(r["LRTPvalue"][hi & (truth < 0.1)])
    # reduced = ~1 takes the closed-form branch (R/fitNbinomGLMs.R:99-137): the mean of normalised counts, which is
    # the NB MLE only for equal size factors -- close to, not equal to, an explicit IRLS fit
    ones = np.ones((m, 1))
    a = pipeline._fit_intercept_only(counts, nf, alpha)
    b = pipeline.fitNbinomGLMs(counts, nf, ones, alpha, engine=oracle)
    conv = b["betaConv"] & ((counts / sf).min(axis=1) > 2)
    assert np.allclose(a["betaMatrix"][conv], b["betaMatrix"][conv], atol=0.1)
    assert np.all(a["logLike"][conv] <= b["logLike"][conv] + 1e-6 * np.abs(b["logLike"][conv]))
    r1 = pipeline.nbinomLRT(counts, nf, full, ones, alpha, engine=oracle)
    assert r1["df"] == 2 and np.all(np.isfinite(r1["LRTPvalue"][ok]))


def test_weighted_quantile_reduces_to_plain_quantile():
    """Hmisc.wtd.quantile with equal weights is R's type-7 quantile."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=501)
    for p in (0.5, 0.9, 0.95):
        assert abs(pipeline.wtd_quantile(x, np.ones_like(x), p) - np.quantile(x, p)) < 1e-12
    # a heavy weight drags the quantile towards its value
    w = np.ones_like(x)
    w[np.argmax(x)] = 200.0
    assert pipeline.wtd_quantile(x, w, 0.9) > np.quantile(x, 0.9)


def test_beta_prior_pipeline_config4_shape(oracle):
    """BASELINE.json config 4's call sequence at a small size: 10-level factor, MLE pass on the standard matrix
    (p = 10), estimateBetaPriorVar, MAP pass on the 11-column expanded matrix with lambda = 1/betaPriorVar."""
    m, levels = 60, 10
    g = (np.arange(m) * levels) // m
    x = synth.design_factor(m, levels)
    d = synth.make_example_counts(500, m, x=x, seed=51, betaSD=0.8)
    counts = d["counts"][d["counts"].sum(axis=1) > 0]
    sf = d["sizeFactors"]
    nf = np.broadcast_to(sf[None, :], counts.shape)
    bm = (counts / sf).mean(axis=1)
    dispFit = 0.1 + 4 / bm
    xs, names, fidx, _ = pipeline.factorDesign([g], expanded=False)
    assert np.array_equal(xs, x)
    r = pipeline.fitGLMsWithPrior(counts, nf, [g], np.clip(dispFit, 1e-8, m), bm, dispFit, engine=oracle)
    assert r["modelMatrix"].shape == (m, 11) and np.linalg.matrix_rank(r["modelMatrix"]) == 10
    pv = r["betaPriorVar"]
    assert pv[0] == 1e6 and np.allclose(pv[1:], pv[1]) and 0.05 < pv[1] < 5.0     # truth: betaSD^2 = 0.64 (log2)
    fit = r["fit"]
    conv = fit["betaConv"] & r["mle"]["betaConv"] & (bm > 20)
    # level effects of the MAP fit sum to ~0 within a gene (symmetric ridge on a rank-deficient design) ...
    assert np.max(np.abs(fit["betaMatrix"][conv, 1:].sum(axis=1))) < 1e-3
    # ... and are shrunken versions of the MLE contrasts
    mle_c = r["mleBetaMatrix"][conv, 1:]                       # level l vs level 1
    map_c = fit["betaMatrix"][conv, 2:] - fit["betaMatrix"][conv, 1:2]
    assert np.median(np.abs(map_c)) < np.median(np.abs(mle_c))
    assert np.corrcoef(map_c.ravel(), mle_c.ravel())[0, 1] > 0.98


def test_optim_fallback_matches_irls_and_rescues_divergence(oracle):
    """test_optim.R:2-39: forceOptim (L-BFGS-B on every row) agrees with IRLS on beta and SE for a badly scaled
    covariate, and the 0/1000 row that IRLS abandons (iter == maxit) gets a finite optim estimate."""
    rng = np.random.default_rng(7)
    m = 20
    x = np.c_[np.ones(m), rng.normal(0, 1000, m)]
    d = synth.make_example_counts(40, m, seed=77)
    counts = d["counts"][d["counts"].min(axis=1) > 3]
    nf = np.ones(counts.shape)
    alpha = np.full(len(counts), 0.1)
    a = pipeline.fitNbinomGLMs(counts, nf, x, alpha, engine=oracle, useOptim=False)
    b = pipeline.fitNbinomGLMs(counts, nf, x, alpha, engine=oracle, forceOptim=True)
    conv = a["betaConv"]
    assert conv.mean() > 0.8
    assert np.allclose(a["betaMatrix"][conv, 0], b["betaMatrix"][conv, 0], atol=1e-4)
    assert np.allclose(a["betaMatrix"][conv, 1] * 1000, b["betaMatrix"][conv, 1] * 1000, atol=5e-3)
    assert np.allclose(a["betaSE"][conv], b["betaSE"][conv], rtol=1e-2)
    y = np.array([[0, 0, 0, 0, 0, 1000, 1000, 0, 0, 0]], dtype=np.int32)
    x2 = np.c_[np.ones(10), np.r_[np.zeros(5), np.ones(5)]]
    r0 = pipeline.fitNbinomGLMs(y, np.ones((1, 10)), x2, np.array([0.1]), engine=oracle, useOptim=False)
    r1 = pipeline.fitNbinomGLMs(y, np.ones((1, 10)), x2, np.array([0.1]), engine=oracle)   # reference default: useOptim
    assert r0["betaIter"][0] == 100 and not r0["betaConv"][0]
    assert np.all(np.isfinite(r1["betaMatrix"])) and np.all(np.abs(r1["betaMatrix"]) <= 30)
    assert r1["logLike"][0] >= r0["logLike"][0] - 1e-6


def test_replace_outliers_semantics():
    """R/core.R:2069-2115 (test_outlier.R:2-31 analogue): only entries above the Cook's cutoff in samples whose cell has
    >= minReplicates samples are replaced, by as.integer(trimmed mean of normalised counts * size factor); the gene
    is flagged even when its outlier sits in a non-replaceable sample."""
    from scipy import stats
    g = np.r_[np.zeros(8), np.ones(8), np.full(3, 2)].astype(int)
    m = len(g)
    x = np.zeros((m, 3))
    x[:, 0] = 1
    x[g == 1, 1] = 1
    x[g == 2, 2] = 1
    rng = np.random.default_rng(0)
    counts = rng.poisson(50, (6, m)).astype(np.int32)
    sf = np.exp(rng.normal(0, 0.2, m))
    cooks = np.full((6, m), 0.1)
    cut = stats.f.ppf(0.99, 3, m - 3)
    cooks[1, 2] = cut * 1.5          # replaceable sample (cell of 8)
    cooks[3, 17] = cut * 2.0         # sample in the cell of 3: flagged, not replaced
    cooks[4, 5] = np.nan             # NA never counts
    new, replace, replaceable = pipeline.replaceOutliers(counts, cooks, sf, x, minReplicates=7)
    assert np.array_equal(replaceable, g < 2)
    assert np.array_equal(replace, [False, True, False, True, False, False])
    changed = np.argwhere(new != counts)
    assert changed.tolist() == [[1, 2]]
    norm = np.sort(counts[1] / sf)
    lo = int(np.floor(m * 0.2))
    assert new[1, 2] == int(norm[lo:m - lo].mean() * sf[2])
    assert pipeline.nOrMoreInCell(x, 3).all() and not pipeline.nOrMoreInCell(x, 9).any()
    with pytest.raises(ValueError):
        pipeline.replaceOutliers(counts, cooks, sf, x, minReplicates=2)


def test_refit_without_outliers_recovers_fold_changes(oracle):
    """test_outlier.R:33-55 idea through the whole host pipeline: planted count outliers get Cook's distances above
    qf(.99, p, m - p), are replaced, and the refit moves the fold changes back to the outlier-free estimates."""
    m = 20
    x = synth.design_condition(m)
    d = synth.make_example_counts(400, m, x=x, seed=5, interceptMean=5.0)
    counts = d["counts"].copy()
    rng = np.random.default_rng(1)
    planted = rng.choice(400, 16, replace=False)
    counts[planted, rng.integers(0, m, 16)] += 100 + 12 * counts[planted].max(axis=1)
    sf = d["sizeFactors"]
    clean = pipeline.DESeq(d["counts"], x, sizeFactors=sf, engine=oracle)
    raw = pipeline.DESeq(counts, x, sizeFactors=sf, engine=oracle)
    fit = pipeline.DESeq(counts, x, sizeFactors=sf, engine=oracle, minReplicatesForReplace=7)
    assert fit["replaceable"].all() and np.all(np.isnan(fit["maxCooks"]))
    hit = np.intersect1d(planted, np.flatnonzero(fit["replace"]))
    assert len(hit) >= 10 and fit["n_replaced"] < 40
    untouched = ~fit["replace"] & ~fit["allZero"]
    assert np.array_equal(fit["betaMatrix"][untouched], raw["betaMatrix"][untouched])
    err_fit = np.abs(fit["betaMatrix"][hit, 1] - clean["betaMatrix"][hit, 1])
    err_raw = np.abs(raw["betaMatrix"][hit, 1] - clean["betaMatrix"][hit, 1])
    assert np.median(err_fit) < 0.25 * np.median(err_raw)
    assert np.array_equal(fit["replaceCounts"][~fit["replace"]], counts[~fit["replace"]])


def test_get_contrast_equals_reparametrised_fit(oracle):
    """R/results.R:760-827: contrast e_k reproduces the fitted coefficient and its SE; a level-vs-level contrast of a
    3-level factor equals the coefficient of the model refitted with the other reference level."""
    m = 18
    g = np.arange(m) % 3
    x = np.c_[np.ones(m), g == 1, g == 2].astype(float)               # reference level A: (Intercept, B_vs_A, C_vs_A)
    d = synth.make_example_counts(300, m, x=x, seed=12, betaSD=0.8)
    counts = d["counts"][d["counts"].sum(axis=1) > 0]
    sf = d["sizeFactors"]
    nf = np.broadcast_to(sf[None, :], counts.shape)
    disp = np.clip(0.05 + 3.0 / (counts / sf).mean(axis=1), 1e-8, m)
    fit = pipeline.nbinomWaldTest(counts, nf, x, disp, engine=oracle)
    ok = fit["betaConv"] & ((counts / sf).min(axis=1) > 1)
    c1 = pipeline.getContrast(counts, nf, x, disp, fit["betaMatrix"], [0, 1, 0], engine=oracle)
    assert np.allclose(c1["log2FoldChange"][ok], fit["betaMatrix"][ok, 1], rtol=0, atol=1e-12)
    assert np.allclose(c1["lfcSE"][ok], fit["betaSE"][ok, 1], rtol=1e-6)
    assert np.allclose(c1["pvalue"][ok], fit["WaldPvalue"][ok, 1], rtol=1e-5, atol=1e-300)
    x2 = np.c_[np.ones(m), g == 0, g == 1].astype(float)              # reference level C: (Intercept, A_vs_C, B_vs_C)
    fit2 = pipeline.nbinomWaldTest(counts, nf, x2, disp, engine=oracle)
    ok &= fit2["betaConv"]
    c = pipeline.getContrast(counts, nf, x, disp, fit["betaMatrix"], [0, 1, -1], engine=oracle)   # B vs C
    assert np.allclose(c["log2FoldChange"][ok], fit2["betaMatrix"][ok, 2], atol=2e-5)
    assert np.allclose(c["lfcSE"][ok], fit2["betaSE"][ok, 2], rtol=1e-4)
    with pytest.raises(ValueError):
        pipeline.getContrast(counts, nf, x, disp, fit["betaMatrix"], [1, -1], engine=oracle)
