"""Host glue (deseq2_b200/pipeline.py, the numpy restatement of the R callers) driven by the oracle engine:
BASELINE.json config 1 (makeExampleDESeqDataSet(n=1000, m=6), ~condition, Wald, CPU plumbing)."""
import numpy as np

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
