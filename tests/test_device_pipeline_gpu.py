"""Device-resident pipeline: pre-step kernel, trend-fit kernel and the full on-device
DESeq() Wald path against the numpy restatement of the R callers (deseq2_b200/pipeline.py) driving the same engine."""
import numpy as np
import pytest

from helpers import DEV, rel_err

pytestmark = pytest.mark.gpu



def _setup(n, m, x=None, seed=1):
    import torch
    from deseq2_b200 import device as D, synth
    d = synth.make_example_counts(n, m, x=x, seed=seed)
    y = D.to_gene_major(d["counts"], torch.device(DEV))
    return d, y


@pytest.mark.parametrize("design", ["condition", "batch"])
def test_prep_kernel_matches_numpy(engine, design, n=2000):
    from deseq2_b200 import device_pipeline as DP, pipeline, synth
    m = 30
    x = synth.design_condition(m) if design == "condition" else synth.design_batch_condition(m, 3)
    d, y = _setup(n, m, x=x, seed=3)
    sf = d["sizeFactors"]
    pr = DP.prep(y, x, sf)
    counts = d["counts"]
    mv = pipeline.getBaseMeansAndVariances(counts, sf)
    nz = ~mv["allZero"]
    assert np.array_equal(pr["allZero"].cpu().numpy().astype(bool), mv["allZero"])
    assert np.max(rel_err(pr["baseMean"].cpu().numpy()[nz], mv["baseMean"][nz])) < 1e-12
    assert np.max(rel_err(pr["baseVar"].cpu().numpy()[nz], mv["baseVar"][nz])) < 1e-10
    norm = counts / sf
    rough = pipeline.roughDispEstimate(norm[nz], x)
    mom = pipeline.momentsDispEstimate(mv["baseMean"][nz], mv["baseVar"][nz], sf)
    a0 = np.minimum(np.maximum(1e-8, np.minimum(rough, mom)), max(10, m))
    assert np.max(rel_err(pr["alpha0"].cpu().numpy()[nz], a0, floor=1e-8)) < 1e-8
    mu = np.maximum(pipeline.linearModelMu(norm[nz], x) * sf, 0.5)
    assert np.max(rel_err(pr["mu_lin"].cpu().numpy()[nz][:, :m], mu)) < 1e-10
    Q, R = np.linalg.qr(x)
    b0 = np.linalg.solve(R, Q.T @ np.log(norm[nz] + 0.1).T).T
    assert np.max(np.abs(pr["beta0"].cpu().numpy().T[nz] - b0)) < 1e-10


def test_trend_kernel_matches_numpy(engine, n=20000):
    import torch
    from deseq2_b200 import device_pipeline as DP, pipeline
    rng = np.random.default_rng(0)
    means = 10 ** rng.uniform(0, 4, n)
    disps = (0.1 + 4 / means) * rng.gamma(8, 1 / 8, n)
    disps[rng.random(n) < 0.02] = 1e-8              # genes at the floor are excluded from the fit
    disps[rng.random(n) < 0.01] *= 40               # outliers trimmed by the residual rule
    ref = pipeline.parametricDispersionFit(means[disps > 1e-6], disps[disps > 1e-6])
    dev = torch.device(DEV)
    out = DP.trend_fit(torch.as_tensor(means, device=dev), torch.as_tensor(disps, device=dev)).cpu().numpy()
    assert out[2] == 0
    assert np.max(rel_err(out[:2], ref)) < 1e-8


@pytest.mark.parametrize("design,n,m", [("condition", 6000, 40), ("batch", 3000, 36), ("factor10", 1500, 120)])
def test_device_pipeline_matches_host_pipeline(engine, design, n, m):
    from deseq2_b200 import device_pipeline as DP, pipeline, synth
    x = {"condition": synth.design_condition(m), "batch": synth.design_batch_condition(m, 3),
         "factor10": synth.design_factor(m, 10)}[design]
    d, y = _setup(n, m, x=x, seed=11)
    host = pipeline.DESeq(d["counts"], x, sizeFactors=d["sizeFactors"], engine=engine)
    dv = DP.DESeq_device(y, x, d["sizeFactors"])
    idx = dv["idx"].cpu().numpy()
    assert np.array_equal(idx, np.flatnonzero(~host["allZero"]))
    assert np.max(rel_err(dv["trendCoefs"].cpu().numpy(), host["trendCoefs"])) < 1e-6
    assert abs(dv["dispPriorVar"] - host["dispPriorVar"]) < 1e-6 * max(1.0, host["dispPriorVar"])
    # per-gene results: inputs of the line searches differ by rounding between numpy and the device pre-steps, so a
    # few knife-edge genes may stop one step apart; everything else must agree to 1e-6
    for k in ("dispGeneEst", "dispMAP", "dispersion"):
        e = rel_err(dv[k].cpu().numpy(), host[k][idx])
        assert np.mean(e < 1e-6) > 0.97, (k, np.mean(e < 1e-6))
        assert np.quantile(e, 0.999) < 5e-2, (k, np.quantile(e, 0.999))
    conv = host["betaConv"][idx] == 1
    e = rel_err(dv["betaMatrix"].cpu().numpy()[conv], host["betaMatrix"][idx][conv], floor=1e-6)
    assert np.mean(e < 1e-5) > 0.97
    pv = dv["WaldPvalue"].cpu().numpy()
    assert np.all((pv[conv] >= 0) & (pv[conv] <= 1))
    se = rel_err(dv["betaSE"].cpu().numpy()[conv], host["betaSE"][idx][conv])
    assert np.mean(se < 1e-5) > 0.97


@pytest.mark.parametrize("design,m", [("condition", 12), ("condition", 60), ("batch", 36), ("covariate", 20),
                                      ("factor10", 200)])
def test_cooks_kernel_matches_numpy(engine, design, m, n=800):
    """robust moments dispersion (per-cell trimmed means), Cook's distances and their maximum
    (R/core.R:2277-2359) on device vs the numpy restatement."""
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    rng = np.random.default_rng(5)
    if design == "condition":
        x = synth.design_condition(m)
    elif design == "batch":
        x = synth.design_batch_condition(m, 3)
    elif design == "factor10":
        x = synth.design_factor(m, 10)
    else:
        x = np.c_[np.ones(m), rng.normal(0, 1, m)]          # every sample its own cell -> trimmedVariance branch
    d = synth.make_example_counts(n, m, x=x if design != "covariate" else None, seed=17)
    counts = d["counts"][d["counts"].sum(axis=1) > 0]
    counts[::7, 0] *= 30                                    # planted outliers (test_outlier.R:33-55 idea)
    sf = d["sizeFactors"]
    n = len(counts)
    mu = np.maximum((counts / sf).mean(axis=1, keepdims=True) * sf[None, :] * rng.uniform(0.8, 1.25, (n, m)), 0.5)
    H = rng.uniform(0.01, 0.6, (n, m))
    dev = torch.device(DEV)
    out = DP.cooks(D.to_gene_major(counts, dev), D.to_gene_major(mu, dev), D.to_gene_major(H, dev), x, sf)
    ref_disp = pipeline.robustMethodOfMomentsDisp(counts, sf, x)
    ref_cooks = pipeline.calculateCooksDistance(counts, mu, H, sf, x)
    ref_max = pipeline.recordMaxCooks(x, ref_cooks)
    assert np.max(rel_err(out["robustDisp"].cpu().numpy(), ref_disp)) < 1e-10
    assert np.max(rel_err(out["cooks"].cpu().numpy()[:, :m], ref_cooks, floor=1e-12)) < 1e-9
    got_max = out["maxCooks"].cpu().numpy()
    if np.all(np.isnan(ref_max)):
        assert np.all(np.isnan(got_max))
    else:
        assert np.max(rel_err(got_max, ref_max, floor=1e-12)) < 1e-9


def test_lrt_device_matches_host(engine, n=3000):
    """BASELINE.json config 5's call sequence (nbinomLRT ~batch+condition vs ~batch) on device vs the host glue."""
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    m = 40
    full = synth.design_batch_condition(m, 2)
    reduced = full[:, :2]
    d = synth.make_example_counts(n, m, x=full, seed=41, betaSD=0.7)
    counts = d["counts"][d["counts"].sum(axis=1) > 0]
    sf = d["sizeFactors"]
    nf = np.broadcast_to(sf[None, :], counts.shape)
    alpha = np.clip(0.1 + 4 / (counts / sf).mean(axis=1), 1e-8, m)
    host = pipeline.nbinomLRT(counts, nf, full, reduced, alpha, engine=engine)
    dev = torch.device(DEV)
    got = DP.nbinomLRT_device(D.to_gene_major(counts, dev), full, reduced, sf, torch.as_tensor(alpha, device=dev))
    ok = host["fullBetaConv"] & host["reducedBetaConv"]
    # the statistic is built from log-likelihoods at the UNCLAMPED means (R/fitNbinomGLMs.R:180-182): genes with fitted
    # means below minmu (zeros in a whole design cell) are part of the comparison
    assert ((counts / sf).min(axis=1) <= 0.5)[ok].sum() > 50
    st = got["LRTStatistic"].cpu().numpy()
    assert np.max(np.abs(st[ok] - host["LRTStatistic"][ok]) / (1 + np.abs(host["deviance"][ok]))) < 1e-9
    pv = got["LRTPvalue"].cpu().numpy()
    big = ok & (host["LRTPvalue"] > 1e-12)
    assert np.max(rel_err(pv[big], host["LRTPvalue"][big])) < 1e-6
    assert np.max(rel_err(got["betaMatrix"].cpu().numpy()[ok], host["betaMatrix"][ok], floor=1e-6)) < 1e-6
