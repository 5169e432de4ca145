"""Device-resident DESeq() Wald path: counts go in once (gene-major, on the GPU), every per-gene step runs on the
device, only a handful of scalars ever reach the host.

Same sequence as deseq2_b200/pipeline.py (the numpy restatement of R/core.R:280-432 with test="Wald",
fitType="parametric", betaPrior=FALSE, no weights, no outlier replacement), but
  * the pre-steps (base mean/variance, rough + moments dispersion, linear-model mu, IRLS start values) are one
    CUDA kernel (b200nb_prep_dev, csrc/pipeline_kernels.cu),
  * the parametric dispersion-trend fit is one single-CTA kernel (b200nb_trend_fit_dev),
  * fitDisp / fitDispGrid / fitBeta are the engine's device entry points,
  * the elementwise rules between them (noIncrease, convergence flags, clamps, outlier rule, MAD, Wald statistic
    and p-value) are a few torch tensor ops -- plumbing, not the product.
This is SURVEY.md section 8(f) rows 2-3 (pre-steps and Wald statistics on device) layered on rows (a)-(e).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import device as D
from .pipeline import modelMatrixGroups

F64 = torch.float64
LN2 = float(np.log(2.0))


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_WS = {}


def _ws(name, shape, dtype, dev):
    """Reusable n x ld work matrices (fitted means, hat diagonals): at 20k x 1000 a fresh 160 MB torch allocation per
    call was measured to cost tens of ms in allocator / cudaMalloc churn, far more than the kernels that fill it.
    The tensors returned by DESeq_device under "mu" / "H" are views of this workspace: valid until the next call."""
    key = (name, str(dev))
    t = _WS.get(key)
    if t is None or t.shape != tuple(shape) or t.dtype != dtype:
        t = torch.empty(shape, dtype=dtype, device=dev)
        _WS[key] = t
    return t


def size_factors(y, m, type="ratio"):
    """b200nb_size_factors_dev: estimateSizeFactorsForMatrix (R/core.R:535-578, locfunc = median) on the device.
    y: gene-major (n, ld) device tensor, m samples.  Returns {"sizeFactors": (m,) device tensor, "loggeomeans": (n,)}.
    Raises like the reference when every gene contains a zero (the only host read: one int)."""
    if type not in ("ratio", "poscounts"):
        raise ValueError("type must be 'ratio' or 'poscounts'")
    L = _lib.lib()
    dev = y.device
    n, ld = y.shape
    out = {"sizeFactors": torch.empty(m, dtype=F64, device=dev), "loggeomeans": torch.empty(n, dtype=F64, device=dev)}
    nfin = torch.zeros(1, dtype=torch.int32, device=dev)
    gm = _ws("sf_ratios", (n, ld), F64, dev)
    cm = _ws("sf_ratios_cm", (n * m,), F64, dev)
    _lib.check(L.b200nb_size_factors_dev(_p(y), 1 if y.dtype == F64 else 0, int(type == "poscounts"), n, m, ld,
                                         _p(out["loggeomeans"]), _p(gm), _p(cm), _p(out["sizeFactors"]), _p(nfin),
                                         _stream()), "size_factors")
    if int(nfin.item()) == 0:
        raise ValueError("every gene contains at least one zero, cannot compute log geometric means")
    return out


def _projection(x):
    """(X'X)^-1 X' (p x m) WITHOUT a BLAS / LAPACK call.  numpy hands `x.T @ x` and linalg.solve to OpenBLAS, whose
    worker threads (one per CPU the process may run on: 64-128 on the GPU hosts) then busy-wait for ~100 ms; under a
    container CPU quota (the 1-GPU lease: cpu.max = 16 CPUs) that spin exhausts the quota and the kernel throttles the
    WHOLE process for the rest of the period -- measured: DESeq_device on the config-4 shape took 100-170 ms instead of
    54 ms, with random 40-90 ms holes in the host timeline (profiles/r02_pipeline_host_stalls.md).  The matrices
    are p x p with p <= 32: einsum's own loops and a Gauss-Jordan sweep with partial pivoting are plenty."""
    x = np.asarray(x, dtype=np.float64)
    m, p = x.shape
    a = np.einsum("ik,il->kl", x, x)                       # X'X (c_einsum loops, no BLAS)
    aug = np.concatenate([a, np.ascontiguousarray(x.T)], axis=1)   # [X'X | X']
    for c in range(p):
        piv = c + int(np.argmax(np.abs(aug[c:, c])))
        if aug[piv, c] == 0.0:
            raise np.linalg.LinAlgError("singular design matrix")
        if piv != c:
            aug[[c, piv]] = aug[[piv, c]]
        aug[c] /= aug[c, c]
        f = aug[:, c].copy()
        f[c] = 0.0
        aug -= f[:, None] * aug[c][None, :]
    return np.ascontiguousarray(aug[:, p:])


def prep(y, x, sizeFactors, minDisp=1e-8, minmu=0.5, want_mu=True, want_beta0=True):
    """b200nb_prep_dev.  y: gene-major (n, ld) int32/float64 device tensor; x: (m, p) numpy; sizeFactors: (m,) numpy."""
    L = _lib.lib()
    dev = y.device
    n, ld = y.shape
    x = np.asarray(x, dtype=np.float64)
    m, p = x.shape
    proj = _projection(x)                                      # (X'X)^-1 X', p x m
    xd = D.x_to_device(x, dev)
    projd = torch.as_tensor(np.ascontiguousarray(proj), device=dev)
    sfd = torch.as_tensor(np.asarray(sizeFactors, dtype=np.float64), device=dev)
    out = {"baseMean": torch.empty(n, dtype=F64, device=dev), "baseVar": torch.empty(n, dtype=F64, device=dev),
           "allZero": torch.empty(n, dtype=torch.int32, device=dev), "alpha0": torch.empty(n, dtype=F64, device=dev),
           "mu_lin": _ws("mu_lin", (n, ld), F64, dev) if want_mu else None,
           "beta0": torch.empty((p, n), dtype=F64, device=dev) if want_beta0 else None}
    rc = L.b200nb_prep_dev(_p(y), 0 if y.dtype == torch.int32 else 1, _p(xd), _p(projd), _p(sfd),
                           float(np.mean(1.0 / np.asarray(sizeFactors))), float(minDisp), float(max(10, m)),
                           float(minmu), n, m, p, ld, _p(out["baseMean"]), _p(out["baseVar"]), _p(out["allZero"]),
                           _p(out["alpha0"]), _p(out["mu_lin"]), _p(out["beta0"]), _stream())
    _lib.check(rc, "prep_dev")
    out["xd"], out["sfd"] = xd, sfd
    return out


def trend_fit(means, disps, minDisp=1e-8):
    """parametricDispersionFit on device: returns a 4-vector tensor (asymptDisp, extraPois, status, rounds)."""
    L = _lib.lib()
    out = torch.empty(4, dtype=F64, device=means.device)
    _lib.check(L.b200nb_trend_fit_dev(_p(means), _p(disps), means.numel(), float(minDisp), _p(out), _stream()),
               "trend_fit_dev")
    return out


def cooks(y, mu, hat, x, sizeFactors, want_matrix=True):
    """b200nb_cooks_dev: Cook's distances, their per-gene maximum and the robust moments dispersion."""
    from .pipeline import designCells
    L = _lib.lib()
    dev = y.device
    n, ld = y.shape
    x = np.asarray(x, dtype=np.float64)
    m, p = x.shape
    cells, sizes = designCells(x)
    order = np.argsort(cells, kind="stable").astype(np.int32)
    ptr = np.r_[0, np.cumsum(sizes)].astype(np.int32)
    ptrd = torch.as_tensor(ptr, device=dev)
    ordd = torch.as_tensor(order, device=dev)
    sfd = torch.as_tensor(np.asarray(sizeFactors, dtype=np.float64), device=dev)
    out = {"cooks": torch.empty((n, ld), dtype=F64, device=dev) if want_matrix else None,
           "maxCooks": torch.empty(n, dtype=F64, device=dev), "robustDisp": torch.empty(n, dtype=F64, device=dev)}
    rc = L.b200nb_cooks_dev(_p(y), 0 if y.dtype == torch.int32 else 1, _p(mu), _p(hat), _p(sfd), _p(ptrd), _p(ordd),
                            len(sizes), n, m, p, ld, _p(out["cooks"]), _p(out["maxCooks"]), _p(out["robustDisp"]),
                            _stream())
    _lib.check(rc, "cooks_dev")
    return out


def _median(v):
    """R's median (mean of the two middle order statistics for even length)."""
    s, _ = torch.sort(v)
    k = s.numel()
    return s[k // 2] if k % 2 else 0.5 * (s[k // 2 - 1] + s[k // 2])


def _grid_refit(y, xd, mu, sel, m, prior_mean, prior_sigmasq, usePrior):
    """fitDispGridWrapper (R/wrappers.R:63-83) on the genes flagged in boolean `sel`; returns (idx, alpha)."""
    idx = torch.nonzero(sel).squeeze(1)
    if idx.numel() == 0:
        return idx, None
    grid = np.linspace(np.log(1e-8), np.log(max(10, m)), 20)
    pm = prior_mean[idx] if prior_mean is not None else torch.zeros(idx.numel(), dtype=F64, device=y.device)
    la = D.fit_disp_grid(y[idx].contiguous(), xd, mu[idx].contiguous(), grid, pm, prior_sigmasq, usePrior)["log_alpha"]
    return idx, torch.exp(la)


def getContrast_device(ynz, x, sizeFactors, dispersion, betaMatrix, contrast, betaPriorVar=None, minmu=0.5):
    """getContrast (R/results.R:760-827) on device tensors: fitBeta's maxit = 0 mode (covariance at the given betas).
    ynz: gene-major counts the model was fitted to; betaMatrix: (n, p) log2-scale coefficients (DESeq_device's
    "betaMatrix").  Returns log2FoldChange, lfcSE, stat, pvalue as device tensors."""
    dev = ynz.device
    x = np.asarray(x, dtype=np.float64)
    m, p = x.shape
    contrast = np.asarray(contrast, dtype=np.float64)
    if contrast.shape != (p,):
        raise ValueError("numeric contrast vector should have one element for every element of 'resultsNames(object)'")
    pv = np.full(p, 1e6) if betaPriorVar is None else np.asarray(betaPriorVar, dtype=np.float64)
    lam = torch.as_tensor(1.0 / (LN2 ** 2 * pv), device=dev)
    sfd = torch.as_tensor(np.asarray(sizeFactors, dtype=np.float64), device=dev)
    beta_nat = (betaMatrix.T * LN2).contiguous()                       # (p, n) = column-major n x p, natural log scale
    r = D.fit_beta(ynz, D.x_to_device(x, dev), sfd, dispersion, torch.as_tensor(contrast, device=dev), beta_nat, lam,
                   1e-8, 0, useQR=False, minmu=minmu, want_hat=False, want_mu=False)
    est, se = r["contrast_num"] / LN2, r["contrast_denom"] / LN2
    stat = est / se
    return {"log2FoldChange": est, "lfcSE": se, "stat": stat, "pvalue": 2.0 * torch.special.ndtr(-stat.abs())}


def _optim_fallback(ysrc, xd, sfd, dispersion, fb, beta0, lam, contrast, maxit, minmu, ll, useOptim=True):
    """R/fitNbinomGLMs.R:186-227 on device results: rows with NA coefficients or non-positive variances -- and, with
    useOptim (the reference's default), rows whose IRLS did not converge -- are refitted by b200nb_beta_optim_dev
    (fitNbinomGLMsOptim, :340-407); their coefficients, standard errors (sandwich at the clamped mean, :386-396),
    fitted means and log-likelihoods are overwritten in `fb` / `ll` in place.  The hat diagonals keep the IRLS values,
    as in the reference (:234).  Returns betaConv (n,) bool.  Costs one host sync (the number of such rows)."""
    beta = fb["beta_mat"]                                             # (p, n), natural-log scale
    betaConv = fb["iter"] < maxit
    rowStable = ~torch.isnan(beta).any(dim=0)
    rowVarPositive = ~(fb["beta_var_mat"] <= 0).any(dim=0)
    need = (~rowStable) | (~rowVarPositive)
    if useOptim:
        need = need | (~betaConv)
    rows = torch.nonzero(need).squeeze(1)
    if rows.numel() == 0:
        return betaConv, 0
    usable = rowStable[rows] & ((beta[:, rows] / LN2).abs() < 30.0).all(dim=0)
    start = torch.where(usable[None, :], beta[:, rows], beta0[:, rows]).contiguous()
    start = torch.nan_to_num(start, nan=0.0)
    ysub = ysrc[rows].contiguous()
    asub = dispersion[rows].contiguous()
    o = D.beta_optim(ysub, xd, sfd, asub, lam, start)
    cov = D.fit_beta(ysub, xd, sfd, asub, contrast, o["beta_mat"], lam, 1e-8, 0, minmu=minmu, want_hat=False,
                     want_mu=False)
    l2 = D.nb_loglik(ysub, xd, sfd, asub, o["beta_mat"], want_mu=ll.get("mu") is not None, minmu=minmu)
    fb["beta_mat"][:, rows] = o["beta_mat"]
    fb["beta_var_mat"][:, rows] = cov["beta_var_mat"]
    fb["contrast_num"][rows] = cov["contrast_num"]
    fb["contrast_denom"][rows] = cov["contrast_denom"]
    ll["logLike"][rows] = l2["logLike"]
    if ll.get("mu") is not None:
        ll["mu"][rows] = l2["mu"]
    betaConv = betaConv.clone()
    betaConv[rows] = betaConv[rows] | (o["converged"] != 0)
    return betaConv, int(rows.numel())


def nbinomLRT_device(ynz, x_full, x_reduced, sizeFactors, dispersion, betaTol=1e-8, maxit=100, minmu=0.5,
                     useOptim=True):
    """nbinomLRT (R/core.R:1787-2012) on device tensors: two IRLS fits, LRT statistic 2 (logLike_full - logLike_reduced)
    (R/core.R:1877) and its chi-square p-value.  The log-likelihoods are evaluated at the UNCLAMPED fitted means
    nf * exp(x beta) as the reference does (R/fitNbinomGLMs.R:180-182) by b200nb_nb_loglik_dev -- NOT taken from the IRLS
    kernel's deviance, which is at the minmu-clamped mean and differs by several units when a design cell is all zeros."""
    dev = ynz.device
    sfd = torch.as_tensor(np.asarray(sizeFactors, dtype=np.float64), device=dev)
    res = {}
    for name, x in (("full", x_full), ("reduced", x_reduced)):
        x = np.asarray(x, dtype=np.float64)
        m, p = x.shape
        pr = prep(ynz, x, sizeFactors, minmu=minmu, want_mu=False)
        contrast = torch.zeros(p, dtype=F64, device=dev)
        contrast[0] = 1.0
        lam = torch.full((p,), 1e-6 / LN2 ** 2, dtype=F64, device=dev)
        res[name] = D.fit_beta(ynz, pr["xd"], sfd, dispersion, contrast, pr["beta0"], lam, betaTol, maxit, minmu=minmu,
                               want_hat=(name == "full"), want_mu=False)
        ll = D.nb_loglik(ynz, pr["xd"], sfd, dispersion, res[name]["beta_mat"], want_mu=(name == "full"))
        res[name]["conv"], _ = _optim_fallback(ynz, pr["xd"], sfd, dispersion, res[name], pr["beta0"], lam, contrast,
                                               maxit, minmu, ll, useOptim=useOptim)
        res[name]["logLike"], res[name]["mu"] = ll["logLike"], ll["mu"]
    df = x_full.shape[1] - x_reduced.shape[1]
    stat = 2.0 * (res["full"]["logLike"] - res["reduced"]["logLike"])
    pval = torch.special.gammaincc(torch.tensor(df / 2.0, dtype=F64, device=dev), torch.clamp(stat, min=0.0) / 2.0)
    return {"LRTStatistic": stat, "LRTPvalue": pval, "deviance": -2.0 * res["full"]["logLike"], "df": df,
            "betaMatrix": (res["full"]["beta_mat"] / LN2).T,
            "betaSE": (torch.sqrt(torch.clamp(res["full"]["beta_var_mat"], min=0.0)) / LN2).T,
            "fullBetaConv": res["full"]["conv"], "reducedBetaConv": res["reduced"]["conv"],
            "mu": res["full"]["mu"], "H": res["full"]["hat_diagonals"]}


def _refit_rows(ysub, x, sizeFactors, tr, varLogDispEsts, dispPriorVar, minDisp, kappa_0, dispTol, maxit, betaTol,
                minmu, outlierSD, useOptim=True):
    """The per-gene chain of DESeq_device (GeneEst -> MAP -> Wald) on a subset of rows with the dispersion trend and
    prior already known: what refitWithoutOutliers (R/core.R:2500-2528) does with objectSub.  Fresh tensors, no
    shared workspace: the subset is small."""
    dev = ysub.device
    m, p = x.shape
    maxDisp = float(max(10, m))
    linearMu = modelMatrixGroups(x) == p
    pr = prep(ysub, x, sizeFactors, minDisp=minDisp, minmu=minmu, want_mu=linearMu)
    xd, sfd, bm, alpha0, beta0 = pr["xd"], pr["sfd"], pr["baseMean"], pr["alpha0"], pr["beta0"]
    if linearMu:
        mu = pr["mu_lin"].clone()
    n = ysub.shape[0]
    contrast = torch.zeros(p, dtype=F64, device=dev)
    contrast[0] = 1.0
    lam = torch.full((p,), 1e-6 / LN2 ** 2, dtype=F64, device=dev)
    min_log_alpha = float(np.log(minDisp / 10))
    if not linearMu:
        mu = D.fit_beta(ysub, xd, sfd, alpha0, contrast, beta0, lam, betaTol, maxit, minmu=minmu, want_hat=False,
                        want_mu=True)["mu"]
    la0 = torch.log(alpha0)
    r = D.fit_disp(ysub, xd, mu, la0, la0, 1.0, min_log_alpha, kappa_0, dispTol, maxit, False)
    dge = torch.clamp(torch.exp(r["log_alpha"]), max=maxDisp)
    dge = torch.where(r["last_lp"] < r["initial_lp"] + r["initial_lp"].abs() / 1e6, alpha0, dge)
    conv = (r["iter"] < maxit) & (r["iter"] != 1)
    gi, ga = _grid_refit(ysub, xd, mu, (~conv) & (dge > minDisp * 10), m, None, 1.0, False)
    if ga is not None:
        dge[gi] = ga
    dge = torch.clamp(dge, minDisp, maxDisp)
    dispFit = tr[0] + tr[1] / bm
    logFit = torch.log(dispFit)
    dispInit = torch.where(dge > 0.1 * dispFit, dge, dispFit)
    rm = D.fit_disp(ysub, xd, mu, torch.log(dispInit), logFit, dispPriorVar, min_log_alpha, kappa_0, dispTol, maxit, True)
    dispMAP = torch.exp(rm["log_alpha"])
    gi2, ga2 = _grid_refit(ysub, xd, mu, rm["iter"] >= maxit, m, logFit, dispPriorVar, True)
    if ga2 is not None:
        dispMAP[gi2] = ga2
    dispMAP = torch.clamp(dispMAP, minDisp, maxDisp)
    dispOutlier = torch.log(dge) > logFit + outlierSD * torch.sqrt(varLogDispEsts)
    dispersion = torch.where(dispOutlier, dge, dispMAP)
    fb = D.fit_beta(ysub, xd, sfd, dispersion, contrast, beta0, lam, betaTol, maxit, minmu=minmu, want_hat=False,
                    want_mu=False)
    ll = D.nb_loglik(ysub, xd, sfd, dispersion, fb["beta_mat"], want_mu=False)
    betaConv, _ = _optim_fallback(ysub, xd, sfd, dispersion, fb, beta0, lam, contrast, maxit, minmu, ll, useOptim=useOptim)
    betaMatrix = fb["beta_mat"] / LN2
    betaSE = torch.sqrt(torch.clamp(fb["beta_var_mat"], min=0.0)) / LN2
    stat = betaMatrix / betaSE
    fb["deviance"] = -2.0 * ll["logLike"]
    return {"baseMean": bm, "dispGeneEst": dge, "dispFit": dispFit, "dispMAP": dispMAP, "dispersion": dispersion,
            "dispOutlier": dispOutlier, "dispGeneIter": r["iter"], "dispIter": rm["iter"], "betaMatrix": betaMatrix.T,
            "betaSE": betaSE.T, "WaldStatistic": stat.T, "WaldPvalue": (2.0 * torch.special.ndtr(-stat.abs())).T,
            "betaIter": fb["iter"], "betaConv": betaConv, "deviance": fb["deviance"]}


def DESeq_device(y, x, sizeFactors=None, minDisp=1e-8, kappa_0=1.0, dispTol=1e-6, maxit=100, betaTol=1e-8, minmu=0.5,
                 outlierSD=2.0, minReplicatesForReplace=np.inf, allgather=None, useOptim=True):
    """y: gene-major (N, ld) device tensor of counts (int32 or float64).  Returns a dict of device tensors over the
    rows with a non-zero sum (`idx` maps them back to the N input rows) plus the trend / prior scalars.
    sizeFactors=None estimates them on the device first (median of ratios, R/core.R:535-578; returned under
    "sizeFactors").  minReplicatesForReplace: the reference's default is 7 (replace count outliers by the trimmed mean
    and refit those genes, R/core.R:419-426, 2069-2115, 2484-2565); the default here is Inf = off (what bench.py's
    full_pipeline times).  With a finite value the per-gene results of the refitted rows are overwritten in place
    and "replace" / "replaceable" / "n_replaced" are added.
    allgather: for the gene-sharded multi-GPU run (deseq2_b200/sharded.py::sharded_DESeq_device): a callable that
    concatenates each of a LIST of per-gene 1-D device tensors over the ranks, in rank order (one collective).  It is used at the one point where the
    reference needs every gene (R/parallel.R:25-28): the dispersion trend and the prior variance are then fitted,
    identically on every rank, to all genes' baseMean / dispGeneEst instead of the shard's."""
    dev = y.device
    x = np.asarray(x, dtype=np.float64)
    m, p = x.shape
    estimated = sizeFactors is None
    if estimated:
        sizeFactors = size_factors(y, m)["sizeFactors"].cpu().numpy()
    import os
    import time
    stage_ms = {}
    _t = [time.perf_counter()]
    debug = bool(os.environ.get("B200NB_PIPE_DEBUG"))

    def mark(name):
        if debug:
            torch.cuda.synchronize()
            now = time.perf_counter()
            stage_ms[name] = round((now - _t[0]) * 1e3, 3)
            _t[0] = now

    maxDisp = float(max(10, m))
    linearMu = modelMatrixGroups(x) == p
    # rows with a zero sum are dropped before anything else (R/core.R:706), so every later array is aligned
    idx = torch.nonzero(y[:, :m].sum(dim=1) != 0).squeeze(1)
    ynz = y if idx.numel() == y.shape[0] else y[idx].contiguous()
    pr = prep(ynz, x, sizeFactors, minDisp=minDisp, minmu=minmu, want_mu=linearMu)
    xd, sfd = pr["xd"], pr["sfd"]
    bm = pr["baseMean"]
    alpha0 = pr["alpha0"]
    beta0 = pr["beta0"]
    n = idx.numel()
    contrast = torch.zeros(p, dtype=F64, device=dev)
    contrast[0] = 1.0
    lam = torch.full((p,), 1e-6 / LN2 ** 2, dtype=F64, device=dev)
    min_log_alpha = float(np.log(minDisp / 10))

    # ---- estimateDispersionsGeneEst (R/core.R:657-860)
    if linearMu:
        mu = pr["mu_lin"]
    else:
        nn0 = ynz.shape[0]
        vec = lambda: torch.empty(nn0, dtype=F64, device=dev)
        out0 = {"beta_mat": torch.empty((p, nn0), dtype=F64, device=dev),
                "beta_var_mat": torch.empty((p, nn0), dtype=F64, device=dev), "iter": vec(), "contrast_num": vec(),
                "contrast_denom": vec(), "deviance": vec(), "hat_diagonals": None,
                "mu": _ws("mu_glm", tuple(ynz.shape), F64, dev)}
        mu = D.fit_beta(ynz, xd, sfd, alpha0, contrast, beta0, lam, betaTol, maxit, minmu=minmu, out=out0)["mu"]
    la0 = torch.log(alpha0)
    mark("mu")
    r = D.fit_disp(ynz, xd, mu, la0, la0, 1.0, min_log_alpha, kappa_0, dispTol, maxit, False)
    mark("fit_disp_mle")
    dge = torch.clamp(torch.exp(r["log_alpha"]), max=maxDisp)
    noIncrease = r["last_lp"] < r["initial_lp"] + r["initial_lp"].abs() / 1e6
    dge = torch.where(noIncrease, alpha0, dge)
    conv = (r["iter"] < maxit) & (r["iter"] != 1)
    gi, ga = _grid_refit(ynz, xd, mu, (~conv) & (dge > minDisp * 10), m, None, 1.0, False)
    if ga is not None:
        dge[gi] = ga
    dge = torch.clamp(dge, minDisp, maxDisp)
    n_refit_geneest = int(gi.numel())
    mark("rules+grid_mle")

    # ---- estimateDispersionsFit + dispersionFunction<- + PriorVar (R/core.R:864-940, R/methods.R:142-190, R/core.R:1135-1208)
    # ONE packed collective for the two vectors the global step needs (sharded.PackedGather accepts a list)
    bm_all, dge_all = allgather([bm, dge]) if allgather is not None else (bm, dge)
    tr = trend_fit(bm_all, dge_all, minDisp)
    dispFit = tr[0] + tr[1] / bm
    above = dge_all >= minDisp * 100
    resid = (torch.log(dge_all) - torch.log(tr[0] + tr[1] / bm_all))[above]
    med = _median(resid)
    varLogDispEsts = (1.4826 * _median((resid - med).abs())) ** 2
    if m - p <= 3 and m > p:
        # 2-vs-2 / 3-vs-2 experiments: the reference's Monte-Carlo KL match (R/core.R:1157-1193) is host glue (numpy);
        # it needs the log-dispersion residuals of all genes: one D2H of two doubles per gene
        from .pipeline import estimateDispersionsPriorVar
        dispPriorVar = estimateDispersionsPriorVar(float(varLogDispEsts.item()), m, p, dge_all.cpu().numpy(),
                                                   (tr[0] + tr[1] / bm_all).cpu().numpy(), minDisp)
    elif m > p:
        expVar = torch.special.polygamma(1, torch.tensor((m - p) / 2.0, dtype=F64, device=dev))
        dispPriorVar = float(torch.clamp(varLogDispEsts - expVar, min=0.25).item())
    else:
        dispPriorVar = float(varLogDispEsts.item())
    status = tr[2].item()
    if status != 0:
        raise FloatingPointError(f"parametric dispersion fit failed on device (status {int(status)})")
    mark("trend+priorvar")

    # ---- estimateDispersionsMAP (R/core.R:943-1131)
    dispInit = torch.where(dge > 0.1 * dispFit, dge, dispFit)
    logFit = torch.log(dispFit)
    rm = D.fit_disp(ynz, xd, mu, torch.log(dispInit), logFit, dispPriorVar, min_log_alpha, kappa_0, dispTol, maxit, True)
    mark("fit_disp_map")
    dispMAP = torch.exp(rm["log_alpha"])
    gi2, ga2 = _grid_refit(ynz, xd, mu, rm["iter"] >= maxit, m, logFit, dispPriorVar, True)
    if ga2 is not None:
        dispMAP[gi2] = ga2
    dispMAP = torch.clamp(dispMAP, minDisp, maxDisp)
    dispOutlier = torch.log(dge) > logFit + outlierSD * torch.sqrt(varLogDispEsts)
    dispersion = torch.where(dispOutlier, dge, dispMAP)
    mark("rules+grid_map")

    # ---- nbinomWaldTest (R/core.R:1332-1565) via fitNbinomGLMs (R/fitNbinomGLMs.R:29-236)
    nn, ldd = ynz.shape
    outb = {"beta_mat": torch.empty((p, nn), dtype=F64, device=dev), "beta_var_mat": torch.empty((p, nn), dtype=F64, device=dev),
            "iter": torch.empty(nn, dtype=F64, device=dev), "contrast_num": torch.empty(nn, dtype=F64, device=dev),
            "contrast_denom": torch.empty(nn, dtype=F64, device=dev), "deviance": torch.empty(nn, dtype=F64, device=dev),
            "hat_diagonals": _ws("H", (nn, ldd), F64, dev), "mu": None}
    fb = D.fit_beta(ynz, xd, sfd, dispersion, contrast, beta0, lam, betaTol, maxit, minmu=minmu, out=outb)
    mark("fit_beta")
    # Cook's distances and the reported deviance use the UNCLAMPED fitted mean nf * exp(x beta) and the log-likelihood
    # at it, which the reference recomputes in R right after the native call (R/fitNbinomGLMs.R:180-182): one small
    # kernel (b200nb_nb_loglik_dev); the IRLS kernel's fused mean / deviance are at the minmu clamp.
    ll = D.nb_loglik(ynz, xd, sfd, dispersion, fb["beta_mat"], out_mu=_ws("mu_cooks", (nn, ldd), F64, dev))
    # rows the IRLS could not fit go to the box-constrained maximiser (R/fitNbinomGLMs.R:203-227, useOptim = TRUE)
    betaConv, n_optim = _optim_fallback(ynz, xd, sfd, dispersion, fb, beta0, lam, contrast, maxit, minmu, ll,
                                        useOptim=useOptim)
    mark("loglik+optim")
    betaMatrix = fb["beta_mat"] / LN2                      # (p, n)
    betaSE = torch.sqrt(torch.clamp(fb["beta_var_mat"], min=0.0)) / LN2
    stat = betaMatrix / betaSE
    pval = 2.0 * torch.special.ndtr(-stat.abs())
    from .pipeline import nOrMoreInCell
    do_replace = bool(np.isfinite(minReplicatesForReplace)) and bool(nOrMoreInCell(x, minReplicatesForReplace).any())
    mu_cooks = ll["mu"]
    ck = cooks(ynz, mu_cooks, fb["hat_diagonals"], x, sizeFactors, want_matrix=do_replace)   # R/core.R:1457-1460
    mark("wald_stats+cooks")
    res = {"stage_ms": stage_ms, "maxCooks": ck["maxCooks"], "idx": idx, "baseMean": bm, "dispGeneEst": dge, "dispFit": dispFit, "dispMAP": dispMAP,
            "dispersion": dispersion, "dispOutlier": dispOutlier, "dispGeneIter": r["iter"], "dispIter": rm["iter"],
            "betaMatrix": betaMatrix.T, "betaSE": betaSE.T, "WaldStatistic": stat.T, "WaldPvalue": pval.T,
            "betaIter": fb["iter"], "betaConv": betaConv, "n_optim": n_optim, "deviance": -2.0 * ll["logLike"], "mu": mu_cooks,
            "H": fb["hat_diagonals"], "trendCoefs": tr[:2], "varLogDispEsts": varLogDispEsts,
            "dispPriorVar": dispPriorVar, "n_refit_geneest": n_refit_geneest, "n_refit_map": int(gi2.numel()),
            "n_input_rows": y.shape[0], "sizeFactors": np.asarray(sizeFactors, dtype=np.float64)}
    if do_replace and m > p:
        _replace_and_refit(res, ynz, ck["cooks"], x, sizeFactors, sfd, tr, varLogDispEsts, dispPriorVar,
                           minReplicatesForReplace, minDisp, kappa_0, dispTol, maxit, betaTol, minmu, outlierSD,
                           useOptim=useOptim)
    return res


def _replace_and_refit(res, ynz, cooksm, x, sizeFactors, sfd, tr, varLogDispEsts, dispPriorVar, minReplicates, minDisp,
                       kappa_0, dispTol, maxit, betaTol, minmu, outlierSD, trim=0.2, useOptim=True):
    """replaceOutliers + refitWithoutOutliers (R/core.R:2069-2115, 2484-2565) on the device results `res` (in place).
    Only the flagged rows move: their counts are gathered, the outlying entries of replaceable samples become
    as.integer(trimmed mean of the normalised counts * size factor), and the rows go through _refit_rows."""
    from scipy import stats as _st
    from .pipeline import nOrMoreInCell
    dev = ynz.device
    m, p = x.shape
    cutoff = float(_st.f.ppf(0.99, p, m - p))
    replaceable = torch.as_tensor(nOrMoreInCell(x, minReplicates), device=dev)
    over = cooksm[:, :m] > cutoff                                   # NaN compares false, like NA in which()/any()
    replace = over.any(dim=1)
    rows = torch.nonzero(replace).squeeze(1)
    res["replace"], res["replaceable"], res["n_replaced"] = replace, replaceable, int(rows.numel())
    if rows.numel() == 0:
        return
    ys = ynz[rows][:, :m].to(F64)
    srt, _ = torch.sort(ys / sfd[None, :], dim=1)
    lo = int(np.floor(m * trim))
    tbm = srt[:, lo:m - lo].mean(dim=1)
    replacement = torch.trunc(tbm[:, None] * sfd[None, :])
    new = torch.where(over[rows] & replaceable[None, :], replacement, ys)
    res["baseMean"] = res["baseMean"].clone()
    res["baseMean"][rows] = (new / sfd[None, :]).mean(dim=1)        # getBaseMeansAndVariances on the new counts
    newAllZero = new.sum(dim=1) == 0
    keep = ~newAllZero
    rr = rows[keep]
    if rr.numel() > 0:
        ysub = torch.zeros((rr.numel(), ynz.shape[1]), dtype=F64, device=dev)
        ysub[:, :m] = new[keep]
        sub = _refit_rows(ysub, x, sizeFactors, tr, varLogDispEsts, dispPriorVar, minDisp, kappa_0, dispTol, maxit,
                          betaTol, minmu, outlierSD, useOptim=useOptim)
        for k, v in sub.items():
            if k == "baseMean":
                continue
            res[k] = res[k].clone()
            res[k][rr] = v.to(res[k].dtype)
        for k in ("betaMatrix", "betaSE", "WaldStatistic", "WaldPvalue", "deviance"):
            res[k][rows[newAllZero]] = float("nan")
        if bool(replaceable.all()):
            res["maxCooks"] = torch.full_like(res["maxCooks"], float("nan"))
        else:                                   # replaceCooks[, replaceable] <- 0, then recordMaxCooks (:2537-2543)
            from .pipeline import designCells
            cells, sizes = designCells(x)
            forCooks = torch.as_tensor(sizes[cells] >= 3, device=dev)
            rc = cooksm[:, :m].clone()
            rc[:, replaceable] = 0.0
            res["maxCooks"] = (rc[:, forCooks].max(dim=1).values if bool(forCooks.any())
                               else torch.full_like(res["maxCooks"], float("nan")))
