"""Host-side glue of DESeq()'s default Wald path around the three native calls.

The reference keeps all of this in R (it "stays R" in a real drop-in); R is not in this
image, so the steps that prepare the inputs of fitDisp / fitBeta and consume their outputs are restated here in
numpy, one function per R function, same names, so the engine can be driven and measured end to end:

  estimateSizeFactorsForMatrix   R/core.R:535-578   (type="ratio")
  getBaseMeansAndVariances       R/core.R:2138-2157
  linearModelMu / roughDispEstimate / momentsDispEstimate   R/core.R:2454-2459, 2422-2437, 2439-2448
  estimateDispersionsGeneEst     R/core.R:657-860
  parametricDispersionFit        R/core.R:2166-2189  (Gamma GLM, identity link, glm.fit IRLS restated)
  dispersionFunction<-           R/methods.R:142-190 (dispFit, varLogDispEsts = mad^2)
  estimateDispersionsPriorVar    R/core.R:1135-1208  (incl. the Monte-Carlo branch for 1..3 residual df, own RNG)
  estimateDispersionsMAP         R/core.R:943-1131
  fitNbinomGLMs / fitNbinomGLMsOptim   R/fitNbinomGLMs.R:29-236, 340-407 (L-BFGS-B fallback, useOptim=TRUE by default as in the reference)
  nbinomWaldTest                 R/core.R:1332-1565  (betas, SEs, Wald statistic and p-value)
  nbinomLRT                      R/core.R:1787-2012  (full vs reduced fit, 2 (l_full - l_reduced), chi-square p-value)
  getContrast                    R/results.R:760-827 (numeric contrasts through fitBeta's maxit = 0 mode)
  replaceOutliers / refit        R/core.R:2069-2115, 2484-2565 (opt-in: DESeq(minReplicatesForReplace=7))
  robustMethodOfMomentsDisp / trimmedCellVariance / calculateCooksDistance / recordMaxCooks   R/core.R:2277-2359
  fitGLMsWithPrior / estimateBetaPriorVar / Hmisc.wtd.quantile / addAllContrasts / averagePriorsOverLevels /
  makeExpandedModelMatrix (additive factor designs)   R/fitNbinomGLMs.R:242-337, R/core.R:1601-1689, 2762-2803, R/expanded.R
  DESeq                          R/core.R:280-432    (test="Wald", fitType="parametric", betaPrior=FALSE)

`engine` is any object with fitDisp / fitDispGrid / fitBeta taking the reference's argument names
(deseq2_b200.wrappers is the product engine and the default; tests and the bench's CPU baseline pass the
oracle).  Not restated (out of scope): local / mean trend fits,
observation weights in the glue, results()/lfcShrink().
"""
from __future__ import annotations

import numpy as np
from scipy import special as _sp

from . import wrappers as _default_engine
from .wrappers import fitBetaWrapper, fitDispGridWrapper, fitDispWrapper

LN2 = np.log(2.0)


def estimateSizeFactorsForMatrix(counts):
    """R/core.R:535-578, type='ratio', locfunc=median."""
    counts = np.asarray(counts, dtype=np.float64)
    with np.errstate(divide="ignore"):
        logc = np.log(counts)
    loggeomeans = logc.mean(axis=1)
    if np.all(np.isinf(loggeomeans)):
        raise ValueError("every gene contains at least one zero, cannot compute log geometric means")
    ok = np.isfinite(loggeomeans)
    sf = np.empty(counts.shape[1])
    for j in range(counts.shape[1]):
        sel = ok & (counts[:, j] > 0)
        sf[j] = np.exp(np.median(logc[sel, j] - loggeomeans[sel]))
    return sf


def getBaseMeansAndVariances(counts, sizeFactors):
    """R/core.R:2138-2157 (no weights)."""
    norm = np.asarray(counts, dtype=np.float64) / sizeFactors[None, :]
    return {"baseMean": norm.mean(axis=1), "baseVar": norm.var(axis=1, ddof=1),
            "allZero": np.asarray(counts).sum(axis=1) == 0}


def linearModelMu(y, x):
    """R/core.R:2454-2459: (y Q)(x R^-1)'."""
    Q, R = np.linalg.qr(x)
    Rinv = np.linalg.inv(R)
    return (y @ Q) @ (x @ Rinv).T


def roughDispEstimate(y, x):
    """R/core.R:2422-2437."""
    mu = np.maximum(1.0, linearModelMu(y, x))
    m, p = x.shape
    est = (((y - mu) ** 2 - mu) / mu ** 2).sum(axis=1) / (m - p)
    return np.maximum(est, 0.0)


def momentsDispEstimate(baseMean, baseVar, sizeFactors):
    """R/core.R:2439-2448."""
    xim = np.mean(1.0 / sizeFactors)
    return (baseVar - xim * baseMean) / baseMean ** 2


def modelMatrixGroups(x):
    """R/core.R:2450-2452: number of distinct rows of the design."""
    return len({tuple(r) for r in np.asarray(x)})


def _stirling_tail(x):
    xi = 1.0 / x
    x2 = xi * xi
    return xi * (1.0 / 12 - x2 * (1.0 / 360 - x2 * (1.0 / 1260 - x2 * (1.0 / 1680 - x2 * (1.0 / 1188)))))


def _lgamma_diff(y, r):
    """lgamma(y + r) - lgamma(r) without the cancellation of two ~r log r numbers (r = 1/alpha reaches 1e8, where the
    direct difference loses ~1e-7 per sample): for r >= 10 the Stirling forms are subtracted analytically."""
    r = np.broadcast_to(r, y.shape)
    big = r >= 10.0
    out = np.empty(y.shape)
    rb, yb = r[big], y[big]
    out[big] = yb * np.log(yb + rb) + (rb - 0.5) * np.log1p(yb / rb) - yb + _stirling_tail(yb + rb) - _stirling_tail(rb)
    out[~big] = _sp.gammaln(y[~big] + r[~big]) - _sp.gammaln(r[~big])
    return out


def nbinomLogLike(counts, mu, disp):
    """R/core.R:2208-2217 without weights: rowSums(dnbinom(counts, mu = mu, size = 1/disp, log = TRUE)).  Evaluated in
    the cancellation-free direct form (the same as the device kernels), with dnbinom_mu's point mass at mu = 0."""
    y = np.asarray(counts, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    alpha = np.asarray(disp, dtype=np.float64)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (_lgamma_diff(y, 1.0 / alpha) - _sp.gammaln(y + 1) + _sp.xlogy(y, mu * alpha)
             - (y + 1.0 / alpha) * np.log1p(mu * alpha))
    t = np.where(mu == 0, np.where(y == 0, 0.0, -np.inf), t)
    return t.sum(axis=1)


def estimateDispersionsGeneEst(counts, sizeFactors, x, engine=None, minDisp=1e-8, kappa_0=1.0, dispTol=1e-6,
                               maxit=100, useCR=True, weightThreshold=1e-2, minmu=0.5, mv=None):
    """R/core.R:657-860 for niter=1, no weights.  `counts` holds only the non-all-zero rows."""
    engine = engine or _default_engine
    counts = np.asarray(counts)
    n, m = counts.shape
    p = x.shape[1]
    if mv is None:
        mv = getBaseMeansAndVariances(counts, sizeFactors)
    norm = counts / sizeFactors[None, :]
    roughDisp = roughDispEstimate(norm, x)
    momentsDisp = momentsDispEstimate(mv["baseMean"], mv["baseVar"], sizeFactors)
    alpha_hat = np.minimum(roughDisp, momentsDisp)
    maxDisp = max(10, m)
    alpha_hat = alpha_init = np.minimum(np.maximum(minDisp, alpha_hat), maxDisp)
    linearMu = modelMatrixGroups(x) == p
    nf = np.broadcast_to(sizeFactors[None, :], (n, m))
    if not linearMu:
        fit = fitNbinomGLMs(counts, nf, x, alpha_hat, engine=engine)
        fitMu = fit["mu"]
    else:
        fitMu = linearModelMu(norm, x) * nf
    fitMu = np.maximum(fitMu, minmu)
    dispRes = fitDispWrapper(ySEXP=counts, xSEXP=x, mu_hatSEXP=fitMu, log_alphaSEXP=np.log(alpha_hat),
                             log_alpha_prior_meanSEXP=np.log(alpha_hat), log_alpha_prior_sigmasqSEXP=1.0,
                             min_log_alphaSEXP=np.log(minDisp / 10), kappa_0SEXP=kappa_0, tolSEXP=dispTol,
                             maxitSEXP=maxit, usePriorSEXP=False, weightsSEXP=None, useWeightsSEXP=False,
                             weightThresholdSEXP=weightThreshold, useCRSEXP=useCR, engine=engine)
    dispIter = dispRes["iter"]
    dispGeneEst = np.minimum(np.exp(dispRes["log_alpha"]), maxDisp)
    noIncrease = dispRes["last_lp"] < dispRes["initial_lp"] + np.abs(dispRes["initial_lp"]) / 1e6
    dispGeneEst[noIncrease] = alpha_init[noIncrease]
    dispGeneEstConv = (dispIter < maxit) & ~(dispIter == 1)
    refitDisp = ~dispGeneEstConv & (dispGeneEst > minDisp * 10)
    if refitDisp.sum() > 0:
        dispGrid = fitDispGridWrapper(y=counts[refitDisp], x=x, mu=fitMu[refitDisp],
                                      logAlphaPriorMean=np.zeros(int(refitDisp.sum())), logAlphaPriorSigmaSq=1.0,
                                      usePrior=False, weightsSEXP=None, useWeightsSEXP=False,
                                      weightThresholdSEXP=weightThreshold, useCRSEXP=useCR, engine=engine)
        dispGeneEst[refitDisp] = dispGrid
    dispGeneEst = np.minimum(np.maximum(dispGeneEst, minDisp), maxDisp)
    return {"dispGeneEst": dispGeneEst, "dispGeneIter": dispIter, "mu": fitMu, "baseMean": mv["baseMean"],
            "baseVar": mv["baseVar"], "n_refit": int(refitDisp.sum()), "dispRes": dispRes}


def _gamma_glm_identity(y, xinv, start, maxit=25, epsilon=1e-8):
    """glm(y ~ I(1/means), family=Gamma(link='identity'), start=start) via glm.fit's IRLS:
    eta = mu, working weights 1/mu^2, working response y; converged when |dev-devold|/(|dev|+0.1) < epsilon."""
    X = np.c_[np.ones_like(xinv), xinv]
    coef = np.asarray(start, dtype=np.float64)
    mu = X @ coef
    dev_old = 2.0 * np.sum(-np.log(y / mu) + (y - mu) / mu)
    converged = False
    for _ in range(maxit):
        w = 1.0 / mu ** 2
        XtW = X.T * w
        coef = np.linalg.solve(XtW @ X, XtW @ y)
        mu = X @ coef
        if np.any(mu <= 0):
            raise FloatingPointError("parametric dispersion fit failed")
        dev = 2.0 * np.sum(-np.log(y / mu) + (y - mu) / mu)
        if abs(dev - dev_old) / (abs(dev) + 0.1) < epsilon:
            converged = True
            break
        dev_old = dev
    return coef, converged


def parametricDispersionFit(means, disps):
    """R/core.R:2166-2189.  Returns (asymptDisp, extraPois)."""
    coefs = np.array([0.1, 1.0])
    it = 0
    while True:
        residuals = disps / (coefs[0] + coefs[1] / means)
        good = (residuals > 1e-4) & (residuals < 15)
        oldcoefs = coefs
        coefs, converged = _gamma_glm_identity(disps[good], 1.0 / means[good], coefs)
        if not np.all(coefs > 0):
            raise FloatingPointError("parametric dispersion fit failed")
        if (np.sum(np.log(coefs / oldcoefs) ** 2) < 1e-6) and converged:
            break
        it += 1
        if it > 10:
            raise FloatingPointError("dispersion fit did not converge")
    return coefs


def _mad(x):
    """stats::mad: 1.4826 * median(|x - median(x)|)."""
    return 1.4826 * np.median(np.abs(x - np.median(x)))


def estimateDispersionsFit(dispGeneEst, baseMean, minDisp=1e-8):
    """R/core.R:864-940 (fitType='parametric') + dispersionFunction<- (R/methods.R:142-190)."""
    useForFit = dispGeneEst > 100 * minDisp
    if useForFit.sum() == 0:
        raise ValueError("all gene-wise dispersion estimates are within 2 orders of magnitude from the minimum value")
    coefs = parametricDispersionFit(baseMean[useForFit], dispGeneEst[useForFit])
    dispFit = coefs[0] + coefs[1] / baseMean
    aboveMinDisp = dispGeneEst >= minDisp * 100
    dispResiduals = np.log(dispGeneEst) - np.log(dispFit)
    varLogDispEsts = _mad(dispResiduals[aboveMinDisp]) ** 2
    return {"dispFit": dispFit, "coefs": coefs, "varLogDispEsts": varLogDispEsts}


def _loess_quadratic(xg, yg, xnew, span):
    """stats::loess(y ~ x, span = span) with its defaults degree = 2, family = "gaussian", evaluated directly at xnew
    (R's default surface = "interpolate" evaluates the same local fits at kd-tree vertices and blends them; on a smooth
    200-point curve the two agree far inside the Monte-Carlo noise of the curve itself): local quadratic least squares
    over the q = floor(n * span + 1e-5) nearest points with tricube weights."""
    n = len(xg)
    q = max(int(np.floor(n * span + 1e-5)), 3)
    out = np.empty(len(xnew))
    for i, x0 in enumerate(xnew):
        d = np.abs(xg - x0)
        idx = np.argpartition(d, q - 1)[:q]
        h = d[idx].max()
        w = (1.0 - (d[idx] / h) ** 3) ** 3 if h > 0 else np.ones(q)
        w = np.maximum(w, 0.0)
        X = np.c_[np.ones(q), xg[idx] - x0, (xg[idx] - x0) ** 2]
        sw = np.sqrt(w)
        coef, *_ = np.linalg.lstsq(X * sw[:, None], yg[idx] * sw, rcond=None)
        out[i] = coef[0]
    return out


def estimateDispersionsPriorVar(varLogDispEsts, m, p, dispGeneEst=None, dispFit=None, minDisp=1e-8):
    """R/core.R:1135-1208.  m - p > 3: MAD-based variance minus the sampling variance trigamma((m-p)/2), floor 0.25.
    1 <= m - p <= 3 (2-vs-2, 3-vs-2 experiments, :1157-1193): the reference matches the histogram of the log
    dispersion residuals against simulated log(chisq_{m-p}/(m-p)) + N(0, v) on a grid of v by KL divergence, smooths
    the curve with loess(span = .2) and takes the argmin; needs dispGeneEst and dispFit.  The reference fixes R's RNG
    (set.seed(2): Mersenne-Twister + inversion) for reproducibility; R's generator is not restated here, numpy's
    PCG64(2) plays its role, so the value agrees with R's within the Monte-Carlo noise of 1e4 draws per grid point
    (a few grid steps of 0.008), not bit for bit -- host glue that stays R in a real drop-in.
    m == p: no sampling variance is subtracted."""
    if m - p <= 3 and m > p:
        if dispGeneEst is None or dispFit is None:
            raise ValueError("residual df <= 3: the Monte-Carlo branch needs dispGeneEst and dispFit")
        dispGeneEst = np.asarray(dispGeneEst, dtype=np.float64)
        resid = np.log(dispGeneEst) - np.log(np.asarray(dispFit, dtype=np.float64))
        above = dispGeneEst >= minDisp * 100
        if above.sum() == 0:
            raise ValueError("no data found which is greater than minDisp")
        rng = np.random.Generator(np.random.PCG64(2))
        brks = np.arange(-20, 21) / 2.0
        obs = resid[above]
        obs = obs[(obs > brks[0]) & (obs < brks[-1])]
        obsDens = np.histogram(obs, bins=brks)[0] / (max(len(obs), 1) * 0.5)
        grid = np.linspace(0.0, 8.0, 200)
        kl = np.empty(len(grid))
        df = m - p
        for i, v in enumerate(grid):
            r = np.log(rng.chisquare(df, 10000)) + rng.normal(0.0, np.sqrt(v), 10000) - np.log(df)
            r = r[(r > brks[0]) & (r < brks[-1])]
            rDens = np.histogram(r, bins=brks)[0] / (max(len(r), 1) * 0.5)
            z = np.r_[obsDens, rDens]
            small = z[z > 0].min()
            kl[i] = np.sum(obsDens * (np.log(obsDens + small) - np.log(rDens + small)))
        fine = np.linspace(0.0, 8.0, 1000)
        fitted = _loess_quadratic(grid, kl, fine, 0.2)
        return float(max(fine[int(np.argmin(fitted))], 0.25))
    if m > p:
        expVarLogDisp = _sp.polygamma(1, (m - p) / 2.0)
        return float(max(varLogDispEsts - expVarLogDisp, 0.25))
    return float(varLogDispEsts)


def estimateDispersionsMAP(counts, x, mu, dispGeneEst, dispFit, dispPriorVar, varLogDispEsts, engine=None,
                           outlierSD=2.0, minDisp=1e-8, kappa_0=1.0, dispTol=1e-6, maxit=100, useCR=True,
                           weightThreshold=1e-2):
    """R/core.R:943-1131, type='DESeq2', no weights."""
    engine = engine or _default_engine
    n, m = np.asarray(counts).shape
    dispInit = np.where(dispGeneEst > 0.1 * dispFit, dispGeneEst, dispFit)
    dispResMAP = fitDispWrapper(ySEXP=counts, xSEXP=x, mu_hatSEXP=mu, log_alphaSEXP=np.log(dispInit),
                                log_alpha_prior_meanSEXP=np.log(dispFit), log_alpha_prior_sigmasqSEXP=dispPriorVar,
                                min_log_alphaSEXP=np.log(minDisp / 10), kappa_0SEXP=kappa_0, tolSEXP=dispTol,
                                maxitSEXP=maxit, usePriorSEXP=True, weightsSEXP=None, useWeightsSEXP=False,
                                weightThresholdSEXP=weightThreshold, useCRSEXP=useCR, engine=engine)
    dispMAP = np.exp(dispResMAP["log_alpha"])
    dispIter = dispResMAP["iter"]
    dispConv = dispIter < maxit
    refitDisp = ~dispConv
    if refitDisp.sum() > 0:
        dispGrid = fitDispGridWrapper(y=np.asarray(counts)[refitDisp], x=x, mu=mu[refitDisp],
                                      logAlphaPriorMean=np.log(dispFit)[refitDisp], logAlphaPriorSigmaSq=dispPriorVar,
                                      usePrior=True, weightsSEXP=None, useWeightsSEXP=False,
                                      weightThresholdSEXP=weightThreshold, useCRSEXP=True, engine=engine)
        dispMAP[refitDisp] = dispGrid
    maxDisp = max(10, m)
    dispMAP = np.minimum(np.maximum(dispMAP, minDisp), maxDisp)
    dispersionFinal = dispMAP.copy()
    dispOutlier = np.log(dispGeneEst) > np.log(dispFit) + outlierSD * np.sqrt(varLogDispEsts)
    dispersionFinal[dispOutlier] = dispGeneEst[dispOutlier]
    return {"dispersion": dispersionFinal, "dispIter": dispIter, "dispOutlier": dispOutlier, "dispMAP": dispMAP,
            "n_refit": int(refitDisp.sum()), "dispRes": dispResMAP}


def fitNbinomGLMsOptim(counts, nf, x, lambda_, rowsForOptim, rowStable, alpha_hat, betaMatrix, betaSE, betaConv,
                       beta_mat, mu, logLike, minmu=0.5):
    """R/fitNbinomGLMs.R:340-407: L-BFGS-B on the penalised NB log-likelihood (log2 scale, box [-30, 30]) for the rows
    the IRLS left unconverged / unstable; standard errors from the sandwich at the optimum.  No weights."""
    from scipy import optimize as _so
    counts = np.asarray(counts, dtype=np.float64)
    lam = np.asarray(lambda_, dtype=np.float64)
    lamNat = lam / LN2 ** 2
    large = 30.0
    betaMatrix, betaSE, betaConv, mu, logLike = (betaMatrix.copy(), betaSE.copy(), betaConv.copy(), mu.copy(),
                                                 logLike.copy())
    for row in rowsForOptim:
        betaRow = betaMatrix[row] if (rowStable[row] and np.all(np.abs(betaMatrix[row]) < large)) else beta_mat[row]
        nfr, k, alpha = nf[row], counts[row], alpha_hat[row]

        def objective(p_):
            mu_row = nfr * 2.0 ** (x @ p_)
            ll = nbinomLogLike(k[None, :], mu_row[None, :], np.array([alpha]))[0]
            logPrior = np.sum(-0.5 * np.log(2 * np.pi / lam) - 0.5 * lam * p_ ** 2)   # dnorm(p, 0, sqrt(1/lambda), log=TRUE)
            v = -(ll + logPrior)
            return v if np.isfinite(v) else 1e300

        o = _so.minimize(objective, np.asarray(betaRow, dtype=np.float64), method="L-BFGS-B",
                         bounds=[(-large, large)] * x.shape[1])
        if o.status == 0:
            betaConv[row] = True
        betaMatrix[row] = o.x
        mu_row = nfr * 2.0 ** (x @ o.x)
        mu[row] = mu_row
        mu_row = np.maximum(mu_row, minmu)
        w = 1.0 / (1.0 / mu_row + alpha)
        xtwx = (x.T * w) @ x
        inv = np.linalg.inv(xtwx + np.diag(lamNat))
        sigma = inv @ xtwx @ inv
        betaSE[row] = np.sqrt(np.maximum(np.diag(sigma), 0.0)) / LN2
        logLike[row] = nbinomLogLike(k[None, :], mu_row[None, :], np.array([alpha]))[0]
    return {"betaMatrix": betaMatrix, "betaSE": betaSE, "betaConv": betaConv, "mu": mu, "logLike": logLike}


def fitNbinomGLMs(counts, nf, x, alpha_hat, lambda_=None, engine=None, betaTol=1e-8, maxit=100, useQR=True,
                  minmu=0.5, useOptim=True, forceOptim=False):
    """R/fitNbinomGLMs.R:29-236 without weights.  As in the reference (:203-227) rows with NA coefficients or
    non-positive variances ALWAYS go to the L-BFGS-B fallback, and with useOptim=TRUE (the reference's default, :30)
    so do the rows whose IRLS did not converge (iter == maxit, which includes the |beta| > 30 sentinel).  Parity tests
    that need the engine's raw output pass useOptim=False."""
    engine = engine or _default_engine
    counts = np.asarray(counts)
    n, m = counts.shape
    p = x.shape[1]
    if lambda_ is None:
        lambda_ = np.full(p, 1e-6)
    norm = counts / nf
    if np.linalg.matrix_rank(x) == p:
        Q, R = np.linalg.qr(x)
        ylog = np.log(norm + 0.1).T
        beta_mat = np.linalg.solve(R, Q.T @ ylog).T
    else:
        beta_mat = np.zeros((n, p))
        beta_mat[:, 0] = np.log(norm.mean(axis=1))
    lambdaNatLogScale = np.asarray(lambda_, dtype=np.float64) / LN2 ** 2
    betaRes = fitBetaWrapper(ySEXP=counts, xSEXP=x, nfSEXP=nf, alpha_hatSEXP=alpha_hat, beta_matSEXP=beta_mat,
                             lambdaSEXP=lambdaNatLogScale, weightsSEXP=None, useWeightsSEXP=False, tolSEXP=betaTol,
                             maxitSEXP=maxit, useQRSEXP=useQR, minmuSEXP=minmu, engine=engine)
    mu = nf * np.exp(betaRes["beta_mat"] @ x.T)
    logLike = nbinomLogLike(counts, mu, alpha_hat)
    betaConv = betaRes["iter"] < maxit
    betaMatrix = betaRes["beta_mat"] / LN2
    betaSE = np.sqrt(np.maximum(betaRes["beta_var_mat"], 0.0)) / LN2
    rowStable = ~np.isnan(betaRes["beta_mat"]).any(axis=1)
    rowVarPositive = ~(betaRes["beta_var_mat"] <= 0).any(axis=1)
    rows = np.flatnonzero((~betaConv | ~rowStable | ~rowVarPositive) if useOptim else (~rowStable | ~rowVarPositive))
    if forceOptim:
        rows = np.arange(n)
    if True:
        if rows.size:
            o = fitNbinomGLMsOptim(counts, np.broadcast_to(nf, counts.shape), x, lambda_, rows, rowStable, alpha_hat,
                                   betaMatrix, betaSE, betaConv, beta_mat / LN2, mu, logLike, minmu=minmu)
            betaMatrix, betaSE, betaConv, mu, logLike = (o["betaMatrix"], o["betaSE"], o["betaConv"], o["mu"],
                                                         o["logLike"])
    return {"logLike": logLike, "betaConv": betaConv, "betaMatrix": betaMatrix, "betaSE": betaSE, "mu": mu,
            "betaIter": betaRes["iter"], "hat_diagonals": betaRes["hat_diagonals"], "betaRes": betaRes}


def nbinomWaldTest(counts, nf, x, dispersion, engine=None, betaTol=1e-8, maxit=100, useQR=True, minmu=0.5,
                   useOptim=True):
    """R/core.R:1332-1565, betaPrior=FALSE, useT=FALSE; Cook's distances are not restated."""
    fit = fitNbinomGLMs(counts, nf, x, dispersion, engine=engine, betaTol=betaTol, maxit=maxit, useQR=useQR,
                        minmu=minmu, useOptim=useOptim)
    with np.errstate(divide="ignore", invalid="ignore"):
        WaldStatistic = fit["betaMatrix"] / fit["betaSE"]
    WaldPvalue = 2.0 * _sp.ndtr(-np.abs(WaldStatistic))
    fit.update({"WaldStatistic": WaldStatistic, "WaldPvalue": WaldPvalue, "deviance": -2.0 * fit["logLike"]})
    return fit


def _r_trimmed_mean(x, trim):
    """R's mean(x, trim=) along the last axis: drop floor(n*trim) order statistics from each end."""
    x = np.sort(np.asarray(x, dtype=np.float64), axis=-1)
    n = x.shape[-1]
    if trim >= 0.5:
        return np.median(x, axis=-1)
    lo = int(np.floor(n * trim))
    return x[..., lo:n - lo].mean(axis=-1)


def _trimfn(n):
    """as.integer(cut(n, breaks=c(0, 3.5, 23.5, Inf))) - 1  (R/core.R:2306)."""
    return 0 if n <= 3.5 else (1 if n <= 23.5 else 2)


_TRIMRATIO = (1.0 / 3.0, 1.0 / 4.0, 1.0 / 8.0)
_SCALE_C = (2.04, 1.86, 1.51)


def designCells(x):
    """Cell (distinct design row) id per sample, in order of first appearance, and the cell sizes."""
    seen, ids = {}, []
    for r in np.asarray(x):
        ids.append(seen.setdefault(tuple(r), len(seen)))
    ids = np.asarray(ids)
    return ids, np.bincount(ids)


def trimmedCellVariance(cnts, cells):
    """R/core.R:2302-2325."""
    levels = np.unique(cells)
    cellMeans = np.empty((cnts.shape[0], len(levels)))
    for k, lvl in enumerate(levels):
        sel = cells == lvl
        cellMeans[:, k] = _r_trimmed_mean(cnts[:, sel], _TRIMRATIO[_trimfn(sel.sum())])
    qmat = cellMeans[:, np.searchsorted(levels, cells)]
    sqerror = (cnts - qmat) ** 2
    varEst = np.empty_like(cellMeans)
    for k, lvl in enumerate(levels):
        sel = cells == lvl
        t = _trimfn(sel.sum())
        varEst[:, k] = _SCALE_C[t] * _r_trimmed_mean(sqerror[:, sel], _TRIMRATIO[t])
    return varEst.max(axis=1)


def trimmedVariance(x):
    """R/core.R:2327-2332."""
    rm = _r_trimmed_mean(x, 1.0 / 8.0)
    return 1.51 * _r_trimmed_mean((x - rm[:, None]) ** 2, 1.0 / 8.0)


def robustMethodOfMomentsDisp(counts, sizeFactors, x):
    """R/core.R:2277-2300."""
    cnts = np.asarray(counts, dtype=np.float64) / sizeFactors[None, :]
    cells, sizes = designCells(x)
    three = sizes[cells] >= 3
    if three.any():
        v = trimmedCellVariance(cnts[:, three], cells[three])
    else:
        v = trimmedVariance(cnts)
    m = cnts.mean(axis=1)
    return np.maximum((v - m) / m ** 2, 0.04)


def calculateCooksDistance(counts, mu, H, sizeFactors, x):
    """R/core.R:2333-2340."""
    p = x.shape[1]
    disp = robustMethodOfMomentsDisp(counts, sizeFactors, x)
    V = mu + disp[:, None] * mu ** 2
    return (np.asarray(counts, dtype=np.float64) - mu) ** 2 / V / p * H / (1 - H) ** 2


def recordMaxCooks(x, cooks):
    """R/core.R:2349-2359: max Cook's distance over the samples that have >= 3 replicates in their cell."""
    cells, sizes = designCells(x)
    samplesForCooks = sizes[cells] >= 3
    m, p = x.shape
    if m > p and samplesForCooks.any():
        return cooks[:, samplesForCooks].max(axis=1)
    return np.full(cooks.shape[0], np.nan)


def nOrMoreInCell(x, n):
    """R/core.R:2366-2371: per sample, does its design cell (distinct design row) hold at least n samples?"""
    cells, sizes = designCells(x)
    return sizes[cells] >= n


def replaceOutliers(counts, cooks, sizeFactors, x, trim=0.2, cooksCutoff=None, minReplicates=7):
    """R/core.R:2069-2115 (size factors, no normalisation-factor matrix).  Returns (counts with the outlying entries of
    replaceable samples set to as.integer(trimmed mean of the normalised counts * size factor), per-gene `replace`
    flag, per-sample `replaceable` flag).  Like the reference, `replace` is raised by an outlier in ANY sample."""
    from scipy import stats as _st
    counts = np.asarray(counts)
    n, m = counts.shape
    p = x.shape[1]
    if minReplicates < 3:
        raise ValueError("at least 3 replicates are necessary in order to indentify a sample as a count outlier")
    if m <= p:
        return counts.copy(), np.zeros(n, bool), np.zeros(m, bool)
    if cooksCutoff is None:
        cooksCutoff = _st.f.ppf(0.99, p, m - p)
    with np.errstate(invalid="ignore"):
        over = cooks > cooksCutoff                      # NA > cutoff is NA in R and drops out of which()/any()
    replace = over.any(axis=1)
    trimBaseMean = _r_trimmed_mean(counts / sizeFactors[None, :], trim)
    replacementCounts = np.trunc(trimBaseMean[:, None] * sizeFactors[None, :]).astype(counts.dtype)   # as.integer
    newCounts = counts.copy()
    newCounts[over] = replacementCounts[over]
    whichSamples = nOrMoreInCell(x, minReplicates)
    out = counts.copy()
    out[:, whichSamples] = newCounts[:, whichSamples]
    return out, replace, whichSamples


def getContrast(counts, nf, x, dispersion, betaMatrix, contrast, betaPriorVar=None, engine=None, minmu=0.5):
    """R/results.R:760-827: a numeric contrast c of the fitted coefficients, batched over genes through fitBeta's
    maxit = 0 mode (no IRLS iteration: covariance at the given betas, c'beta and sqrt(c' Sigma c)).  `counts` are the
    counts the model was fitted to (replaceCounts after an outlier refit), `betaMatrix` is on the log2 scale; rows are
    the non-all-zero genes.  Returns log2FoldChange, lfcSE, stat, pvalue (useT = FALSE)."""
    engine = engine or _default_engine
    counts = np.asarray(counts)
    n, m = counts.shape
    p = x.shape[1]
    contrast = np.asarray(contrast, dtype=np.float64)
    if contrast.shape != (p,):
        raise ValueError("numeric contrast vector should have one element for every element of 'resultsNames(object)'")
    if betaPriorVar is None:
        betaPriorVar = np.full(p, 1e6)                                    # betaPrior = FALSE (R/core.R:1421)
    lam = 1.0 / (LN2 ** 2 * np.asarray(betaPriorVar, dtype=np.float64))
    r = engine.fitBeta(ySEXP=counts, xSEXP=x, nfSEXP=np.broadcast_to(nf, counts.shape), alpha_hatSEXP=dispersion,
                       contrastSEXP=contrast, beta_matSEXP=LN2 * np.asarray(betaMatrix, dtype=np.float64),
                       lambdaSEXP=lam, weightsSEXP=None, useWeightsSEXP=False, tolSEXP=1e-8, maxitSEXP=0,
                       useQRSEXP=False, minmuSEXP=minmu)
    est = np.asarray(r["contrast_num"]).reshape(n) / LN2
    se = np.asarray(r["contrast_denom"]).reshape(n) / LN2
    with np.errstate(divide="ignore", invalid="ignore"):
        stat = est / se
    return {"log2FoldChange": est, "lfcSE": se, "stat": stat, "pvalue": 2.0 * _sp.ndtr(-np.abs(stat))}


def _fit_intercept_only(counts, nf, alpha_hat):
    """R/fitNbinomGLMs.R:99-137: reduced model ~1 with the default wide prior needs no IRLS (no native call)."""
    norm = counts / nf
    betaMatrix = np.log2(norm.mean(axis=1))[:, None]
    mu = nf * (2.0 ** betaMatrix)
    logLike = nbinomLogLike(counts, mu, alpha_hat)
    w = 1.0 / (1.0 / mu + alpha_hat[:, None])
    xtwx = w.sum(axis=1)
    return {"logLike": logLike, "betaConv": np.ones(len(counts), bool), "betaMatrix": betaMatrix,
            "betaSE": (np.sqrt(1.0 / xtwx) / LN2)[:, None], "mu": mu, "betaIter": np.ones(len(counts)),
            "hat_diagonals": w / xtwx[:, None]}


def nbinomLRT(counts, nf, full, reduced, dispersion, engine=None, betaTol=1e-8, maxit=100, useQR=True, minmu=0.5,
              useOptim=True):
    """R/core.R:1787-2012 with user-supplied model matrices, betaPrior=FALSE; Cook's distances are not restated.
    `full` / `reduced` are m x p model matrices (reduced nested in full)."""
    from scipy import stats as _st
    counts = np.asarray(counts)
    if reduced.shape[1] >= full.shape[1]:
        raise ValueError("less than one degree of freedom, perhaps full and reduced models are not in the correct order")
    fullModel = fitNbinomGLMs(counts, nf, full, dispersion, engine=engine, betaTol=betaTol, maxit=maxit, useQR=useQR,
                              minmu=minmu, useOptim=useOptim)
    if reduced.shape[1] == 1 and np.all(reduced == 1):
        reducedModel = _fit_intercept_only(counts, nf, dispersion)
    else:
        reducedModel = fitNbinomGLMs(counts, nf, reduced, dispersion, engine=engine, betaTol=betaTol, maxit=maxit,
                                     useQR=useQR, minmu=minmu, useOptim=useOptim)
    df = full.shape[1] - reduced.shape[1]
    LRTStatistic = 2.0 * (fullModel["logLike"] - reducedModel["logLike"])
    LRTPvalue = _st.chi2.sf(LRTStatistic, df)
    return {"LRTStatistic": LRTStatistic, "LRTPvalue": LRTPvalue, "deviance": -2.0 * fullModel["logLike"],
            "betaMatrix": fullModel["betaMatrix"], "betaSE": fullModel["betaSE"], "fullBetaConv": fullModel["betaConv"],
            "reducedBetaConv": reducedModel["betaConv"], "betaIter": fullModel["betaIter"], "mu": fullModel["mu"],
            "hat_diagonals": fullModel["hat_diagonals"], "df": df}


# ---------------------------------------------------------------- beta prior (config 4: betaPrior=TRUE, expanded matrix)

def factorDesign(factors, expanded=False):
    """Model matrix of an additive design of factors, ~f1 + f2 + ... (no interactions).
    factors: list of integer level-code vectors (0 = reference level).  standard: intercept + one indicator per
    non-reference level (model.matrix treatment contrasts); expanded: intercept + one indicator for EVERY level
    (makeExpandedModelMatrix, R/expanded.R:1-18).  Returns (matrix, names, factor index per column (-1 = intercept),
    level per column)."""
    factors = [np.asarray(f) for f in factors]
    m = len(factors[0])
    cols, names, fidx, lvl = [np.ones(m)], ["Intercept"], [-1], [-1]
    for k, f in enumerate(factors):
        for l in range(0 if expanded else 1, int(f.max()) + 1):
            cols.append((f == l).astype(np.float64))
            names.append(f"f{k}{l}")
            fidx.append(k)
            lvl.append(l)
    return np.stack(cols, axis=1), names, np.asarray(fidx), np.asarray(lvl)


def wtd_quantile(x, weights, prob):
    """Hmisc.wtd.quantile(x, weights, prob, type='quantile', normwt=TRUE)  (R/core.R:2762-2803)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    keep = ~(np.isnan(w) | (w == 0))
    x, w = x[keep], w[keep]
    w = w * len(x) / w.sum()                       # normwt
    order = np.argsort(x, kind="stable")
    xs, ws = x[order], w[order]
    ux, inv = np.unique(xs, return_inverse=True)   # wtd.table: sum of weights per distinct x
    wts = np.bincount(inv, weights=ws)
    n = wts.sum()
    o = 1 + (n - 1) * prob
    low = max(np.floor(o), 1.0)
    high = min(low + 1, n)
    frac = o % 1
    cum = np.cumsum(wts)

    def at(v):      # approx(cumsum(wts), x, xout=v, method='constant', f=1, rule=2): smallest x with cum >= v
        i = np.searchsorted(cum, v - 1e-12 * max(1.0, v), side="left")
        return ux[min(i, len(ux) - 1)]

    return (1 - frac) * at(low) + frac * at(high)


def matchWeightedUpperQuantileForVariance(x, weights, upperQuantile=0.05):
    """R/core.R:2416-2419."""
    from scipy import stats as _st
    sdEst = wtd_quantile(np.abs(x), weights, 1 - upperQuantile) / _st.norm.ppf(1 - upperQuantile / 2)
    return sdEst ** 2


def estimateBetaPriorVar(betaMatrix, names, fidx, baseMean, dispFit, expandedType=True, upperQuantile=0.05):
    """R/core.R:1601-1689 (betaPriorMethod='weighted').  betaMatrix: MLE log2 fold changes on the STANDARD matrix
    (columns described by names / fidx from factorDesign(expanded=False)).  Returns the prior variances for the
    standard columns and, when expandedType, for the expanded matrix columns (averagePriorsOverLevels,
    R/expanded.R:20-73) given `levels_per_factor` inferred from fidx."""
    weights = 1.0 / (1.0 / baseMean + dispFit)
    cols = {nm: betaMatrix[:, k] for k, nm in enumerate(names)}
    colsets = [(nm, betaMatrix[:, k], fidx[k]) for k, nm in enumerate(names)]
    if expandedType:                                # addAllContrasts (R/expanded.R:76-100)
        for f in sorted(set(int(v) for v in fidx if v >= 0)):
            M = betaMatrix[:, fidx == f]
            nlev = M.shape[1]
            for j in range(nlev - 1):
                for i in range(j + 1, nlev):
                    colsets.append((f"f{f}Cntrst", M[:, i] - M[:, j], f))

    def one(x):
        fin = np.abs(x) < 10
        if fin.sum() == 0:
            return 1e6
        return matchWeightedUpperQuantileForVariance(x[fin], weights[fin], upperQuantile)

    pv = [(nm, one(x) if nm != "Intercept" else 1e6, f) for nm, x, f in colsets]
    std = np.array([v for (nm, v, f) in pv[:len(names)]])
    if not expandedType:
        return std, None
    mean_by_factor = {}
    for f in sorted(set(int(v) for v in fidx if v >= 0)):
        mean_by_factor[f] = float(np.mean([v for (nm, v, ff) in pv if ff == f]))
    return std, mean_by_factor


def fitGLMsWithPrior(counts, nf, factors, dispersion, baseMean, dispFit, engine=None, betaTol=1e-8, maxit=100,
                     useQR=True, minmu=0.5):
    """R/fitNbinomGLMs.R:242-337 for an additive factor design with the expanded model matrix (the default when
    betaPrior=TRUE): MLE fit on the standard matrix -> estimateBetaPriorVar -> MAP fit on the expanded matrix with
    lambda = 1 / betaPriorVar."""
    x_std, names, fidx, _ = factorDesign(factors, expanded=False)
    mle = fitNbinomGLMs(counts, nf, x_std, dispersion, engine=engine, betaTol=betaTol, maxit=maxit, useQR=useQR,
                        minmu=minmu)
    _, mean_by_factor = estimateBetaPriorVar(mle["betaMatrix"], names, fidx, baseMean, dispFit, expandedType=True)
    x_exp, enames, efidx, _ = factorDesign(factors, expanded=True)
    betaPriorVar = np.array([1e6 if f < 0 else mean_by_factor[int(f)] for f in efidx])
    if np.any(betaPriorVar == 0):
        raise ValueError("beta prior variances are equal to zero for some variables")
    fit = fitNbinomGLMs(counts, nf, x_exp, dispersion, lambda_=1.0 / betaPriorVar, engine=engine, betaTol=betaTol,
                        maxit=maxit, useQR=useQR, minmu=minmu)
    return {"fit": fit, "H": mle["hat_diagonals"], "mu": mle["mu"], "betaPriorVar": betaPriorVar,
            "modelMatrix": x_exp, "mleBetaMatrix": mle["betaMatrix"], "names": enames, "mle": mle}


def DESeq(counts, x, sizeFactors=None, engine=None, minReplicatesForReplace=np.inf, useOptim=True):
    """R/core.R:280-432 with test='Wald', fitType='parametric', betaPrior=FALSE.
    minReplicatesForReplace: the reference's default is 7 (outlier replacement + refit whenever a design cell has >= 7
    samples, R/core.R:419-426, refitWithoutOutliers :2484-2565); the default here is Inf = off, which is what the
    engine parity tests and bench.py's `full_pipeline` exercise.  With a finite value the result also carries
    `replace` (genes refitted), `replaceable` (samples) and `replaceCounts`.
    Returns per-gene vectors over ALL rows; all-zero rows carry NaN (buildDataFrameWithNARows, R/core.R:2232)."""
    engine = engine or _default_engine
    counts = np.asarray(counts)
    N, m = counts.shape
    p = x.shape[1]
    if sizeFactors is None:
        sizeFactors = estimateSizeFactorsForMatrix(counts)
    mvAll = getBaseMeansAndVariances(counts, sizeFactors)
    nz = ~mvAll["allZero"]
    cnz = counts[nz]
    mv = {k: v[nz] for k, v in mvAll.items()}
    ge = estimateDispersionsGeneEst(cnz, sizeFactors, x, engine=engine, mv=mv)
    tf = estimateDispersionsFit(ge["dispGeneEst"], ge["baseMean"])
    dispPriorVar = estimateDispersionsPriorVar(tf["varLogDispEsts"], m, p, ge["dispGeneEst"], tf["dispFit"])
    mp = estimateDispersionsMAP(cnz, x, ge["mu"], ge["dispGeneEst"], tf["dispFit"], dispPriorVar,
                                tf["varLogDispEsts"], engine=engine)
    nf = np.broadcast_to(sizeFactors[None, :], cnz.shape)
    wt = nbinomWaldTest(cnz, nf, x, mp["dispersion"], engine=engine, useOptim=useOptim)
    cooks = calculateCooksDistance(cnz, wt["mu"], wt["hat_diagonals"], sizeFactors, x)      # R/core.R:1457-1460
    per_gene = {"baseMean": mv["baseMean"], "dispGeneEst": ge["dispGeneEst"], "dispFit": tf["dispFit"],
                "dispMAP": mp["dispMAP"], "dispersion": mp["dispersion"], "dispOutlier": mp["dispOutlier"].astype(float),
                "betaMatrix": wt["betaMatrix"], "betaSE": wt["betaSE"], "WaldStatistic": wt["WaldStatistic"],
                "WaldPvalue": wt["WaldPvalue"], "betaConv": wt["betaConv"].astype(float), "betaIter": wt["betaIter"],
                "deviance": wt["deviance"], "dispGeneIter": ge["dispGeneIter"].astype(float),
                "dispIter": mp["dispIter"].astype(float), "maxCooks": recordMaxCooks(x, cooks)}
    extra = {}
    if np.isfinite(minReplicatesForReplace) and nOrMoreInCell(x, minReplicatesForReplace).any():
        # ---- refitWithoutOutliers (R/core.R:2484-2565) on the rows that had a count replaced
        newc, replace, replaceable = replaceOutliers(cnz, cooks, sizeFactors, x, minReplicates=minReplicatesForReplace)
        nrefit = int(replace.sum())
        newAllZero = replace & (newc.sum(axis=1) == 0)
        if nrefit > 0:
            mvNew = getBaseMeansAndVariances(newc, sizeFactors)
            per_gene["baseMean"] = mvNew["baseMean"]
        if nrefit > 0 and nrefit > int(newAllZero.sum()):
            rr = replace & ~newAllZero
            sub = newc[rr]
            ge2 = estimateDispersionsGeneEst(sub, sizeFactors, x, engine=engine)
            dispFit2 = tf["coefs"][0] + tf["coefs"][1] / ge2["baseMean"]       # dispersionFunction(object)(baseMean)
            mp2 = estimateDispersionsMAP(sub, x, ge2["mu"], ge2["dispGeneEst"], dispFit2, dispPriorVar,
                                         tf["varLogDispEsts"], engine=engine)
            wt2 = nbinomWaldTest(sub, np.broadcast_to(sizeFactors[None, :], sub.shape), x, mp2["dispersion"],
                                 engine=engine, useOptim=useOptim)
            upd = {"dispGeneEst": ge2["dispGeneEst"], "dispFit": dispFit2, "dispMAP": mp2["dispMAP"],
                   "dispersion": mp2["dispersion"], "dispOutlier": mp2["dispOutlier"].astype(float),
                   "betaMatrix": wt2["betaMatrix"], "betaSE": wt2["betaSE"], "WaldStatistic": wt2["WaldStatistic"],
                   "WaldPvalue": wt2["WaldPvalue"], "betaConv": wt2["betaConv"].astype(float),
                   "betaIter": wt2["betaIter"], "deviance": wt2["deviance"],
                   "dispGeneIter": ge2["dispGeneIter"].astype(float), "dispIter": mp2["dispIter"].astype(float)}
            for k, v in upd.items():
                per_gene[k] = per_gene[k].copy()
                per_gene[k][rr] = v
            for k in ("betaMatrix", "betaSE", "WaldStatistic", "WaldPvalue", "betaConv", "betaIter", "deviance"):
                per_gene[k][newAllZero] = np.nan                                 # "results" columns (:2534)
            if replaceable.all():
                per_gene["maxCooks"] = np.full(len(cnz), np.nan)
            else:
                rc = cooks.copy()
                rc[:, replaceable] = 0.0
                per_gene["maxCooks"] = recordMaxCooks(x, rc)
        full_counts = counts.copy()
        full_counts[nz] = newc
        rep_all = np.zeros(N, bool)
        rep_all[nz] = replace
        extra = {"replace": rep_all, "replaceable": replaceable, "replaceCounts": full_counts, "n_replaced": nrefit}

    def full(v):
        out = np.full((N,) + v.shape[1:], np.nan)
        out[nz] = v
        return out

    res = {k: full(np.asarray(v, dtype=np.float64)) for k, v in per_gene.items()}
    res.update({"sizeFactors": sizeFactors, "allZero": mvAll["allZero"], "dispPriorVar": dispPriorVar,
                "trendCoefs": tf["coefs"], "varLogDispEsts": tf["varLogDispEsts"], "n_refit_geneest": ge["n_refit"],
                "n_refit_map": mp["n_refit"]})
    res["baseMean"] = np.where(nz, res["baseMean"], 0.0)
    res.update(extra)
    return res
