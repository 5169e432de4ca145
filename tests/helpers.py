"""Shared helpers for the parity tests: realistic fitDisp / fitBeta inputs built with the host glue."""
import os

import numpy as np

from deseq2_b200 import pipeline, synth

# `-m gpu` tests run on "cuda"; when B200NB_LIB points at the emulated engine (tests/simt_emu/_build/libb200nb_emu*.so,
# see tests/test_emulated_kernels.py) the same tests run with CPU tensors: in the emulator device pointers are host
# pointers (tests/conftest.py then also replaces the CUDA stream getters of the torch wrappers by a null stream).
EMULATED = "libb200nb_emu" in os.path.basename(os.environ.get("B200NB_LIB", ""))
DEV = "cpu" if EMULATED else "cuda"

DISP_KEYS = ("log_alpha", "last_change", "initial_lp", "initial_dlp", "last_lp", "last_dlp", "last_d2lp")


def make_case(n, m, x=None, seed=1, drop_zero=True, **kw):
    d = synth.make_example_counts(n, m, x=x, seed=seed, **kw)
    counts = d["counts"]
    if drop_zero:
        counts = counts[counts.sum(axis=1) > 0]
    sf = d["sizeFactors"]
    x = d["x"]
    mv = pipeline.getBaseMeansAndVariances(counts, sf)
    norm = counts / sf[None, :]
    rough = pipeline.roughDispEstimate(norm, x)
    mom = pipeline.momentsDispEstimate(mv["baseMean"], mv["baseVar"], sf)
    alpha0 = np.minimum(np.maximum(1e-8, np.minimum(rough, mom)), max(10, m))
    nf = np.broadcast_to(sf[None, :], counts.shape).copy()
    if pipeline.modelMatrixGroups(x) == x.shape[1]:
        mu = np.maximum(pipeline.linearModelMu(norm, x) * nf, 0.5)
    else:
        mu = None
    Q, R = np.linalg.qr(x)
    beta0 = np.linalg.solve(R, Q.T @ np.log(norm + 0.1).T).T
    return dict(counts=counts, x=x, sf=sf, nf=nf, mu=mu, alpha0=alpha0, beta0=beta0, baseMean=mv["baseMean"],
                trueDisp=d["trueDisp"])


def disp_args(c, mu, log_alpha, prior_mean=None, sigmasq=1.0, usePrior=False, tol=1e-6, maxit=100, weights=None,
              useWeights=False, useCR=True, y=None):
    return dict(ySEXP=c["counts"] if y is None else y, xSEXP=c["x"], mu_hatSEXP=mu, log_alphaSEXP=log_alpha,
                log_alpha_prior_meanSEXP=log_alpha if prior_mean is None else prior_mean,
                log_alpha_prior_sigmasqSEXP=sigmasq, min_log_alphaSEXP=np.log(1e-8 / 10), kappa_0SEXP=1.0,
                tolSEXP=tol, maxitSEXP=maxit, usePriorSEXP=usePrior, weightsSEXP=weights, useWeightsSEXP=useWeights,
                weightThresholdSEXP=1e-2, useCRSEXP=useCR)


def beta_args(c, alpha, beta0=None, lam=None, useQR=True, maxit=100, tol=1e-8, weights=None, useWeights=False,
              contrast=None, nf=None, x=None, y=None):
    x = c["x"] if x is None else x
    p = x.shape[1]
    return dict(ySEXP=c["counts"] if y is None else y, xSEXP=x, nfSEXP=c["nf"] if nf is None else nf,
                alpha_hatSEXP=alpha, contrastSEXP=np.r_[1.0, np.zeros(p - 1)] if contrast is None else contrast,
                beta_matSEXP=c["beta0"] if beta0 is None else beta0,
                lambdaSEXP=np.full(p, 1e-6) / np.log(2) ** 2 if lam is None else lam, weightsSEXP=weights,
                useWeightsSEXP=useWeights, tolSEXP=tol, maxitSEXP=maxit, useQRSEXP=useQR, minmuSEXP=0.5)


def rel_err(a, b, floor=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
