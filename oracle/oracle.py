"""ctypes binding of the CPU oracle (oracle/nbglm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under deseq2_b200/ may import this module.

Function names, argument names and returned dict keys are those of the reference's native
routines (src/DESeq2.cpp:164,283,469; list members :268-276, :458-464, :512).
Matrices are numpy arrays in R layout (genes x samples); they are converted to Fortran
(column-major) order before the call, exactly what R hands to .Call.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnbglm_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nbglm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_dnbinom_mu_log.restype = C.c_double
        _lib.oracle_dnbinom_mu_log.argtypes = [C.c_double] * 3
        for f in ("oracle_digamma", "oracle_trigamma", "oracle_lgamma"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double]
        _lib.oracle_log_posterior_row.restype = C.c_double
    return _lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def fitDisp(ySEXP, xSEXP, mu_hatSEXP, log_alphaSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
            min_log_alphaSEXP, kappa_0SEXP, tolSEXP, maxitSEXP, usePriorSEXP, weightsSEXP, useWeightsSEXP,
            weightThresholdSEXP, useCRSEXP, with_margin=False):
    y = _f(ySEXP); x = _f(xSEXP); mu = _f(mu_hatSEXP)
    n, m = y.shape
    p = x.shape[1]
    la = np.ascontiguousarray(log_alphaSEXP, dtype=np.float64)
    pm = np.ascontiguousarray(log_alpha_prior_meanSEXP, dtype=np.float64)
    w = _f(weightsSEXP) if weightsSEXP is not None else None
    out = {k: np.zeros(n) for k in ("log_alpha", "last_change", "initial_lp", "initial_dlp", "last_lp",
                                    "last_dlp", "last_d2lp")}
    out["iter"] = np.zeros(n, dtype=np.int32)
    out["iter_accept"] = np.zeros(n, dtype=np.int32)
    margin = np.zeros(n) if with_margin else None
    lib().oracle_fit_disp(
        _p(y), _p(x), _p(mu), _p(la), _p(pm), C.c_double(log_alpha_prior_sigmasqSEXP), C.c_double(min_log_alphaSEXP),
        C.c_double(kappa_0SEXP), C.c_double(tolSEXP), C.c_int(int(maxitSEXP)), C.c_int(bool(usePriorSEXP)), _p(w),
        C.c_int(bool(useWeightsSEXP)), C.c_double(weightThresholdSEXP), C.c_int(bool(useCRSEXP)),
        C.c_int(n), C.c_int(m), C.c_int(p),
        _p(out["log_alpha"]), out["iter"].ctypes.data_as(_ip), out["iter_accept"].ctypes.data_as(_ip),
        _p(out["last_change"]), _p(out["initial_lp"]), _p(out["initial_dlp"]), _p(out["last_lp"]),
        _p(out["last_dlp"]), _p(out["last_d2lp"]), _p(margin))
    if with_margin:
        out["margin"] = margin
    return out


def fitDispGrid(ySEXP, xSEXP, mu_hatSEXP, disp_gridSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
                usePriorSEXP, weightsSEXP, useWeightsSEXP, weightThresholdSEXP, useCRSEXP):
    y = _f(ySEXP); x = _f(xSEXP); mu = _f(mu_hatSEXP)
    n, m = y.shape
    p = x.shape[1]
    grid = np.ascontiguousarray(disp_gridSEXP, dtype=np.float64)
    pm = np.ascontiguousarray(log_alpha_prior_meanSEXP, dtype=np.float64)
    w = _f(weightsSEXP) if weightsSEXP is not None else None
    la = np.zeros(n)
    lib().oracle_fit_disp_grid(
        _p(y), _p(x), _p(mu), _p(grid), C.c_int(len(grid)), _p(pm), C.c_double(log_alpha_prior_sigmasqSEXP),
        C.c_int(bool(usePriorSEXP)), _p(w), C.c_int(bool(useWeightsSEXP)), C.c_double(weightThresholdSEXP),
        C.c_int(bool(useCRSEXP)), C.c_int(n), C.c_int(m), C.c_int(p), _p(la))
    return {"log_alpha": la}


def fitBeta(ySEXP, xSEXP, nfSEXP, alpha_hatSEXP, contrastSEXP, beta_matSEXP, lambdaSEXP, weightsSEXP,
            useWeightsSEXP, tolSEXP, maxitSEXP, useQRSEXP, minmuSEXP):
    y = _f(ySEXP); x = _f(xSEXP); nf = _f(nfSEXP)
    n, m = y.shape
    p = x.shape[1]
    alpha = np.ascontiguousarray(alpha_hatSEXP, dtype=np.float64)
    contrast = np.ascontiguousarray(contrastSEXP, dtype=np.float64)
    beta = np.array(beta_matSEXP, dtype=np.float64, order="F", copy=True).reshape(n, p, order="F")
    lam = np.ascontiguousarray(lambdaSEXP, dtype=np.float64)
    w = _f(weightsSEXP) if weightsSEXP is not None else None
    var = np.zeros((n, p), order="F")
    it = np.zeros(n)
    H = np.zeros((n, m), order="F")
    cn = np.zeros((n, 1)); cd = np.zeros((n, 1)); dev = np.zeros(n)
    lib().oracle_fit_beta(
        _p(y), _p(x), _p(nf), _p(alpha), _p(contrast), _p(beta), _p(lam), _p(w), C.c_int(bool(useWeightsSEXP)),
        C.c_double(tolSEXP), C.c_int(int(maxitSEXP)), C.c_int(bool(useQRSEXP)), C.c_double(minmuSEXP),
        C.c_int(n), C.c_int(m), C.c_int(p), _p(var), _p(it), _p(H), _p(cn), _p(cd), _p(dev))
    return {"beta_mat": beta, "beta_var_mat": var, "iter": it, "hat_diagonals": H, "contrast_num": cn,
            "contrast_denom": cd, "deviance": dev}


def log_posterior_row(yrow, murow, x, log_alpha, prior_mean=0.0, prior_sigmasq=1.0, usePrior=False, wrow=None,
                      useWeights=False, weightThreshold=1e-2, useCR=True, deriv=0):
    yrow = np.ascontiguousarray(yrow, dtype=np.float64); murow = np.ascontiguousarray(murow, dtype=np.float64)
    x = _f(x)
    w = np.ascontiguousarray(wrow, dtype=np.float64) if wrow is not None else None
    return lib().oracle_log_posterior_row(
        _p(yrow), _p(murow), _p(w), _p(x), C.c_int(x.shape[0]), C.c_int(x.shape[1]), C.c_double(log_alpha),
        C.c_double(prior_mean), C.c_double(prior_sigmasq), C.c_int(bool(usePrior)), C.c_int(bool(useWeights)),
        C.c_double(weightThreshold), C.c_int(bool(useCR)), C.c_int(deriv))


def dnbinom_mu_log(x, size, mu):
    return lib().oracle_dnbinom_mu_log(float(x), float(size), float(mu))
