"""ctypes binding of oracle/_ref/libdeseq2_ref.so = the REFERENCE's own src/DESeq2.cpp, compiled unchanged.

TEST INFRASTRUCTURE ONLY (like oracle/oracle.py): imported by tests/ (oracle == reference pin), __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  Nothing under deseq2_b200/ may import this module.

The library is built in the build container by `make -C oracle ref` (g++ reads /root/reference/src/DESeq2.cpp in place
against the stand-in Rcpp / Armadillo / Rmath headers of oracle/ref_standin/; the reference source is never copied)
and travels to the GPU box as a built file (oracle/_ref/ is git-ignored, not gpurun-ignored).  Function names, argument
names and returned dict keys are the reference's (src/DESeq2.cpp:164,283,469; :268-276, :458-464, :512).
`nthreads` > 1 = contiguous gene chunks, one independent reference call per chunk (R/parallel.R:9-10).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libdeseq2_ref.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def available() -> bool:
    return os.path.exists(_SO) or os.path.exists("/root/reference/src/DESeq2.cpp")


def build() -> str:
    if os.path.exists("/root/reference/src/DESeq2.cpp"):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libdeseq2_ref.so is absent and /root/reference is not here to build it")
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_source.restype = C.c_char_p
    return _lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference raised: " + lib().ref_last_error().decode())


def _is_int(y):
    return int(np.issubdtype(np.asarray(y).dtype, np.integer))


def fitDisp(ySEXP, xSEXP, mu_hatSEXP, log_alphaSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
            min_log_alphaSEXP, kappa_0SEXP, tolSEXP, maxitSEXP, usePriorSEXP, weightsSEXP, useWeightsSEXP,
            weightThresholdSEXP, useCRSEXP, nthreads=1):
    yi = _is_int(ySEXP)
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
    _check(lib().ref_fit_disp(
        _p(y), C.c_int(yi), _p(x), _p(mu), _p(la), _p(pm), C.c_double(log_alpha_prior_sigmasqSEXP),
        C.c_double(min_log_alphaSEXP), C.c_double(kappa_0SEXP), C.c_double(tolSEXP), C.c_int(int(maxitSEXP)),
        C.c_int(bool(usePriorSEXP)), _p(w), C.c_int(bool(useWeightsSEXP)), C.c_double(weightThresholdSEXP),
        C.c_int(bool(useCRSEXP)), C.c_int(n), C.c_int(m), C.c_int(p), C.c_int(nthreads),
        _p(out["log_alpha"]), out["iter"].ctypes.data_as(_ip), out["iter_accept"].ctypes.data_as(_ip),
        _p(out["last_change"]), _p(out["initial_lp"]), _p(out["initial_dlp"]), _p(out["last_lp"]),
        _p(out["last_dlp"]), _p(out["last_d2lp"])))
    return out


def fitDispGrid(ySEXP, xSEXP, mu_hatSEXP, disp_gridSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
                usePriorSEXP, weightsSEXP, useWeightsSEXP, weightThresholdSEXP, useCRSEXP, nthreads=1):
    yi = _is_int(ySEXP)
    y = _f(ySEXP); x = _f(xSEXP); mu = _f(mu_hatSEXP)
    n, m = y.shape
    p = x.shape[1]
    grid = np.ascontiguousarray(disp_gridSEXP, dtype=np.float64)
    pm = np.ascontiguousarray(log_alpha_prior_meanSEXP, dtype=np.float64)
    w = _f(weightsSEXP) if weightsSEXP is not None else None
    la = np.zeros(n)
    _check(lib().ref_fit_disp_grid(
        _p(y), C.c_int(yi), _p(x), _p(mu), _p(grid), C.c_int(len(grid)), _p(pm),
        C.c_double(log_alpha_prior_sigmasqSEXP), C.c_int(bool(usePriorSEXP)), _p(w), C.c_int(bool(useWeightsSEXP)),
        C.c_double(weightThresholdSEXP), C.c_int(bool(useCRSEXP)), C.c_int(n), C.c_int(m), C.c_int(p),
        C.c_int(nthreads), _p(la)))
    return {"log_alpha": la}


def fitBeta(ySEXP, xSEXP, nfSEXP, alpha_hatSEXP, contrastSEXP, beta_matSEXP, lambdaSEXP, weightsSEXP,
            useWeightsSEXP, tolSEXP, maxitSEXP, useQRSEXP, minmuSEXP, nthreads=1):
    yi = _is_int(ySEXP)
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
    _check(lib().ref_fit_beta(
        _p(y), C.c_int(yi), _p(x), _p(nf), _p(alpha), _p(contrast), _p(beta), _p(lam), _p(w),
        C.c_int(bool(useWeightsSEXP)), C.c_double(tolSEXP), C.c_int(int(maxitSEXP)), C.c_int(bool(useQRSEXP)),
        C.c_double(minmuSEXP), C.c_int(n), C.c_int(m), C.c_int(p), C.c_int(nthreads), _p(var), _p(it), _p(H),
        _p(cn), _p(cd), _p(dev)))
    return {"beta_mat": beta, "beta_var_mat": var, "iter": it, "hat_diagonals": H, "contrast_num": cn,
            "contrast_denom": cd, "deviance": dev}
