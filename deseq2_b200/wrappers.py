"""Host-side mirror of the reference's R -> native shim for the hot path.

Same function names, argument names, defaults and error behaviour as
  R/RcppExports.R:4-14   fitDisp / fitBeta / fitDispGrid   (thin .Call stubs)
  R/wrappers.R:26,63,102 fitDispWrapper / fitDispGridWrapper / fitBetaWrapper (NA screening, default
                         contrast c(1,0,...), 20-point log-alpha grid, exp() of the grid result)
but the call lands in libb200nb.so (include/b200nb.h) instead of the Rcpp routines in src/DESeq2.cpp.
Matrices are numpy arrays shaped like the R objects (genes x samples); they are handed over in
column-major order, which is what R's .Call would pass.  Results come back as dicts keyed by the
reference's list member names.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_Y_INT32, _Y_F64 = 0, 1


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _ymat(y):
    y = np.asarray(y)
    if y.ndim != 2:
        raise ValueError("the count matrix must be two-dimensional (genes x samples)")
    if y.dtype == np.int32:
        return np.asfortranarray(y), _Y_INT32
    if np.issubdtype(y.dtype, np.integer):
        # R's INTSXP is 32 bit; a wider integer that does not fit is an error, never a silent wrap
        if y.size and (y.max() > np.iinfo(np.int32).max or y.min() < np.iinfo(np.int32).min):
            raise ValueError("integer counts exceed the int32 range of R's integer type; pass them as float64")
        return np.asfortranarray(y.astype(np.int32)), _Y_INT32
    return np.asfortranarray(y, dtype=np.float64), _Y_F64


def _vec_n(a, n, name, fname, broadcast=False):
    """A length-n double vector; the C ABI reads exactly n doubles, so a wrong length would be an out-of-bounds read.
    broadcast=True accepts a scalar and repeats it (R recycles a length-1 lambda)."""
    v = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if broadcast and v.size == 1 and n != 1:
        v = np.full(n, v[0])
    if v.size != n:
        raise ValueError(f"{fname}: {name} has length {v.size}, expected {n}")
    return v


def _fmat(a):
    return np.asfortranarray(a, dtype=np.float64)


class _PinnedBlock:
    """A block of the engine's page-locked result memory (b200nb_host_alloc); returned to the pool when the numpy
    arrays viewing it are gone.  The R shim does the same through allocVector3's custom allocator."""
    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            _lib.lib().b200nb_host_free(C.c_void_p(self.ptr))
        except Exception:   # interpreter shutdown
            pass


def _result_vectors(spec):
    """Fresh 1-D / small 2-D result arrays described by [(name, shape, dtype)], carved out of ONE block of the engine's
    page-locked pool when together they are >= 256 KB (each then receives its data by DMA), else ordinary arrays."""
    sizes = [int(np.prod(shape)) * np.dtype(dt).itemsize for _, shape, dt in spec]
    offs = np.cumsum([0] + [(s + 63) & ~63 for s in sizes])
    total = int(offs[-1])
    block = None
    if total >= (1 << 18):
        ptr = _lib.lib().b200nb_host_alloc(total)
        if ptr:
            block = np.asarray(_PinnedBlock(ptr, total))
    out = {}
    for (name, shape, dt), off, sz in zip(spec, offs, sizes):
        if block is None:
            out[name] = np.empty(shape, dtype=dt, order="F")
        else:
            out[name] = block[off:off + sz].view(dt).reshape(shape, order="F")
    return out


def _result_matrix(n, m):
    """Fresh n x m float64 column-major result matrix.  Large ones live in page-locked memory of the engine's pool: the
    device-to-host copy is then one DMA into the array itself (no staging through host threads, no first-touch page
    faults); small ones, or when page-locked memory cannot be had, are ordinary numpy arrays."""
    nbytes = int(n) * int(m) * 8
    if nbytes >= (1 << 20):
        ptr = _lib.lib().b200nb_host_alloc(nbytes)
        if ptr:
            return np.asarray(_PinnedBlock(ptr, nbytes)).view(np.float64).reshape((n, m), order="F")
    return np.empty((n, m), order="F")


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _weights(weightsSEXP, useWeightsSEXP, shape=None, fname=""):
    if not useWeightsSEXP or weightsSEXP is None:
        return None
    w = _fmat(weightsSEXP)
    if shape is not None and w.shape != shape:
        raise ValueError(f"{fname}: weights have shape {w.shape}, expected {shape}")
    return w


def fitDisp(ySEXP, xSEXP, mu_hatSEXP, log_alphaSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
            min_log_alphaSEXP, kappa_0SEXP, tolSEXP, maxitSEXP, usePriorSEXP, weightsSEXP, useWeightsSEXP,
            weightThresholdSEXP, useCRSEXP):
    """R/RcppExports.R:4 -> src/DESeq2.cpp:164.  Returns the nine-member list of :268-276 as a dict."""
    L = _lib.lib()
    _lib.require_device()
    y, yt = _ymat(ySEXP)
    x = _fmat(xSEXP)
    mu = _fmat(mu_hatSEXP)
    n, m = y.shape
    p = x.shape[1]
    if x.shape[0] != m or mu.shape != (n, m):
        raise ValueError("fitDisp: non-conformable arguments")
    la = _vec_n(log_alphaSEXP, n, "log_alpha", "fitDisp")
    pm = _vec_n(log_alpha_prior_meanSEXP, n, "log_alpha_prior_mean", "fitDisp")
    w = _weights(weightsSEXP, useWeightsSEXP, (n, m), "fitDisp")
    out = _result_vectors([(k, (n,), np.float64) for k in ("log_alpha", "last_change", "initial_lp", "initial_dlp",
                                                           "last_lp", "last_dlp", "last_d2lp")]
                          + [("iter", (n,), np.int32), ("iter_accept", (n,), np.int32)])
    rc = L.b200nb_fit_disp(_ptr(y), yt, _ptr(x), _ptr(mu), _ptr(la), _ptr(pm), float(log_alpha_prior_sigmasqSEXP),
                           float(min_log_alphaSEXP), float(kappa_0SEXP), float(tolSEXP), int(maxitSEXP),
                           int(bool(usePriorSEXP)), _ptr(w), int(w is not None), float(weightThresholdSEXP),
                           int(bool(useCRSEXP)), n, m, p,
                           _ptr(out["log_alpha"]), _ptr(out["iter"]), _ptr(out["iter_accept"]),
                           _ptr(out["last_change"]), _ptr(out["initial_lp"]), _ptr(out["initial_dlp"]),
                           _ptr(out["last_lp"]), _ptr(out["last_dlp"]), _ptr(out["last_d2lp"]))
    _lib.check(rc, "fitDisp")
    return out


def fitDispGrid(ySEXP, xSEXP, mu_hatSEXP, disp_gridSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
                usePriorSEXP, weightsSEXP, useWeightsSEXP, weightThresholdSEXP, useCRSEXP):
    """R/RcppExports.R:12 -> src/DESeq2.cpp:469."""
    L = _lib.lib()
    _lib.require_device()
    y, yt = _ymat(ySEXP)
    x = _fmat(xSEXP)
    mu = _fmat(mu_hatSEXP)
    n, m = y.shape
    p = x.shape[1]
    if x.shape[0] != m or mu.shape != (n, m):
        raise ValueError("fitDispGrid: non-conformable arguments")
    grid = _vec(disp_gridSEXP)
    pm = _vec_n(log_alpha_prior_meanSEXP, n, "log_alpha_prior_mean", "fitDispGrid")
    w = _weights(weightsSEXP, useWeightsSEXP, (n, m), "fitDispGrid")
    la = np.empty(n)
    rc = L.b200nb_fit_disp_grid(_ptr(y), yt, _ptr(x), _ptr(mu), _ptr(grid), len(grid), _ptr(pm),
                                float(log_alpha_prior_sigmasqSEXP), int(bool(usePriorSEXP)), _ptr(w),
                                int(w is not None), float(weightThresholdSEXP), int(bool(useCRSEXP)), n, m, p,
                                _ptr(la))
    _lib.check(rc, "fitDispGrid")
    return {"log_alpha": la}


def fitBeta(ySEXP, xSEXP, nfSEXP, alpha_hatSEXP, contrastSEXP, beta_matSEXP, lambdaSEXP, weightsSEXP,
            useWeightsSEXP, tolSEXP, maxitSEXP, useQRSEXP, minmuSEXP, return_mu=False):
    """R/RcppExports.R:8 -> src/DESeq2.cpp:283.  Returns the seven-member list of :458-464 as a dict
    (plus "mu" when return_mu=True: the fused fitted mean the reference recomputes at R/fitNbinomGLMs.R:180)."""
    L = _lib.lib()
    _lib.require_device()
    y, yt = _ymat(ySEXP)
    x = _fmat(xSEXP)
    nf = _fmat(nfSEXP)
    n, m = y.shape
    p = x.shape[1]
    if x.shape[0] != m or nf.shape != (n, m):
        raise ValueError("fitBeta: non-conformable arguments")
    alpha = _vec_n(alpha_hatSEXP, n, "alpha_hat", "fitBeta")
    contrast = _vec_n(contrastSEXP, p, "contrast", "fitBeta")
    beta0 = np.asarray(beta_matSEXP, dtype=np.float64)
    if beta0.size != n * p:
        raise ValueError(f"fitBeta: beta_mat has {beta0.size} elements, expected {n} x {p}")
    beta0 = _fmat(beta0.reshape(n, p))
    lam = _vec_n(lambdaSEXP, p, "lambda", "fitBeta", broadcast=True)
    w = _weights(weightsSEXP, useWeightsSEXP, (n, m), "fitBeta")
    sm = _result_vectors([("beta", (n, p), np.float64), ("var", (n, p), np.float64), ("it", (n,), np.float64),
                          ("cn", (n, 1), np.float64), ("cd", (n, 1), np.float64), ("dev", (n,), np.float64)])
    beta, var, it, cn, cd, dev = (sm[k] for k in ("beta", "var", "it", "cn", "cd", "dev"))
    H = _result_matrix(n, m)
    mu = _result_matrix(n, m) if return_mu else None
    rc = L.b200nb_fit_beta(_ptr(y), yt, _ptr(x), _ptr(nf), _ptr(alpha), _ptr(contrast), _ptr(beta0), _ptr(lam),
                           _ptr(w), int(w is not None), float(tolSEXP), int(maxitSEXP), int(bool(useQRSEXP)),
                           float(minmuSEXP), n, m, p, _ptr(beta), _ptr(var), _ptr(it), _ptr(H), _ptr(cn), _ptr(cd),
                           _ptr(dev), _ptr(mu))
    _lib.check(rc, "fitBeta")
    out = {"beta_mat": beta, "beta_var_mat": var, "iter": it, "hat_diagonals": H, "contrast_num": cn,
           "contrast_denom": cd, "deviance": dev}
    if return_mu:
        out["mu"] = mu
    return out


# ---------------------------------------------------------------- R/wrappers.R

def _na_check(fname, **kwargs):
    bad = [k for k, v in kwargs.items() if v is not None and np.any(np.isnan(np.asarray(v, dtype=np.float64)))]
    if bad:
        raise ValueError(f"in call to {fname}, the following arguments contain NA: {', '.join(bad)}")


def fitDispWrapper(ySEXP, xSEXP, mu_hatSEXP, log_alphaSEXP, log_alpha_prior_meanSEXP, log_alpha_prior_sigmasqSEXP,
                   min_log_alphaSEXP, kappa_0SEXP, tolSEXP, maxitSEXP, usePriorSEXP, weightsSEXP, useWeightsSEXP,
                   weightThresholdSEXP, useCRSEXP, engine=None):
    """R/wrappers.R:26-42."""
    args = dict(ySEXP=ySEXP, xSEXP=xSEXP, mu_hatSEXP=mu_hatSEXP, log_alphaSEXP=log_alphaSEXP,
                log_alpha_prior_meanSEXP=log_alpha_prior_meanSEXP,
                log_alpha_prior_sigmasqSEXP=log_alpha_prior_sigmasqSEXP, min_log_alphaSEXP=min_log_alphaSEXP,
                kappa_0SEXP=kappa_0SEXP, tolSEXP=tolSEXP, maxitSEXP=maxitSEXP, usePriorSEXP=usePriorSEXP,
                weightsSEXP=weightsSEXP, useWeightsSEXP=useWeightsSEXP, weightThresholdSEXP=weightThresholdSEXP,
                useCRSEXP=useCRSEXP)
    _na_check("fitDisp", **args)
    return (engine.fitDisp if engine is not None else fitDisp)(**args)


def fitDispGridWrapper(y, x, mu, logAlphaPriorMean, logAlphaPriorSigmaSq, usePrior, weightsSEXP, useWeightsSEXP,
                       weightThresholdSEXP, useCRSEXP, engine=None):
    """R/wrappers.R:63-83: builds the 20-point grid on [log 1e-8, log max(10, ncol(y))], returns exp(log_alpha)."""
    _na_check("fitDispGridWrapper", y=y, x=x, mu=mu, logAlphaPriorMean=logAlphaPriorMean,
              logAlphaPriorSigmaSq=logAlphaPriorSigmaSq, weightsSEXP=weightsSEXP)
    minLogAlpha = np.log(1e-8)
    maxLogAlpha = np.log(max(10, np.asarray(y).shape[1]))
    dispGrid = np.linspace(minLogAlpha, maxLogAlpha, 20)
    f = engine.fitDispGrid if engine is not None else fitDispGrid
    logAlpha = f(ySEXP=y, xSEXP=x, mu_hatSEXP=mu, disp_gridSEXP=dispGrid, log_alpha_prior_meanSEXP=logAlphaPriorMean,
                 log_alpha_prior_sigmasqSEXP=logAlphaPriorSigmaSq, usePriorSEXP=usePrior, weightsSEXP=weightsSEXP,
                 useWeightsSEXP=useWeightsSEXP, weightThresholdSEXP=weightThresholdSEXP,
                 useCRSEXP=useCRSEXP)["log_alpha"]
    return np.exp(logAlpha)


def fitBetaWrapper(ySEXP, xSEXP, nfSEXP, alpha_hatSEXP, beta_matSEXP, lambdaSEXP, weightsSEXP, useWeightsSEXP,
                   tolSEXP, maxitSEXP, useQRSEXP, minmuSEXP, contrastSEXP=None, engine=None, **kw):
    """R/wrappers.R:102-119: default contrast c(1, 0, ..., 0); NA screening."""
    if contrastSEXP is None:
        contrastSEXP = np.r_[1.0, np.zeros(np.asarray(xSEXP).shape[1] - 1)]
    args = dict(ySEXP=ySEXP, xSEXP=xSEXP, nfSEXP=nfSEXP, alpha_hatSEXP=alpha_hatSEXP, contrastSEXP=contrastSEXP,
                beta_matSEXP=beta_matSEXP, lambdaSEXP=lambdaSEXP, weightsSEXP=weightsSEXP,
                useWeightsSEXP=useWeightsSEXP, tolSEXP=tolSEXP, maxitSEXP=maxitSEXP, useQRSEXP=useQRSEXP,
                minmuSEXP=minmuSEXP)
    _na_check("fitBeta", **args)
    return (engine.fitBeta if engine is not None else fitBeta)(**args, **kw)
