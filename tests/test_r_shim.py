"""The R .Call shim (deseq2_b200/r_shim/deseq2_b200_shim.c; replaces /root/reference/src/RcppExports.cpp:16-94) is
compiled UNCHANGED against a mock of R's C API (tests/mock_r/) and driven the way R drives it: 15 / 13 / 11 SEXP
arguments with R's types (integer count matrix, numeric matrices with a dim attribute, length-1 scalars, logical
flags), a named list back.  R itself is not installed here, so this is the strongest check of the R-facing half of
the drop-in boundary available without R.

CPU test: the shim is linked against tests/mock_r/engine_stub_oracle.c (the C ABI implemented on the oracle), so
every difference from a direct oracle call is a marshalling bug.  GPU test: the same shim object code is linked
against the product library deseq2_b200/libb200nb.so and compared with the oracle at the parity tolerance.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import DISP_KEYS, beta_args, disp_args, make_case, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_r")
SHIM = os.path.join(ROOT, "deseq2_b200", "r_shim", "deseq2_b200_shim.c")
LGLSXP, INTSXP, REALSXP, VECSXP = 10, 13, 14, 19

DISP_NAMES = ["log_alpha", "iter", "iter_accept", "last_change", "initial_lp", "initial_dlp", "last_lp", "last_dlp",
              "last_d2lp"]                                                    # src/DESeq2.cpp:268-276
BETA_NAMES = ["beta_mat", "beta_var_mat", "iter", "hat_diagonals", "contrast_num", "contrast_denom", "deviance"]  # :458-464


def _build(tmp, name, extra_src, link):
    so = os.path.join(tmp, name)
    cmd = ["/usr/bin/gcc", "-O1", "-fPIC", "-shared", "-Wall", "-Werror=implicit-function-declaration",
           "-Werror=incompatible-pointer-types", "-Werror=int-conversion", "-I", MOCK, "-o", so, SHIM,
           os.path.join(MOCK, "mock_r.c")] + extra_src + link
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return so


class MockR:
    """Just enough of an R session: builds SEXPs from numpy, does .Call, reads the result list back."""

    def __init__(self, so):
        self.lib = L = C.CDLL(so)
        vp = C.c_void_p
        L.mock_real.restype = vp; L.mock_real.argtypes = [vp, C.c_long, C.c_int, C.c_int]
        L.mock_int.restype = vp; L.mock_int.argtypes = [vp, C.c_long, C.c_int, C.c_int, C.c_int]
        L.mock_string.restype = vp; L.mock_string.argtypes = [C.c_char_p]
        L.mock_nil.restype = vp
        L.mock_dot_call.restype = vp; L.mock_dot_call.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.mock_last_error.restype = C.c_char_p
        L.mock_typeof.argtypes = [vp]
        L.mock_length.restype = C.c_long; L.mock_length.argtypes = [vp]
        L.mock_data.restype = vp; L.mock_data.argtypes = [vp]
        L.mock_has_dim.argtypes = [vp]
        L.mock_dim.argtypes = [vp, C.c_int]
        L.mock_list_elt.restype = vp; L.mock_list_elt.argtypes = [vp, C.c_long]
        L.mock_name.restype = C.c_char_p; L.mock_name.argtypes = [vp, C.c_long]
        L.mock_run_init.argtypes = [vp]
        L.mock_registered_name.restype = C.c_char_p
        L.mock_registered_fun.restype = vp

    def sexp(self, v):
        L = self.lib
        if v is None:
            return L.mock_nil()
        if isinstance(v, str):
            return L.mock_string(v.encode())
        if isinstance(v, (bool, np.bool_)):
            a = np.array([int(v)], dtype=np.int32)
            return L.mock_int(a.ctypes.data, 1, -1, -1, 1)
        a = np.asarray(v)
        if a.dtype.kind in "iu" or a.dtype == np.bool_:
            lgl = int(a.dtype == np.bool_)
            a = np.asfortranarray(a, dtype=np.int32)
            mk = lambda n, r, c: L.mock_int(a.ctypes.data, n, r, c, lgl)
        else:
            a = np.asfortranarray(a, dtype=np.float64)
            mk = lambda n, r, c: L.mock_real(a.ctypes.data, n, r, c)
        if a.ndim == 2:
            return mk(a.size, a.shape[0], a.shape[1])
        a = a.reshape(-1)
        return mk(a.size, -1, -1)

    def value(self, s):
        """SEXP -> (numpy array copy in R's shape, typeof)."""
        L = self.lib
        t, n = L.mock_typeof(s), L.mock_length(s)
        ct = {INTSXP: C.c_int32, LGLSXP: C.c_int32, REALSXP: C.c_double}[t]
        a = np.ctypeslib.as_array(C.cast(L.mock_data(s), C.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
        if L.mock_has_dim(s):
            a = a.reshape((L.mock_dim(s, 0), L.mock_dim(s, 1)), order="F")
        return a, t

    def dot_call(self, symbol, *args):
        L = self.lib
        L.mock_reset()
        sx = [self.sexp(a) for a in args]
        before = [self.value(s)[0] if s != L.mock_nil() and L.mock_typeof(s) in (INTSXP, LGLSXP, REALSXP) else None
                  for s in sx]
        arr = (C.c_void_p * len(sx))(*sx)
        fn = C.cast(getattr(L, symbol), C.c_void_p)
        res = L.mock_dot_call(fn, len(sx), arr)
        assert L.mock_protect_underflow() == 0, "UNPROTECT of more than was PROTECTed"
        for s, b in zip(sx, before):                      # .Call arguments are R-owned: must come back untouched
            if b is not None:
                assert np.array_equal(self.value(s)[0], b, equal_nan=True), "the shim modified an input"
        if not res:
            raise RuntimeError(L.mock_last_error().decode())
        assert L.mock_protect_depth() == 0, "PROTECT / UNPROTECT imbalance on the normal return path"
        assert L.mock_typeof(res) == VECSXP
        out, types = {}, {}
        for i in range(L.mock_length(res)):
            name = L.mock_name(res, i)
            assert name is not None, "result list has no names attribute"
            out[name.decode()], types[name.decode()] = self.value(L.mock_list_elt(res, i))
        return out, types


def _ordered(d):
    return list(d.values())


@pytest.fixture(scope="module")
def r_oracle(oracle, tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("shim_oracle"))
    odir = os.path.join(ROOT, "oracle")
    so = _build(tmp, "DESeq2_oracle.so", [os.path.join(MOCK, "engine_stub_oracle.c")],
                ["-L" + odir, "-lnbglm_oracle", "-Wl,-rpath," + odir])
    return MockR(so)


@pytest.fixture(scope="module")
def r_engine(tmp_path_factory):
    import deseq2_b200
    tmp = str(tmp_path_factory.mktemp("shim_engine"))
    deseq2_b200.lib()                                    # fails loudly if the engine has not been built
    path = deseq2_b200.lib_path()                        # = -L<dir> -lb200nb of INTEGRATION.md section 2
    so = _build(tmp, "DESeq2.so", [], [path, "-Wl,-rpath," + os.path.dirname(path)])
    return MockR(so)


def _case():
    c = make_case(60, 8, seed=5)
    ones = np.ones(c["counts"].shape)
    la = np.log(c["alpha0"])
    return c, ones, la


def test_registration_table(r_oracle):
    """R_init_DESeq2 registers the three .Call routines with the reference's arities (src/RcppExports.cpp:84-94)
    and switches dynamic lookup off."""
    L = r_oracle.lib
    n = L.mock_run_init(C.cast(L.R_init_DESeq2, C.c_void_p))
    got = {L.mock_registered_name(i).decode(): L.mock_registered_nargs(i) for i in range(n)}
    assert got == {"_DESeq2_fitDisp": 15, "_DESeq2_fitBeta": 13, "_DESeq2_fitDispGrid": 11}
    for i in range(n):
        assert L.mock_registered_fun(i) == C.cast(getattr(L, L.mock_registered_name(i).decode()), C.c_void_p).value
    assert L.mock_dynamic_symbols() == 0


def test_fitDisp_marshalling(r_oracle, oracle):
    c, ones, la = _case()
    a = disp_args(c, c["mu"], la, weights=ones)
    before = r_oracle.lib.mock_interrupt_checks()
    got, types = r_oracle.dot_call("_DESeq2_fitDisp", *_ordered(a))
    assert r_oracle.lib.mock_interrupt_checks() == before + 1          # user interrupts are polled, as in the reference
    assert list(got) == DISP_NAMES
    assert [types[k] for k in DISP_NAMES] == [REALSXP, INTSXP, INTSXP] + [REALSXP] * 6     # src/DESeq2.cpp:181-190
    ref = oracle.fitDisp(**a)
    for k in DISP_NAMES:
        assert got[k].shape == (len(la),) and np.array_equal(got[k], ref[k], equal_nan=True), k
    # R passes whatever storage mode the caller has: a double count matrix, integer maxit, numeric flags, an
    # integer design matrix, and a weights matrix that is really used
    w = np.random.default_rng(3).uniform(0.2, 1.0, ones.shape)
    b = disp_args(c, c["mu"], la, prior_mean=la + 0.3, sigmasq=0.7, usePrior=True, weights=w, useWeights=True, maxit=25)
    alt = dict(b, ySEXP=c["counts"].astype(np.float64), xSEXP=c["x"].astype(np.int32), maxitSEXP=np.array([25.0]),
               usePriorSEXP=np.array([1.0]), useWeightsSEXP=np.array([1], dtype=np.int32))
    got2, _ = r_oracle.dot_call("_DESeq2_fitDisp", *_ordered(alt))
    ref2 = oracle.fitDisp(**b)
    for k in DISP_NAMES:
        assert np.array_equal(got2[k], ref2[k], equal_nan=True), k
    assert not np.array_equal(got2["log_alpha"], got["log_alpha"])


def test_fitDispGrid_marshalling(r_oracle, oracle):
    c, ones, la = _case()
    grid = np.linspace(np.log(1e-8), np.log(10), 20)
    a = dict(ySEXP=c["counts"], xSEXP=c["x"], mu_hatSEXP=c["mu"], disp_gridSEXP=grid, log_alpha_prior_meanSEXP=la,
             log_alpha_prior_sigmasqSEXP=1.0, usePriorSEXP=False, weightsSEXP=ones, useWeightsSEXP=False,
             weightThresholdSEXP=1e-2, useCRSEXP=True)
    got, types = r_oracle.dot_call("_DESeq2_fitDispGrid", *_ordered(a))
    assert list(got) == ["log_alpha"] and types["log_alpha"] == REALSXP                    # src/DESeq2.cpp:512
    assert np.array_equal(got["log_alpha"], oracle.fitDispGrid(**a)["log_alpha"])


def test_fitBeta_marshalling(r_oracle, oracle):
    c, ones, la = _case()
    n, m = c["counts"].shape
    p = c["x"].shape[1]
    for useQR in (True, False):
        a = beta_args(c, c["alpha0"], weights=ones, useQR=useQR)
        got, types = r_oracle.dot_call("_DESeq2_fitBeta", *_ordered(a))
        assert list(got) == BETA_NAMES and all(t == REALSXP for t in types.values())       # iter is numeric (:317)
        assert got["beta_mat"].shape == (n, p) and got["beta_var_mat"].shape == (n, p)
        assert got["hat_diagonals"].shape == (n, m) and got["iter"].shape == (n,) and got["deviance"].shape == (n,)
        assert got["contrast_num"].shape == (n, 1) and got["contrast_denom"].shape == (n, 1)   # arma::mat (:318-319)
        ref = oracle.fitBeta(**a)
        for k in BETA_NAMES:
            assert np.array_equal(got[k].reshape(np.shape(ref[k])), ref[k], equal_nan=True), k
    # getContrast's use (R/results.R:797-812): maxit = 0, a real contrast, beta_mat = the fitted coefficients
    fit = oracle.fitBeta(**beta_args(c, c["alpha0"], weights=ones))
    a0 = beta_args(c, c["alpha0"], beta0=fit["beta_mat"], maxit=0, contrast=np.array([0.0, 1.0]), weights=ones)
    got0, _ = r_oracle.dot_call("_DESeq2_fitBeta", *_ordered(a0))
    ref0 = oracle.fitBeta(**a0)
    assert np.array_equal(got0["contrast_num"].ravel(), ref0["contrast_num"].ravel())
    assert np.array_equal(got0["contrast_denom"].ravel(), ref0["contrast_denom"].ravel())
    assert np.array_equal(got0["beta_mat"], fit["beta_mat"])


def test_errors_become_R_errors(r_oracle):
    """A y that is neither integer nor numeric is refused with an R error, and the protect stack is clean after."""
    c, ones, la = _case()
    a = disp_args(c, c["mu"], la, weights=ones)
    bad = dict(a, ySEXP=c["counts"] > 0)                               # a logical matrix
    with pytest.raises(RuntimeError, match="integer or numeric"):
        r_oracle.dot_call("_DESeq2_fitDisp", *_ordered(bad))
    got, _ = r_oracle.dot_call("_DESeq2_fitDisp", *_ordered(a))       # the session is still usable
    assert np.all(np.isfinite(got["log_alpha"]))


def test_engine_shim_links_and_has_no_cpu_fallback(r_engine):
    """The shim links against the product library exactly as INTEGRATION.md section 2 says.  Without a GPU the
    engine's status code must surface as an R error carrying b200nb_last_error() -- never a silent CPU result."""
    import deseq2_b200
    if deseq2_b200.lib().b200nb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    c, ones, la = _case()
    with pytest.raises(RuntimeError, match=r"fitDisp \(b200nb\): "):
        r_engine.dot_call("_DESeq2_fitDisp", *_ordered(disp_args(c, c["mu"], la, weights=ones)))
    with pytest.raises(RuntimeError, match=r"fitBeta \(b200nb\): "):
        r_engine.dot_call("_DESeq2_fitBeta", *_ordered(beta_args(c, c["alpha0"], weights=ones)))


@pytest.mark.gpu
def test_engine_through_R_boundary(r_engine, oracle):
    """.Call('_DESeq2_fitDisp' / '_DESeq2_fitBeta' / '_DESeq2_fitDispGrid') on the CUDA engine vs the oracle: the same
    cases and the same acceptance rules as tests/test_parity_gpu.py (1e-6 relative, exact iteration counts), but
    entered through the R-facing symbols instead of the ctypes wrappers."""
    from test_parity_gpu import _compare_beta, _compare_disp
    c = make_case(800, 37, seed=13)
    a = disp_args(c, c["mu"], np.log(c["alpha0"]), weights=np.ones(c["counts"].shape))
    got, types = r_engine.dot_call("_DESeq2_fitDisp", *_ordered(a))
    assert list(got) == DISP_NAMES and types["iter"] == INTSXP and types["iter_accept"] == INTSXP
    _compare_disp(got, oracle.fitDisp(**a, with_margin=True), ".Call fitDisp")

    c = make_case(800, 37, seed=33)
    alpha = np.clip(0.1 + 4.0 / c["baseMean"], 1e-8, 10)
    b = beta_args(c, alpha, useQR=False, weights=np.ones(c["counts"].shape))
    gb, types = r_engine.dot_call("_DESeq2_fitBeta", *_ordered(b))
    assert list(gb) == BETA_NAMES and all(t == REALSXP for t in types.values())
    rb = oracle.fitBeta(**b)
    gb = {k: v.reshape(np.shape(rb[k])) for k, v in gb.items()}
    _compare_beta(gb, rb, ".Call fitBeta")

    c = make_case(400, 30, seed=23)
    grid = np.linspace(np.log(1e-8), np.log(30), 20)
    g = dict(ySEXP=c["counts"], xSEXP=c["x"], mu_hatSEXP=c["mu"], disp_gridSEXP=grid,
             log_alpha_prior_meanSEXP=np.log(0.1 + 4 / c["baseMean"]), log_alpha_prior_sigmasqSEXP=0.5,
             usePriorSEXP=True, weightsSEXP=np.ones(c["counts"].shape), useWeightsSEXP=False,
             weightThresholdSEXP=1e-2, useCRSEXP=True)
    gg, _ = r_engine.dot_call("_DESeq2_fitDispGrid", *_ordered(g))
    og = oracle.fitDispGrid(**g)["log_alpha"]
    assert np.mean(np.abs(gg["log_alpha"] - og) < 1e-9) > 0.995
    assert np.max(np.abs(gg["log_alpha"] - og)) < 2.0 * (grid[1] - grid[0]) / 9.5


def test_large_hat_matrix_lives_in_the_engines_pinned_pool(r_oracle, oracle):
    """hat_diagonals of >= 1 MB is allocated through allocVector3's custom allocator in the engine's page-locked pool
    (b200nb_host_alloc), so that the device-to-host copy is one DMA into the R object; R's gc returns the block
    (b200nb_host_free); if page-locked memory cannot be had the shim falls back to malloc.  Same numbers either way."""
    L = r_oracle.lib
    c = make_case(4200, 32, seed=6)                     # 4200 x 32 doubles > 1 MB
    n = len(c["counts"])
    ones = np.ones(c["counts"].shape)
    a = beta_args(c, c["alpha0"], weights=ones)
    ref = oracle.fitBeta(**a)
    a0, f0 = L.stub_host_allocs(), L.stub_host_frees()
    got, _ = r_oracle.dot_call("_DESeq2_fitBeta", *_ordered(a))
    assert L.stub_host_allocs() == a0 + 1                 # exactly the one large matrix
    assert got["hat_diagonals"].shape == (n, 32)
    assert np.array_equal(got["hat_diagonals"], ref["hat_diagonals"])
    L.mock_reset()                                        # R's gc
    assert L.stub_host_frees() == f0 + 1 and L.mock_custom_allocations() == 0
    L.stub_set_fail_alloc(1)                              # no page-locked memory: ordinary memory, same result
    try:
        got2, _ = r_oracle.dot_call("_DESeq2_fitBeta", *_ordered(a))
        assert np.array_equal(got2["hat_diagonals"], ref["hat_diagonals"])
        L.mock_reset()
        assert L.stub_host_frees() == f0 + 1
    finally:
        L.stub_set_fail_alloc(0)
