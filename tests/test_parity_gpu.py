"""GPU parity tests: the CUDA engine (through the C ABI, host buffers, exactly what the R shim would pass)
against the CPU oracle on identical seeded inputs.  Tolerances: 1e-6 relative on every floating-point output
(north_star), exact on iteration counters for every gene whose line-search decisions are not knife-edge
(oracle `margin` > 1e-9: two correct fp64 evaluations cannot legitimately disagree there)."""
import numpy as np
import pytest

from helpers import DISP_KEYS, beta_args, disp_args, make_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-6


def test_device_special_functions(engine):
    import ctypes as C

    import mpmath as mp
    from deseq2_b200 import _lib
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(1e-4, 40, 4000), 10 ** rng.uniform(-6, 9, 4000), np.arange(1, 200) * 0.5])
    lg, dg, tg = (np.empty_like(x) for _ in range(3))
    P = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(_lib.lib().b200nb_test_special(P(x), len(x), P(lg), P(dg), P(tg)), "test_special")
    mp.mp.dps = 40
    idx = rng.choice(len(x), 600, replace=False)
    for i in idx:
        xi = mp.mpf(float(x[i]))
        rl, rd, rt = float(mp.loggamma(xi)), float(mp.digamma(xi)), float(mp.polygamma(1, xi))
        # absolute accuracy on the scale max(1,|f|): lgamma/digamma cross zero at x = 1, 2 / 1.4616
        assert abs(lg[i] - rl) <= 2e-14 * max(1.0, abs(rl)), (x[i], lg[i], rl)
        assert abs(dg[i] - rd) <= 2e-14 * max(1.0, abs(rd)), (x[i], dg[i], rd)
        assert abs(tg[i] - rt) <= 2e-14 * max(1.0, abs(rt)), (x[i], tg[i], rt)


ROBUST = 64.0   # decision margin, in units of the a-priori fp64 rounding bound (see oracle_fit_disp)


def _compare_disp(g, o, name, min_robust=0.9):
    robust = o["margin"] > ROBUST
    same = (g["iter"] == o["iter"]) & (g["iter_accept"] == o["iter_accept"])
    # every gene whose decisions are not knife-edge must follow the oracle's control flow exactly
    bad = np.flatnonzero(~same & robust)
    assert bad.size == 0, (f"{name}: {bad.size} robust genes differ in iter/iter_accept, e.g. gene {bad[:5]} "
                           f"gpu iter {g['iter'][bad[:5]]} oracle iter {o['iter'][bad[:5]]} "
                           f"margin {o['margin'][bad[:5]]}")
    assert robust.mean() >= min_robust, f"{name}: only {robust.mean():.3f} of genes have robust decisions"
    sel = robust & same
    for k in DISP_KEYS:
        if k in ("last_dlp", "last_change", "last_d2lp", "initial_dlp"):
            # derivatives / differences of nearly equal numbers: absolute error on the scale of the posterior
            err = np.abs(g[k][sel] - o[k][sel]) / (np.abs(o["last_lp"][sel]) * 1e-3 + np.abs(o[k][sel]) + 1e-12)
        else:
            err = rel_err(g[k][sel], o[k][sel])
        assert np.nanmax(err) < TOL, f"{name}: {k} max rel err {np.nanmax(err):.3e}"
    assert np.max(rel_err(np.exp(g["log_alpha"][sel]), np.exp(o["log_alpha"][sel]))) < TOL
    # knife-edge genes (either side may take the other branch): both must still end at the same posterior value
    rest = ~sel
    if rest.any():
        d = np.abs(g["last_lp"][rest] - o["last_lp"][rest]) / (1.0 + np.abs(o["last_lp"][rest]))
        assert np.nanmax(d) < 1e-5, f"{name}: knife-edge genes end at different posterior values ({np.nanmax(d):.2e})"
    return robust.mean(), same.mean()


@pytest.mark.parametrize("n,m,seed", [(3000, 100, 11), (1500, 6, 12), (800, 37, 13)])
def test_fit_disp_mle_parity(engine, oracle, n, m, seed):
    c = make_case(n, m, seed=seed)
    a = disp_args(c, c["mu"], np.log(c["alpha0"]))
    g = engine.fitDisp(**a)
    o = oracle.fitDisp(**a, with_margin=True)
    _compare_disp(g, o, f"mle {n}x{m}", min_robust=0.9 if m >= 30 else 0.3)


def test_fit_disp_map_parity(engine, oracle):
    c = make_case(3000, 100, seed=21)
    a0 = disp_args(c, c["mu"], np.log(c["alpha0"]))
    mle = oracle.fitDisp(**a0)
    fit = 0.1 + 4.0 / c["baseMean"]
    a = disp_args(c, c["mu"], mle["log_alpha"], prior_mean=np.log(fit), sigmasq=0.6, usePrior=True)
    _compare_disp(engine.fitDisp(**a), oracle.fitDisp(**a, with_margin=True), "map")


def test_fit_disp_f64_counts_and_no_cr(engine, oracle):
    c = make_case(500, 20, seed=22)
    a = disp_args(c, c["mu"], np.log(c["alpha0"]), useCR=False, y=c["counts"].astype(np.float64))
    _compare_disp(engine.fitDisp(**a), oracle.fitDisp(**a, with_margin=True), "f64/noCR", min_robust=0.5)


def test_fit_disp_grid_parity(engine, oracle):
    c = make_case(400, 30, seed=23)
    grid = np.linspace(np.log(1e-8), np.log(30), 20)
    kw = dict(ySEXP=c["counts"], xSEXP=c["x"], mu_hatSEXP=c["mu"], disp_gridSEXP=grid,
              log_alpha_prior_meanSEXP=np.log(0.1 + 4 / c["baseMean"]), log_alpha_prior_sigmasqSEXP=0.5,
              usePriorSEXP=True, weightsSEXP=None, useWeightsSEXP=False, weightThresholdSEXP=1e-2, useCRSEXP=True)
    g = engine.fitDispGrid(**kw)["log_alpha"]
    o = oracle.fitDispGrid(**kw)["log_alpha"]
    # argmax on a grid: identical grid point unless two grid values tie to rounding
    assert np.mean(np.abs(g - o) < 1e-9) > 0.995
    assert np.max(np.abs(g - o)) < 2.0 * (grid[1] - grid[0]) / 9.5


def _compare_beta(g, o, name, tol=TOL):
    assert np.array_equal(g["iter"], o["iter"]), f"{name}: iter differs on {np.sum(g['iter'] != o['iter'])} genes"
    for k in ("beta_mat", "beta_var_mat", "hat_diagonals", "contrast_num", "contrast_denom", "deviance"):
        err = rel_err(g[k], o[k], floor=1e-8)
        assert np.nanmax(err) < tol, f"{name}: {k} max rel err {np.nanmax(err):.3e}"
    se = rel_err(np.sqrt(g["beta_var_mat"]), np.sqrt(o["beta_var_mat"]))
    assert np.nanmax(se) < tol


@pytest.mark.parametrize("n,m,seed,useQR", [(3000, 100, 31, True), (1500, 6, 32, True), (800, 37, 33, False)])
def test_fit_beta_parity(engine, oracle, n, m, seed, useQR):
    c = make_case(n, m, seed=seed)
    alpha = np.clip(0.1 + 4.0 / c["baseMean"], 1e-8, 10)
    a = beta_args(c, alpha, useQR=useQR)
    _compare_beta(engine.fitBeta(**a), oracle.fitBeta(**a), f"beta {n}x{m}")


def test_fit_beta_maxit0_contrast(engine, oracle):
    """results() re-entry (R/results.R:797-807): maxit=0, numeric contrast, only the covariance block runs."""
    c = make_case(600, 24, seed=34)
    alpha = np.clip(0.1 + 4.0 / c["baseMean"], 1e-8, 10)
    fit = oracle.fitBeta(**beta_args(c, alpha))
    a = beta_args(c, alpha, beta0=fit["beta_mat"], maxit=0, useQR=False, contrast=np.array([0.0, 1.0]))
    g, o = engine.fitBeta(**a), oracle.fitBeta(**a)
    assert np.all(g["iter"] == 0)
    _compare_beta(g, o, "maxit0")
