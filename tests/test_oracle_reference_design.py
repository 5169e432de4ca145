"""The oracle pinned the way the reference pins its own native code: agreement
with independent solvers at 1e-6 -- tests/testthat/test_betaFitting.R, test_dispersions.R, test_QR.R,
test_optim.R, test_weights.R, test_parallel.R -- with our own seeds (the reference stores no golden numbers)."""
import numpy as np
import scipy.optimize as so
import scipy.special as sp

from helpers import beta_args, disp_args, make_case

LN2 = np.log(2.0)


def _one_gene(seed=1, m=10):
    rng = np.random.default_rng(seed)
    y = rng.poisson(20, m).astype(float)[None, :]
    x = np.c_[np.ones(m), np.r_[np.zeros(m // 2), np.ones(m - m // 2)]]
    return y, x


def test_beta_irls_equals_textbook_and_optim(oracle):
    """test_betaFitting.R:2-47: fitBeta == 100 steps of textbook IRLS == Nelder-Mead on the penalised likelihood."""
    y, x = _one_gene()
    m = y.shape[1]
    alpha = 0.5
    lam = np.array([1e-6, 2.0]) / LN2 ** 2
    r = oracle.fitBeta(y, x, np.ones((1, m)), [alpha], [1, 0], [[np.log(20.0), 0.0]], lam, None, False, 1e-8, 100,
                       True, 0.5)
    b = np.array([np.log(20.0), 0.0])
    for _ in range(100):
        mu = np.exp(x @ b)
        w = mu / (1 + alpha * mu)
        z = np.log(mu) + (y[0] - mu) / mu
        b = np.linalg.solve((x.T * w) @ x + np.diag(lam), x.T @ (w * z))
    assert np.allclose(r["beta_mat"][0], b, rtol=0, atol=1e-6)

    def obj(bb):
        mu = np.exp(x @ bb)
        size = 1 / alpha
        ll = np.sum(sp.gammaln(y[0] + size) - sp.gammaln(size) - sp.gammaln(y[0] + 1) + size * np.log(size / (size + mu))
                    + y[0] * np.log(mu / (size + mu)))
        return -(ll - 0.5 * np.sum(lam * bb * bb))

    res = so.minimize(obj, [3.0, 0.0], method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-14, maxiter=20000))
    assert np.allclose(r["beta_mat"][0], res.x, atol=1e-6)
    # standard errors = sqrt(diag((X'WX+L)^-1 X'WX (X'WX+L)^-1))
    mu = np.exp(x @ b)
    w = mu / (1 + alpha * mu)
    A = (x.T * w) @ x
    Ai = np.linalg.inv(A + np.diag(lam))
    assert np.allclose(r["beta_var_mat"][0], np.diag(Ai @ A @ Ai), rtol=1e-6)


def test_disp_map_equals_brent_and_derivatives(oracle):
    """test_dispersions.R:35-111: fitDisp MAP == 1-d optimiser of the same posterior; analytic == numeric derivs."""
    y, x = _one_gene(seed=4)
    rng = np.random.default_rng(5)
    y = rng.negative_binomial(2.0, 2.0 / (2.0 + 30.0), y.shape[1]).astype(float)[None, :]
    mu = np.full_like(y, y.mean())
    pm, s2 = np.log(0.2), 1.0
    d = oracle.fitDisp(y, x, mu, [np.log(0.5)], [pm], s2, np.log(1e-9), 1.0, 1e-16, 100, True, None, False, 1e-2, True)
    f = lambda a: -oracle.log_posterior_row(y[0], mu[0], x, a, pm, s2, True)
    br = so.minimize_scalar(f, bounds=(-10, 5), method="bounded", options=dict(xatol=1e-12))
    assert abs(d["log_alpha"][0] - br.x) < 1e-6
    a0, h = np.log(0.5), 1e-5
    assert abs(d["initial_dlp"][0] - (f(a0 - h) - f(a0 + h)) / (2 * h)) < 1e-6 * max(1, abs(d["initial_dlp"][0]))
    a1, h = d["log_alpha"][0], 1e-4
    num = -(f(a1 + h) - 2 * f(a1) + f(a1 - h)) / h ** 2
    assert abs(d["last_d2lp"][0] - num) < 1e-5 * max(1, abs(num))
    # derivative helper agrees with fitDisp's own report
    assert abs(oracle.log_posterior_row(y[0], mu[0], x, a0, pm, s2, True, deriv=1) - d["initial_dlp"][0]) < 1e-12


def test_qr_equals_normal_equations(oracle):
    """test_QR.R:2-9."""
    c = make_case(100, 12, seed=3)
    alpha = np.clip(0.1 + 4 / c["baseMean"], 1e-8, 12)
    a = oracle.fitBeta(**beta_args(c, alpha, useQR=True))
    b = oracle.fitBeta(**beta_args(c, alpha, useQR=False))
    assert np.allclose(a["beta_mat"], b["beta_mat"], atol=1e-6)
    assert np.array_equal(a["iter"], b["iter"])


def test_divergence_sentinel(oracle):
    """test_optim.R:29-39: a 0/1000 row must come back with iter == maxit."""
    y = np.array([[0, 0, 0, 0, 0, 1000, 1000, 0, 0, 0]], dtype=float)
    x = np.c_[np.ones(10), np.r_[np.zeros(5), np.ones(5)]]
    r = oracle.fitBeta(y, x, np.ones((1, 10)), [0.1], [1, 0], [[1.0, 1.0]], np.full(2, 1e-6) / LN2 ** 2, None, False,
                       1e-8, 100, True, 0.5)
    assert r["iter"][0] == 100


def test_weight_zero_equals_dropped_sample(oracle):
    """test_weights.R:6-19: weight 0 == sample removed, for coefficients, SE and deviance."""
    c = make_case(40, 10, seed=6)
    alpha = np.clip(0.1 + 4 / c["baseMean"], 1e-8, 10)
    w = np.ones(c["counts"].shape)
    w[:, 0] = 0.0
    a = oracle.fitBeta(**beta_args(c, alpha, weights=w, useWeights=True))
    sub = dict(c)
    sub["counts"], sub["nf"], sub["x"] = c["counts"][:, 1:], c["nf"][:, 1:], c["x"][1:]
    b = oracle.fitBeta(**beta_args(sub, alpha))
    assert np.allclose(a["beta_mat"], b["beta_mat"], atol=1e-8)
    assert np.allclose(a["beta_var_mat"], b["beta_var_mat"], rtol=1e-7)
    assert np.allclose(a["deviance"], b["deviance"], rtol=1e-9)


def test_chunked_equals_whole(oracle):
    """test_parallel.R:12-37: gene chunks processed separately == one call (genes are independent)."""
    c = make_case(60, 8, seed=7)
    a = disp_args(c, c["mu"], np.log(c["alpha0"]))
    whole = oracle.fitDisp(**a)
    idx = np.array_split(np.arange(len(c["counts"])), 3)
    for k in ("log_alpha", "iter", "last_lp"):
        parts = []
        for ii in idx:
            sub = dict(a)
            for key in ("ySEXP", "mu_hatSEXP", "log_alphaSEXP", "log_alpha_prior_meanSEXP"):
                sub[key] = a[key][ii]
            parts.append(oracle.fitDisp(**sub)[k])
        assert np.array_equal(np.concatenate(parts), whole[k])


def test_grid_brackets_the_optimum(oracle):
    c = make_case(30, 12, seed=8)
    grid = np.linspace(np.log(1e-8), np.log(12), 20)
    kw = dict(ySEXP=c["counts"], xSEXP=c["x"], mu_hatSEXP=c["mu"], disp_gridSEXP=grid,
              log_alpha_prior_meanSEXP=np.zeros(len(c["counts"])), log_alpha_prior_sigmasqSEXP=1.0, usePriorSEXP=False,
              weightsSEXP=None, useWeightsSEXP=False, weightThresholdSEXP=1e-2, useCRSEXP=True)
    g = oracle.fitDispGrid(**kw)["log_alpha"]
    a = disp_args(c, c["mu"], np.log(c["alpha0"]), tol=1e-12)
    ls = oracle.fitDisp(**a)
    ok = (ls["iter"] < 100) & (ls["log_alpha"] > np.log(1e-7))
    delta = grid[1] - grid[0]
    assert np.all(np.abs(g[ok] - ls["log_alpha"][ok]) <= delta * (2.0 / 19) + 1e-3)
