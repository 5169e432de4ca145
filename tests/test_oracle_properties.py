"""Size-independent properties of the oracle (the checker must itself be trustworthy): sample-permutation
invariance, gene independence, scale equivariance of the size factors, monotone Armijo ascent."""
import numpy as np

from helpers import beta_args, disp_args, make_case


def test_sample_permutation_invariance(oracle):
    c = make_case(60, 14, seed=91)
    rng = np.random.default_rng(1)
    perm = rng.permutation(14)
    alpha = np.clip(0.1 + 4 / c["baseMean"], 1e-8, 14)
    a = oracle.fitBeta(**beta_args(c, alpha))
    cp = dict(c)
    cp["counts"], cp["nf"], cp["x"] = c["counts"][:, perm], c["nf"][:, perm], c["x"][perm]
    b = oracle.fitBeta(**beta_args(cp, alpha))
    assert np.array_equal(a["iter"], b["iter"])
    assert np.allclose(a["beta_mat"], b["beta_mat"], rtol=1e-9, atol=1e-11)
    assert np.allclose(a["hat_diagonals"][:, perm], b["hat_diagonals"], rtol=1e-8)
    d1 = oracle.fitDisp(**disp_args(c, c["mu"], np.log(c["alpha0"])), with_margin=True)
    d2 = oracle.fitDisp(**disp_args(cp, c["mu"][:, perm], np.log(c["alpha0"])), with_margin=True)
    ok = (d1["margin"] > 64) & (d2["margin"] > 64)
    assert ok.mean() > 0.7 and np.array_equal(d1["iter"][ok], d2["iter"][ok])
    assert np.allclose(d1["log_alpha"][ok], d2["log_alpha"][ok], rtol=1e-8, atol=1e-10)


def test_offset_equivariance(oracle):
    """Doubling every normalisation factor shifts the intercept by -log 2 and changes nothing else."""
    c = make_case(50, 12, seed=92)
    alpha = np.clip(0.1 + 4 / c["baseMean"], 1e-8, 12)
    a = oracle.fitBeta(**beta_args(c, alpha))
    b0 = c["beta0"].copy()
    b0[:, 0] -= np.log(2.0)
    b = oracle.fitBeta(**beta_args(c, alpha, nf=2.0 * c["nf"], beta0=b0))
    conv = (a["iter"] < 100) & (b["iter"] < 100) & (c["counts"].min(axis=1) > 2)     # away from the minmu clamp
    assert np.allclose(a["beta_mat"][conv, 0] - np.log(2.0), b["beta_mat"][conv, 0], atol=1e-6)
    assert np.allclose(a["beta_mat"][conv, 1], b["beta_mat"][conv, 1], atol=1e-6)
    assert np.allclose(a["deviance"][conv], b["deviance"][conv], rtol=1e-7)


def test_line_search_never_decreases_the_posterior(oracle):
    c = make_case(200, 10, seed=93)
    d = oracle.fitDisp(**disp_args(c, c["mu"], np.log(c["alpha0"])))
    moved = d["iter_accept"] > 0
    assert np.all(d["last_lp"][moved] >= d["initial_lp"][moved] - 1e-9 * np.abs(d["initial_lp"][moved]))
    assert np.all(d["iter_accept"] <= d["iter"]) and np.all(d["iter"] <= 100)
    # no accepted step => log_alpha unchanged
    still = d["iter_accept"] == 0
    assert np.array_equal(d["log_alpha"][still], np.log(c["alpha0"])[still])


def test_grid_matches_dense_evaluation(oracle):
    """fitDispGrid returns a point of its own fine grid that maximises the posterior over that grid."""
    c = make_case(12, 10, seed=94)
    grid = np.linspace(np.log(1e-8), np.log(10), 20)
    kw = dict(ySEXP=c["counts"], xSEXP=c["x"], mu_hatSEXP=c["mu"], disp_gridSEXP=grid,
              log_alpha_prior_meanSEXP=np.zeros(len(c["counts"])), log_alpha_prior_sigmasqSEXP=1.0, usePriorSEXP=False,
              weightsSEXP=None, useWeightsSEXP=False, weightThresholdSEXP=1e-2, useCRSEXP=True)
    g = oracle.fitDispGrid(**kw)["log_alpha"]
    delta = grid[1] - grid[0]
    for i in range(len(g)):
        lp = np.array([oracle.log_posterior_row(c["counts"][i], c["mu"][i], c["x"], a) for a in grid])
        a_hat = grid[np.argmax(lp)]
        fine = np.linspace(a_hat - delta, a_hat + delta, 20)
        lpf = np.array([oracle.log_posterior_row(c["counts"][i], c["mu"][i], c["x"], a) for a in fine])
        assert abs(g[i] - fine[np.argmax(lpf)]) < 1e-12
