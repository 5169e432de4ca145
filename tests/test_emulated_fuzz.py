"""Randomised shapes and options through the emulated engine (the CUDA kernels' source on CPU fibers, tests/simt_emu/)
against the oracle.  Fixed seeds, so the test is deterministic; the generator covers what hand-written cases tend to
miss: m around the 4- / 32- / 128-sample unroll boundaries, m barely above p, 1-gene and 1-sample problems, p from 1
to 32 (register-resident p <= 4 kernels and the shared-memory general-p kernels up to B200NB_MAX_P), factor and continuous designs,
integer / double / non-integer counts, all-zero genes, observation weights with entries below the Cox-Reid threshold,
prior and CR switches, tiny maxit, ridge penalties, size-factor vs gene-wise normalisation, both QR flags.
A 400-seed run of the same generator (scripts/fuzz_emulated.py) found no discrepancy on valid inputs."""
import numpy as np
import pytest

from helpers import rel_err

M_CHOICES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 17, 31, 32, 33, 40, 63, 64, 65, 70, 127, 129, 200]


def make_problem(s):
    rng = np.random.default_rng(s)
    n = int(rng.integers(1, 30))
    m = int(rng.choice(M_CHOICES))
    p = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 12, 32]))
    if m <= p:
        p = max(1, m - 1)
    kind = str(rng.choice(["factor", "cov"]))
    if p == 1:
        x = np.ones((m, 1))
    elif kind == "factor":
        g = np.arange(m) % p
        x = np.zeros((m, p))
        x[:, 0] = 1
        for k in range(1, p):
            x[g == k, k] = 1
    else:
        x = np.c_[np.ones(m), rng.normal(0, 1, (m, p - 1))]
    if np.linalg.matrix_rank(x) < p:
        return None
    base = 10 ** rng.uniform(-0.5, 4, (n, 1))
    beta = rng.normal(0, 0.5, (n, p))
    beta[:, 0] = 0
    sf = np.exp(rng.normal(0, 0.3, m))
    mu_true = base * np.exp(beta @ x.T) * sf
    alpha_true = 10 ** rng.uniform(-3, 0.5, (n, 1))
    y = rng.negative_binomial(1 / alpha_true, 1 / (1 + mu_true * alpha_true)).astype(np.int32)
    if rng.random() < 0.2:
        y[rng.integers(0, n)] = 0
    yy = y
    if rng.random() < 0.3:
        yy = y.astype(np.float64) + (0.5 if rng.random() < 0.15 else 0.0)
    useW = bool(rng.random() < 0.4)
    w = rng.uniform(0.05, 1, (n, m)) if useW else None
    if useW and rng.random() < 0.5:
        w[rng.random((n, m)) < 0.1] = 1e-3
    mu = np.maximum(mu_true * np.exp(rng.normal(0, 0.1, (n, m))), 0.5)
    la = np.log(np.clip(alpha_true[:, 0] * np.exp(rng.normal(0, 1, n)), 1e-8, max(10, m)))
    usePrior, useCR = bool(rng.random() < 0.5), bool(rng.random() < 0.85)
    disp = dict(ySEXP=yy, xSEXP=x, mu_hatSEXP=mu, log_alphaSEXP=la, log_alpha_prior_meanSEXP=la + rng.normal(0, .5, n),
                log_alpha_prior_sigmasqSEXP=float(rng.uniform(0.3, 2)), min_log_alphaSEXP=np.log(1e-9), kappa_0SEXP=1.0,
                tolSEXP=1e-6, maxitSEXP=int(rng.choice([1, 3, 100])), usePriorSEXP=usePrior, weightsSEXP=w,
                useWeightsSEXP=useW, weightThresholdSEXP=1e-2, useCRSEXP=useCR)
    grid = dict(ySEXP=yy, xSEXP=x, mu_hatSEXP=mu, disp_gridSEXP=np.linspace(np.log(1e-8), np.log(max(10, m)), 20),
                log_alpha_prior_meanSEXP=disp["log_alpha_prior_meanSEXP"],
                log_alpha_prior_sigmasqSEXP=disp["log_alpha_prior_sigmasqSEXP"], usePriorSEXP=usePrior, weightsSEXP=w,
                useWeightsSEXP=useW, weightThresholdSEXP=1e-2, useCRSEXP=useCR)
    lam = np.full(p, 1e-6) / np.log(2) ** 2
    if rng.random() < 0.3 and p > 1:
        lam[1:] = rng.uniform(0.1, 3, p - 1)
    nf = np.broadcast_to(sf, (n, m)).copy() if rng.random() < 0.6 else np.exp(rng.normal(0, 0.3, (n, m)))
    b0 = np.zeros((n, p))
    b0[:, 0] = np.log(np.maximum((y / sf).mean(axis=1), 0.1))
    beta_kw = dict(ySEXP=yy, xSEXP=x, nfSEXP=nf, alpha_hatSEXP=np.clip(alpha_true[:, 0], 1e-8, 10),
                   contrastSEXP=rng.normal(0, 1, p), beta_matSEXP=b0, lambdaSEXP=lam, weightsSEXP=w, useWeightsSEXP=useW,
                   tolSEXP=1e-8, maxitSEXP=int(rng.choice([0, 2, 100])), useQRSEXP=bool(rng.random() < 0.5),
                   minmuSEXP=0.5)
    # Genes whose Cox-Reid matrix is rank deficient once the rows with weight <= threshold are dropped have an undefined
    # (log 0 / NaN) posterior in the reference too; R refuses such weights up front (getAndCheckWeights,
    # R/core.R:2697-2752), so they are not part of the contract.
    valid = np.ones(n, bool)
    if useW and useCR:
        for i in range(n):
            xs = x[w[i] > 1e-2]
            xs = xs[:, np.any(xs != 0, axis=0)] if xs.size else xs
            valid[i] = (xs.size > 0 and xs.shape[1] > 0 and np.linalg.matrix_rank(xs) == xs.shape[1]
                        and np.linalg.cond(xs) < 1e6)
    tag = f"seed {s}: n={n} m={m} p={p} {kind} weights={useW} prior={usePrior} CR={useCR} y={yy.dtype}"
    return dict(disp=disp, grid=grid, beta=beta_kw, valid=valid, n=n, usePrior=usePrior, tag=tag)


def check_problem(E, O, P, robust):
    n, valid, tag = P["n"], P["valid"], P["tag"]
    g, o = E.fitDisp(**P["disp"]), O.fitDisp(**P["disp"], with_margin=True)
    ok = (o["margin"] > robust) & np.isfinite(o["last_lp"]) & valid
    assert np.array_equal(g["iter"][ok], o["iter"][ok]), (tag, "iter", g["iter"], o["iter"], o["margin"])
    assert np.array_equal(g["iter_accept"][ok], o["iter_accept"][ok]), (tag, "iter_accept")
    for k, floor in (("log_alpha", 1e-6), ("initial_lp", 1e-3), ("last_lp", 1e-3)):
        e = rel_err(g[k][ok], o[k][ok], floor=floor)
        assert e.size == 0 or np.nanmax(e) < 1e-6, (tag, k, np.nanmax(e))
    assert np.array_equal(np.isfinite(g["last_lp"])[valid], np.isfinite(o["last_lp"])[valid]), (tag, "finiteness")
    gg, og = E.fitDispGrid(**P["grid"])["log_alpha"], O.fitDispGrid(**P["grid"])["log_alpha"]
    f = np.isfinite(og) & valid
    if P["usePrior"] and f.sum() > 4:      # without a prior the posterior is flat to rounding as alpha -> 0: ties
        assert np.mean(np.abs(gg[f] - og[f]) < 1e-6) >= 0.8, (tag, "grid", gg, og)
    kb = P["beta"]
    gb, ob = E.fitBeta(**kb), O.fitBeta(**kb)
    conv = (ob["iter"] < max(kb["maxitSEXP"], 1)) | (kb["maxitSEXP"] == 0)
    conv &= np.all(np.isfinite(ob["beta_mat"]), axis=1) & np.all(np.abs(ob["beta_mat"]) < 25, axis=1)
    same = gb["iter"] == ob["iter"]
    if conv.sum() > 5:
        assert np.mean(same[conv]) >= 0.9, (tag, "beta iter", gb["iter"], ob["iter"])
    sel = conv & same
    for k in ("beta_mat", "beta_var_mat", "hat_diagonals", "deviance", "contrast_num", "contrast_denom"):
        a, b = np.asarray(gb[k]).reshape(n, -1)[sel], np.asarray(ob[k]).reshape(n, -1)[sel]
        e = np.abs(a - b) / (np.abs(b) + 1e-6 * (1 + np.max(np.abs(b), axis=1, keepdims=True)))
        assert e.size == 0 or np.nanmax(e) < 2e-6, (tag, k, np.nanmax(e))


@pytest.mark.parametrize("block", range(8))
def test_emulated_fuzz(emu, oracle, block):
    from test_parity_gpu import ROBUST
    for s in range(7000 + 12 * block, 7000 + 12 * (block + 1)):
        P = make_problem(s)
        if P is not None:
            check_problem(emu, oracle, P, ROBUST)


@pytest.mark.parametrize("order", ["reverse", "random:11"])
def test_emulated_fuzz_is_lane_order_independent(order):
    """The emulator gives the CPU to the lanes of a warp in a configurable order between barriers (SIMT_EMU_ORDER,
    fixed when the library is loaded, hence a subprocess).  The hardware promises no order, so the kernels must give
    the same answers top-down and shuffled as bottom-up; a shared-memory hand-off without a barrier would not."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_emulated.py"), "9000", "30"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, SIMT_EMU_ORDER=order))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().endswith("30 seeds, 0 failures"), r.stdout[-2000:]
