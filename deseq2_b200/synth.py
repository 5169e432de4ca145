"""Synthetic count matrices with the semantics of makeExampleDESeqDataSet (R/core.R:459-498):
beta0 ~ N(interceptMean=4, interceptSD=2) on the log2 scale, other betas ~ N(0, betaSD), dispersion
alpha_i = 4 / 2^beta0 + 0.1, K_ij ~ NB(mu_ij = s_j 2^(x_j beta_i), size = 1/alpha_i), int32.
numpy Generator(PCG64(seed)); the reference uses R's RNG, so values differ but the law is the same.
Design matrices for the BASELINE.json configs are built here too.
"""
from __future__ import annotations

import numpy as np


def design_condition(m: int) -> np.ndarray:
    """~condition, two levels, ceil(m/2) A then floor(m/2) B (R/core.R:463-465): columns [Intercept, B_vs_A]."""
    nA = (m + 1) // 2
    return np.c_[np.ones(m), np.r_[np.zeros(nA), np.ones(m - nA)]]


def design_batch_condition(m: int, n_batch: int = 3) -> np.ndarray:
    """~batch + condition: n_batch batches x 2 conditions, balanced, treatment contrasts: p = n_batch + 1."""
    batch = np.arange(m) % n_batch
    cond = (np.arange(m) // n_batch) % 2
    cols = [np.ones(m)] + [(batch == b).astype(float) for b in range(1, n_batch)] + [cond.astype(float)]
    return np.stack(cols, axis=1)


def design_factor(m: int, levels: int) -> np.ndarray:
    """~group with `levels` levels (treatment contrasts): p = levels."""
    g = (np.arange(m) * levels) // m
    cols = [np.ones(m)] + [(g == l).astype(float) for l in range(1, levels)]
    return np.stack(cols, axis=1)


def design_factor_expanded(m: int, levels: int) -> np.ndarray:
    """Expanded model matrix (R/expanded.R:1-18): intercept + one indicator per level: p = levels + 1, rank levels."""
    g = (np.arange(m) * levels) // m
    cols = [np.ones(m)] + [(g == l).astype(float) for l in range(levels)]
    return np.stack(cols, axis=1)


def make_example_counts(n: int, m: int, x: np.ndarray | None = None, seed: int = 20260923, betaSD: float = 1.0,
                        interceptMean: float = 4.0, interceptSD: float = 2.0, size_factor_sd: float = 0.25,
                        sizeFactors: np.ndarray | None = None):
    """Returns dict(counts int32 n x m, x, sizeFactors, trueBeta (log2), trueDisp).  `sizeFactors` fixes the size
    factors (gene shards of one experiment generated on different ranks share them)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if x is None:
        x = design_condition(m)
    p = x.shape[1]
    beta = np.empty((n, p))
    beta[:, 0] = rng.normal(interceptMean, interceptSD, n)
    for k in range(1, p):
        beta[:, k] = rng.normal(0.0, betaSD, n)
    disp = 4.0 / 2.0 ** beta[:, 0] + 0.1
    sf = np.exp(rng.normal(0.0, size_factor_sd, m)) if size_factor_sd > 0 else np.ones(m)
    sf = sf / np.exp(np.mean(np.log(sf)))
    if sizeFactors is not None:
        sf = np.asarray(sizeFactors, dtype=np.float64)
    mu = (2.0 ** (beta @ x.T)) * sf[None, :]
    size = 1.0 / disp[:, None]
    prob = size / (size + mu)
    counts = rng.negative_binomial(np.broadcast_to(size, mu.shape), prob)
    counts = np.minimum(counts, np.iinfo(np.int32).max).astype(np.int32)
    return {"counts": counts, "x": x, "sizeFactors": sf, "trueBeta": beta, "trueDisp": disp}
