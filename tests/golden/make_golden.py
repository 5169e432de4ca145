"""Generate the golden input/output vectors in tests/golden/*.npz with the CPU oracle.

The reference stores no numeric fixtures for this path (its tests compare against independent R solvers,
SURVEY.md 8c) and cannot be run in this image (no R), so these vectors pin the ORACLE's outputs on fixed
seeded inputs; tests/test_oracle_reference_design.py pins the oracle itself against independent solvers.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import beta_args, disp_args, make_case  # noqa: E402
from deseq2_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def save(name, inputs, outputs):
    flat = {f"in_{k}": np.asarray(v) for k, v in inputs.items() if v is not None}
    flat.update({f"out_{k}": np.asarray(v) for k, v in outputs.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **flat)
    print(name, {k: v.shape for k, v in flat.items() if v.ndim > 0 and v.size > 4})


def main():
    rng = np.random.default_rng(2026)
    # 1. fitDisp MLE, ~condition, int counts
    c = make_case(96, 12, seed=101)
    a = disp_args(c, c["mu"], np.log(c["alpha0"]))
    save("fitdisp_mle_p2", a, O.fitDisp(**a, with_margin=True))
    # 2. fitDisp MAP with observation weights (some below the Cox-Reid threshold), ~batch+condition (p=3)
    x3 = synth.design_batch_condition(12, n_batch=2)
    c3 = make_case(64, 12, x=x3, seed=102)
    alpha3 = np.clip(0.1 + 4 / c3["baseMean"], 1e-8, 12)
    mu3 = np.maximum(c3["nf"] * np.exp(O.fitBeta(**beta_args(c3, alpha3))["beta_mat"] @ x3.T), 0.5)
    w = rng.uniform(0.3, 1.0, c3["counts"].shape)
    w[rng.random(w.shape) < 0.08] = 1e-3
    w = np.maximum(w / w.max(axis=1, keepdims=True), 1e-6)
    a = disp_args(c3, mu3, np.log(alpha3), prior_mean=np.log(0.1 + 4 / c3["baseMean"]), sigmasq=0.7, usePrior=True,
                  weights=w, useWeights=True)
    save("fitdisp_map_weights_p3", a, O.fitDisp(**a, with_margin=True))
    # 3. fitBeta, QR branch
    alpha = np.clip(0.1 + 4 / c["baseMean"], 1e-8, 12)
    b = beta_args(c, alpha)
    save("fitbeta_p2", b, O.fitBeta(**b))
    # 4. fitBeta with weights, p=3, normal-equation branch, ridge on the last coefficient
    lam = np.array([1e-6, 1e-6, 0.5]) / np.log(2) ** 2
    b = beta_args(c3, alpha3, weights=w, useWeights=True, useQR=False, lam=lam, contrast=np.array([0, 0, 1.0]))
    save("fitbeta_weights_p3", b, O.fitBeta(**b))
    # 5. fitDispGrid
    c5 = make_case(32, 12, seed=103)
    kw = dict(ySEXP=c5["counts"], xSEXP=c5["x"], mu_hatSEXP=c5["mu"], disp_gridSEXP=np.linspace(np.log(1e-8), np.log(12), 20),
              log_alpha_prior_meanSEXP=np.log(0.1 + 4 / c5["baseMean"]), log_alpha_prior_sigmasqSEXP=0.5,
              usePriorSEXP=True, weightsSEXP=None, useWeightsSEXP=False, weightThresholdSEXP=1e-2, useCRSEXP=True)
    save("fitdispgrid_p2", kw, O.fitDispGrid(**kw))


if __name__ == "__main__":
    main()
