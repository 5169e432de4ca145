"""Time the three host-buffer C-ABI calls separately (e2e breakdown)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from deseq2_b200 import wrappers as W
w = bench.build_workload(50000, 100, 20260925, W)
for rep in range(3):
    c, x, mu = w["counts"], w["x"], w["mu"]
    common = dict(ySEXP=c, xSEXP=x, mu_hatSEXP=mu, min_log_alphaSEXP=bench.MIN_LOG_ALPHA, kappa_0SEXP=1.0, tolSEXP=1e-6,
                  maxitSEXP=100, weightsSEXP=None, useWeightsSEXP=False, weightThresholdSEXP=1e-2, useCRSEXP=True)
    t0 = time.perf_counter()
    W.fitDisp(log_alphaSEXP=w["log_alpha0"], log_alpha_prior_meanSEXP=w["log_alpha0"], log_alpha_prior_sigmasqSEXP=1.0, usePriorSEXP=False, **common)
    t1 = time.perf_counter()
    W.fitDisp(log_alphaSEXP=w["log_dispInit"], log_alpha_prior_meanSEXP=w["log_dispFit"], log_alpha_prior_sigmasqSEXP=w["priorVar"], usePriorSEXP=True, **common)
    t2 = time.perf_counter()
    W.fitBeta(ySEXP=c, xSEXP=x, nfSEXP=w["nf"], alpha_hatSEXP=w["dispersion"], contrastSEXP=np.r_[1.0, 0.0], beta_matSEXP=w["beta0"],
              lambdaSEXP=w["lam"], weightsSEXP=None, useWeightsSEXP=False, tolSEXP=1e-8, maxitSEXP=100, useQRSEXP=True, minmuSEXP=0.5)
    t3 = time.perf_counter()
print("threads", os.environ.get("B200NB_STAGE_THREADS", "default"), "ms: fitDisp %.2f fitDisp %.2f fitBeta %.2f total %.2f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
