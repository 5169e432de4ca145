"""Config-4 shape (m = 1000, 10-level factor, p = 10: general-p kernels): shared-memory rows vs global row scratch
(B200NB_GENERIC_ROWS=smem|global, read per launch), and the iteration counts of the two dispersion fits of the
device pipeline.  usage: python scripts/c4_ab.py [genes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deseq2_b200 import device as D, device_pipeline as DP, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = 1000
x = synth.design_factor(m, 10)
d = synth.make_example_counts(n, m, x=x, seed=11, betaSD=0.5)
dev = torch.device("cuda")
y = D.to_gene_major(d["counts"], dev)
res = DP.DESeq_device(y, x, d["sizeFactors"])
print("pipeline: mean iter fitDisp MLE %.1f, MAP %.1f; iter==100: MLE %d, MAP %d of %d genes; MAP grid refits %d" % (
    res["dispGeneIter"].double().mean().item(), res["dispIter"].double().mean().item(),
    int((res["dispGeneIter"] >= 100).sum()), int((res["dispIter"] >= 100).sum()), res["idx"].numel(), res["n_refit_map"]))
pr = DP.prep(y, x, d["sizeFactors"])
la0 = torch.log(pr["alpha0"])
lfit = torch.log(res["dispFit"])
lam = torch.full((10,), 1e-6 / np.log(2) ** 2, dtype=torch.float64, device=dev)
con = torch.zeros(10, dtype=torch.float64, device=dev)
con[0] = 1
ev = lambda: torch.cuda.Event(enable_timing=True)


def t(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = ev(), ev()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for mode in os.environ.get("C4_AB_MODES", "smem,global").split(","):
    os.environ["B200NB_GENERIC_ROWS"] = mode
    t1 = t(lambda: D.fit_disp(y, pr["xd"], pr["mu_lin"], la0, la0, 1.0, float(np.log(1e-9)), 1.0, 1e-6, 100, False))
    t2 = t(lambda: D.fit_disp(y, pr["xd"], pr["mu_lin"], torch.log(res["dispGeneEst"]), lfit, res["dispPriorVar"],
                              float(np.log(1e-9)), 1.0, 1e-6, 100, True))
    t3 = t(lambda: D.fit_beta(y, pr["xd"], pr["sfd"], res["dispersion"], con, pr["beta0"], lam, 1e-8, 100))
    print(f"rows in {mode:6s}: fitDisp MLE {t1:.2f} ms, fitDisp MAP {t2:.2f} ms, fitBeta {t3:.2f} ms  ({n} genes x {m})")
os.environ.pop("B200NB_GENERIC_ROWS", None)
