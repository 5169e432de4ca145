"""Small-p kernels vs the segmented general-p kernels on the other BASELINE shapes: run once plainly and once with
B200NB_FORCE_GENERIC=1 (read once per process) and compare.  usage: python scripts/shape_ab.py C2|C3|C5|<m>:<design> [genes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deseq2_b200 import device as D, device_pipeline as DP, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
shapes = {"C2": (100, synth.design_condition(100), 50000), "C3": (500, synth.design_batch_condition(500, 3), 25000),
          "C5": (200, synth.design_batch_condition(200, 2), 50000), "M300": (300, synth.design_condition(300), 30000),
          "M1000p2": (1000, synth.design_condition(1000), 10000)}
m, x, n = shapes[name]
if len(sys.argv) > 2:
    n = int(sys.argv[2])
p = x.shape[1]
d = synth.make_example_counts(n, m, x=x, seed=21, betaSD=0.5)
dev = torch.device("cuda")
y = D.to_gene_major(d["counts"], dev)
os.environ["B200NB_PIPE_DEBUG"] = "1"
res = DP.DESeq_device(y, x, d["sizeFactors"])
res = DP.DESeq_device(y, x, d["sizeFactors"])
os.environ.pop("B200NB_PIPE_DEBUG")
pr = DP.prep(y, x, d["sizeFactors"])
mu = res["mu"] if res["mu"].shape == y.shape else pr["mu_lin"]
la0 = torch.log(pr["alpha0"])
lfit = torch.log(res["dispFit"])
lam = torch.full((p,), 1e-6 / np.log(2) ** 2, dtype=torch.float64, device=dev)
con = torch.zeros(p, dtype=torch.float64, device=dev)
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


t1 = t(lambda: D.fit_disp(y, pr["xd"], mu, la0, la0, 1.0, float(np.log(1e-9)), 1.0, 1e-6, 100, False))
t2 = t(lambda: D.fit_disp(y, pr["xd"], mu, torch.log(res["dispGeneEst"]), lfit, res["dispPriorVar"], float(np.log(1e-9)),
                          1.0, 1e-6, 100, True))
t3 = t(lambda: D.fit_beta(y, pr["xd"], pr["sfd"], res["dispersion"], con, pr["beta0"], lam, 1e-8, 100))
tag = "general-p kernels forced" if os.environ.get("B200NB_FORCE_GENERIC") else "small-p kernels"
print(f"{name} ({y.shape[0]} x {m}, p = {p}) {tag:26s}: fitDisp MLE {t1:6.3f} ms, MAP {t2:6.3f} ms, fitBeta {t3:6.3f} ms | "
      f"DESeq_device stages {res['stage_ms']}")
