"""Config-4 shape: the pre-step kernel per design group vs streaming (B200NB_PREP_GROUPED=0) and the Cook's kernel,
CUDA events, best of 5.  usage: python scripts/c4_prep_cooks_ab.py [genes]"""
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
sf = d["sizeFactors"]
ev = lambda: torch.cuda.Event(enable_timing=True)


def t(fn, reps=5):
    out = fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = ev(), ev()
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


for mode in ("0", "1"):
    os.environ["B200NB_PREP_GROUPED"] = mode
    ms, pr = t(lambda: DP.prep(y, x, sf))
    print(f"prep grouped={mode}: {ms:.3f} ms (events around DP.prep incl. its host glue)  ({n} genes x {m})")
    if mode == "0":
        ref = {k: pr[k].clone() for k in ("alpha0", "mu_lin", "beta0", "baseMean")}
    else:
        for k, v in ref.items():
            print(f"    max rel diff {k}: {((pr[k] - v).abs() / v.abs().clamp(min=1e-12)).max().item():.2e}")
os.environ.pop("B200NB_PREP_GROUPED")
mu = pr["mu_lin"]
H = torch.full_like(mu, 0.01)
ms, ck = t(lambda: DP.cooks(y, mu, H, x, sf, want_matrix=False))
print(f"cooks (maxCooks + robust dispersion, no matrix): {ms:.3f} ms")
ms, ck = t(lambda: DP.cooks(y, mu, H, x, sf, want_matrix=True))
print(f"cooks (with the n x m matrix): {ms:.3f} ms")
