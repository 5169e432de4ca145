import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200NB_PIPE_DEBUG"] = "1"
from deseq2_b200 import device as D, device_pipeline as DP, synth
for name, n, m, x in [("C2", 20000, 100, synth.design_condition(100)), ("C4", 8000, 1000, synth.design_factor(1000, 10)), ("C4", 20000, 1000, synth.design_factor(1000, 10))]:
    d = synth.make_example_counts(n, m, x=x, seed=11, betaSD=0.5)
    y = D.to_gene_major(d["counts"], torch.device("cuda"))
    DP.DESeq_device(y, x, d["sizeFactors"])
    r = DP.DESeq_device(y, x, d["sizeFactors"])
    print(name, n, m, r["stage_ms"], "refits", r["n_refit_geneest"], r["n_refit_map"])
