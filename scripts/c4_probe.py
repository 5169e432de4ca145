import os, sys, time, gc
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200NB_PIPE_DEBUG"] = "1"
from deseq2_b200 import device as D, device_pipeline as DP, synth
n, m = 20000, 1000
x = synth.design_factor(m, 10)
d = synth.make_example_counts(n, m, x=x, seed=11, betaSD=0.5)
dev = torch.device("cuda")
y = D.to_gene_major(d["counts"], dev)
for rep in range(5):
    if rep == 3:
        gc.disable()
    s0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    r = DP.DESeq_device(y, x, d["sizeFactors"])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    s1 = torch.cuda.memory_stats()
    print(rep, "total %.1f ms" % dt, {k: v for k, v in r["stage_ms"].items() if v > 5},
          "cudaMalloc calls", s1["num_device_alloc"] - s0["num_device_alloc"], "cudaFree calls", s1["num_device_free"] - s0["num_device_free"],
          "reserved GB %.2f" % (s1["reserved_bytes.all.current"] / 1e9), "gc" if gc.isenabled() else "nogc")
    del r
