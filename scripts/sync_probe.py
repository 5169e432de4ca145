"""Config-4 shape: the engine calls of the device pipeline one at a time, each followed by a device synchronisation.
Per call: kernel time by CUDA events, host time inside the call, wall time until the synchronisation returned and the
CPU time this thread burnt meanwhile -- is a late return GPU work nobody sees, or the host not running?
usage: python scripts/sync_probe.py [genes]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deseq2_b200 import device as D, device_pipeline as DP, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
m = 1000
x = synth.design_factor(m, 10)
sf = np.exp(np.random.Generator(np.random.PCG64(20260925)).normal(0.0, 0.25, m))
sf = sf / np.exp(np.mean(np.log(sf)))
d = synth.make_example_counts(n, m, x=x, seed=20260923 + 2 + 17, sizeFactors=sf, betaSD=0.5)
dev = torch.device("cuda")
y = D.to_gene_major(d["counts"], dev)
res = DP.DESeq_device(y, x, sf)
pr = DP.prep(y, x, sf)
la0 = torch.log(pr["alpha0"])
lfit = torch.log(res["dispFit"])
lam = torch.full((10,), 1e-6 / np.log(2) ** 2, dtype=torch.float64, device=dev)
con = torch.zeros(10, dtype=torch.float64, device=dev)
con[0] = 1
print("cpus", len(os.sched_getaffinity(0)), "loadavg", open("/proc/loadavg").read().strip(), flush=True)
calls = {
    "fit_disp_mle": lambda: D.fit_disp(y, pr["xd"], pr["mu_lin"], la0, la0, 1.0, float(np.log(1e-9)), 1.0, 1e-6, 100, False),
    "fit_disp_map": lambda: D.fit_disp(y, pr["xd"], pr["mu_lin"], torch.log(res["dispGeneEst"]), lfit, res["dispPriorVar"],
                                       float(np.log(1e-9)), 1.0, 1e-6, 100, True),
    "fit_beta": lambda: D.fit_beta(y, pr["xd"], pr["sfd"], res["dispersion"], con, pr["beta0"], lam, 1e-8, 100),
    "cooks": lambda: DP.cooks(y, res["mu"], res["H"], x, sf, want_matrix=False),
    "tiny_torch": lambda: torch.log(la0.abs() + 1.0),
}
worst = {}
for rep in range(12):
    row = []
    for name, fn in calls.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, c0 = time.perf_counter(), time.thread_time()
        e0.record()
        r = fn()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2, c2 = time.perf_counter(), time.thread_time()
        ev = e0.elapsed_time(e1)
        row.append(f"{name} ev {ev:.2f} call {1e3 * (t1 - t0):.2f} wall {1e3 * (t2 - t0):.2f} cpu {1e3 * (c2 - c0):.2f}")
        worst[name] = max(worst.get(name, 0.0), 1e3 * (t2 - t0) - ev)
        del r
    print(f"rep {rep}: " + " | ".join(row), flush=True)
print("largest wall - kernel per call [ms]:", {k: round(v, 2) for k, v in worst.items()})
