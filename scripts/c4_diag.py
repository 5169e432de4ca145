"""Where does DESeq_device spend its time on the config-4 shape?  Wraps every engine call of the device pipeline in
CUDA events (no extra synchronisation), prints GPU ms per call and the iteration statistics of the two dispersion fits.
usage: python scripts/c4_diag.py [genes] [samples] [levels]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deseq2_b200 import device as D, device_pipeline as DP, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
lv = int(sys.argv[3]) if len(sys.argv) > 3 else 10
x = synth.design_factor(m, lv)
sf = np.exp(np.random.Generator(np.random.PCG64(20260925)).normal(0.0, 0.25, m))
sf = sf / np.exp(np.mean(np.log(sf)))
d = synth.make_example_counts(n, m, x=x, seed=20260923 + 2 + 17, sizeFactors=sf, betaSD=0.5)
dev = torch.device("cuda")
y = D.to_gene_major(d["counts"], dev)
log = []


def wrap(mod, name):
    f = getattr(mod, name)

    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        r = f(*a, **k)
        e1.record()
        log.append((name, e0, e1, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(mod, name, g)


for nm in ("fit_disp", "fit_beta", "nb_loglik", "fit_disp_grid", "beta_optim"):
    if hasattr(D, nm):
        wrap(D, nm)
for nm in ("prep", "trend_fit", "cooks"):
    if hasattr(DP, nm):
        wrap(DP, nm)

for rep in range(3):
    log.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    res = DP.DESeq_device(y, x, sf)
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"run {rep}: DESeq_device {a.elapsed_time(b):.2f} ms on the GPU, {wall:.2f} ms wall; calls (GPU ms / host ms in the call): "
          + ", ".join(f"{nm} {e0.elapsed_time(e1):.2f}/{h:.2f}" for nm, e0, e1, h in log))
print("iter MLE mean %.1f max %d; MAP mean %.1f max %d; genes %d" % (
    res["dispGeneIter"].double().mean().item(), int(res["dispGeneIter"].max()), res["dispIter"].double().mean().item(),
    int(res["dispIter"].max()), res["idx"].numel()))
os.environ["B200NB_PIPE_DEBUG"] = "1"
print("stage wall ms with a sync after every stage:", DP.DESeq_device(y, x, sf)["stage_ms"])
del os.environ["B200NB_PIPE_DEBUG"]
import cProfile
import gc
import pstats
for rep in range(2):
    st0 = torch.cuda.memory_stats()
    g0 = [s_["collections"] for s_ in gc.get_stats()]
    pr_ = cProfile.Profile()
    t0 = time.perf_counter()
    pr_.enable()
    res = DP.DESeq_device(y, x, sf)
    torch.cuda.synchronize()
    pr_.disable()
    wall = (time.perf_counter() - t0) * 1e3
    st1 = torch.cuda.memory_stats()
    g1 = [s_["collections"] for s_ in gc.get_stats()]
    print(f"cProfile run {rep}: wall {wall:.1f} ms; cudaMalloc calls {st1['num_device_alloc'] - st0['num_device_alloc']}, cudaFree calls "
          f"{st1['num_device_free'] - st0['num_device_free']}, alloc retries {st1['num_alloc_retries'] - st0['num_alloc_retries']}, "
          f"reserved {st1['reserved_bytes.all.current'] / 1e9:.2f} GB, gc collections {[b_ - a_ for a_, b_ in zip(g0, g1)]}")
    pstats.Stats(pr_).sort_stats("tottime").print_stats(14)
if os.environ.get("C4_DIAG_PROFILER", "0") == "1":
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        DP.DESeq_device(y, x, sf)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
