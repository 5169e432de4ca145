"""Host/GPU timeline of one DESeq_device run on the config-4 shape: every engine call and every synchronising torch call
(nonzero, item, as_tensor, sort) logged with its host wall interval and the GPU interval between two events recorded
around it.  usage: python scripts/c4_timeline.py [genes]"""
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
log = []
T0 = [0.0]


def wrap(owner, name, label=None):
    f = getattr(owner, name)

    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, c0 = time.perf_counter(), time.thread_time()
        e0.record()
        r = f(*a, **k)
        e1.record()
        log.append((label or name, 1e3 * (t0 - T0[0]), 1e3 * (time.perf_counter() - t0), 1e3 * (time.thread_time() - c0), e0, e1))
        return r
    setattr(owner, name, g)


for nm in ("fit_disp", "fit_beta", "nb_loglik", "fit_disp_grid", "beta_optim"):
    wrap(D, nm)
for nm in ("prep", "trend_fit", "cooks", "_median"):
    wrap(DP, nm)
wrap(torch, "nonzero")
wrap(torch, "as_tensor")
wrap(torch.Tensor, "item", "item")
wrap(torch.Tensor, "numel", "numel") if False else None

def cpu_stat():
    out = {}
    for path in ("/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.stat"):
        try:
            for line in open(path):
                k, v = line.split()
                out[k] = int(v)
            break
        except OSError:
            pass
    return out


for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        print(path, open(path).read().strip())
    except OSError:
        pass
try:
    import threadpoolctl
    print("thread pools:", [(q["user_api"], q["internal_api"], q["num_threads"]) for q in threadpoolctl.threadpool_info()])
except Exception as ex:
    print("threadpoolctl:", ex)
print("torch threads", torch.get_num_threads(), "affinity", len(os.sched_getaffinity(0)), "process threads",
      len(os.listdir("/proc/self/task")))
LIMIT = os.environ.get("C4_LIMIT_THREADS") == "1"
if LIMIT:
    import threadpoolctl
    threadpoolctl.threadpool_limits(1)
    torch.set_num_threads(1)
    print("limited BLAS / OpenMP / torch pools to one thread")

for rep in range(6):
    log.clear()
    cs0 = cpu_stat()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T0[0] = time.perf_counter()
    a.record()
    res = DP.DESeq_device(y, x, sf)
    b.record()
    t_ret = 1e3 * (time.perf_counter() - T0[0])
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - T0[0])
    cs1 = cpu_stat()
    print(f"run {rep}: GPU {a.elapsed_time(b):.2f} ms, host returned at {t_ret:.2f} ms, wall {wall:.2f} ms; cgroup throttled "
          f"{cs1.get('nr_throttled', 0) - cs0.get('nr_throttled', 0)} periods, {(cs1.get('throttled_usec', 0) - cs0.get('throttled_usec', 0)) / 1e3:.1f} ms; "
          f"cpu used {(cs1.get('usage_usec', 0) - cs0.get('usage_usec', 0)) / 1e3:.1f} ms; threads {len(os.listdir('/proc/self/task'))}")
    if rep == 5 and not LIMIT:
        print("  %-12s %9s %9s %9s %9s %9s" % ("call", "start", "host ms", "cpu ms", "gpu start", "gpu ms"))
        for nm, ts, th, tc, e0, e1 in log:
            print("  %-12s %9.2f %9.2f %9.2f %9.2f %9.2f" % (nm, ts, th, tc, a.elapsed_time(e0), e0.elapsed_time(e1)))
