"""Host-path probe (round 2): the three host-buffer C-ABI calls of one C2 step, timed per call
  (a) through deseq2_b200.wrappers (fresh numpy result arrays each call, like R's allocVector),
  (b) through the raw ctypes symbols with PREALLOCATED, already touched result arrays (what is left is the library),
so that Python / allocation overhead and library time can be told apart.  Every step starts with b200nb_cache_clear()
(a new DESeq() run shares nothing with the previous one).  Knobs come from the environment (one process per setting:
scripts/e2e_sweep2.sh).  Prints one line."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import deseq2_b200
from deseq2_b200 import wrappers as W

n, m = int(os.environ.get("PROBE_GENES", 50000)), int(os.environ.get("PROBE_SAMPLES", 100))
REPS = int(os.environ.get("PROBE_REPS", 9))
NUMA = ""
if os.environ.get("PROBE_WORLD"):   # one of several concurrent single-GPU processes (scripts/contention_probe.sh)
    try:   # the main thread (first touch of the buffers) on the GPU's NUMA node, like bench.bind_to_gpu_numa_node
        node = int(open(f"/sys/bus/pci/devices/{os.environ['PROBE_PCI'].lower()[-12:]}/numa_node").read())
        cpus = set()
        for tok in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = tok.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0))
        NUMA = f"node{node}"
    except Exception as ex:
        NUMA = "unbound(" + repr(ex)[:60] + ")"
w = bench.build_workload(n, m, 20260925, W)
L = deseq2_b200.lib()
c, x, mu = w["counts"], w["x"], w["mu"]
ng = len(c)
P = lambda a: C.c_void_p(a.ctypes.data)
common = dict(ySEXP=c, xSEXP=x, mu_hatSEXP=mu, min_log_alphaSEXP=bench.MIN_LOG_ALPHA, kappa_0SEXP=1.0, tolSEXP=1e-6,
              maxitSEXP=100, weightsSEXP=None, useWeightsSEXP=False, weightThresholdSEXP=1e-2, useCRSEXP=True)


def step_wrappers():
    L.b200nb_cache_clear()
    t0 = time.perf_counter()
    r1 = W.fitDisp(log_alphaSEXP=w["log_alpha0"], log_alpha_prior_meanSEXP=w["log_alpha0"], log_alpha_prior_sigmasqSEXP=1.0,
                   usePriorSEXP=False, **common)
    t1 = time.perf_counter()
    r2 = W.fitDisp(log_alphaSEXP=w["log_dispInit"], log_alpha_prior_meanSEXP=w["log_dispFit"],
                   log_alpha_prior_sigmasqSEXP=w["priorVar"], usePriorSEXP=True, **common)
    t2 = time.perf_counter()
    r3 = W.fitBeta(ySEXP=c, xSEXP=x, nfSEXP=w["nf"], alpha_hatSEXP=w["dispersion"], contrastSEXP=np.r_[1.0, 0.0],
                   beta_matSEXP=w["beta0"], lambdaSEXP=w["lam"], weightsSEXP=None, useWeightsSEXP=False, tolSEXP=1e-8,
                   maxitSEXP=100, useQRSEXP=True, minmuSEXP=0.5)
    t3 = time.perf_counter()
    del r1, r2, r3
    t4 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2, t4 - t3


od = [np.zeros(ng) for _ in range(7)]
oi = [np.zeros(ng, dtype=np.int32) for _ in range(2)]
bo, bv = np.zeros((ng, 2), order="F"), np.zeros((ng, 2), order="F")
it, cn, cd, dv = (np.zeros(ng) for _ in range(4))
H = np.zeros((ng, m), order="F")
contrast = np.r_[1.0, 0.0]


def step_raw():
    L.b200nb_cache_clear()
    t0 = time.perf_counter()
    rc = L.b200nb_fit_disp(P(c), 0, P(x), P(mu), P(w["log_alpha0"]), P(w["log_alpha0"]), 1.0, bench.MIN_LOG_ALPHA, 1.0, 1e-6,
                           100, 0, None, 0, 1e-2, 1, ng, m, 2, P(od[0]), P(oi[0]), P(oi[1]), *[P(a) for a in od[1:]])
    t1 = time.perf_counter()
    rc |= L.b200nb_fit_disp(P(c), 0, P(x), P(mu), P(w["log_dispInit"]), P(w["log_dispFit"]), float(w["priorVar"]),
                            bench.MIN_LOG_ALPHA, 1.0, 1e-6, 100, 1, None, 0, 1e-2, 1, ng, m, 2, P(od[0]), P(oi[0]), P(oi[1]),
                            *[P(a) for a in od[1:]])
    t2 = time.perf_counter()
    rc |= L.b200nb_fit_beta(P(c), 0, P(x), P(w["nf"]), P(w["dispersion"]), P(contrast), P(w["beta0"]), P(w["lam"]), None, 0,
                            1e-8, 100, 1, 0.5, ng, m, 2, P(bo), P(bv), P(it), P(H), P(cn), P(cd), P(dv), None)
    t3 = time.perf_counter()
    assert rc == 0
    return t1 - t0, t2 - t1, t3 - t2, 0.0


def barrier(tag):
    d, world = os.environ.get("PROBE_BARRIER_DIR"), int(os.environ.get("PROBE_WORLD", 0))
    if not d or world < 2:
        return
    open(os.path.join(d, f"{tag}.{os.environ.get('CUDA_VISIBLE_DEVICES', '0')}"), "w").close()
    t0 = time.time()
    while len([f for f in os.listdir(d) if f.startswith(tag + ".")]) < world and time.time() - t0 < 120:
        time.sleep(0.002)


def run(f, reps=REPS):
    for _ in range(3):
        f()
    barrier(f.__name__)
    t = np.array([f() for _ in range(reps)])
    return np.median(t, axis=0) * 1e3


st0 = (C.c_longlong * 6)()
L.b200nb_host_stats(st0, 6)
a = run(step_wrappers)
st1 = (C.c_longlong * 6)()
L.b200nb_host_stats(st1, 6)
b = run(step_raw)
per_step = [(st1[i] - st0[i]) / (REPS + 3) for i in range(6)]
knobs = " ".join(f"{k[7:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("B200NB_") and k != "B200NB_LIB")
print(f"{NUMA + ' ' if NUMA else ''}{knobs or 'default'} | wrappers: disp {a[0]:.2f} disp {a[1]:.2f} beta {a[2]:.2f} free {a[3]:.2f} total {a[:3].sum():.2f} ms"
      f" | raw preallocated: disp {b[0]:.2f} disp {b[1]:.2f} beta {b[2]:.2f} total {b[:3].sum():.2f} ms"
      f" | per step: H2D {per_step[0] / 1e6:.1f} MB, D2H {per_step[1] / 1e6:.1f} MB, cache hits {per_step[2]:.1f}, "
      f"served from cache {per_step[4] / 1e6:.1f} MB, hashed {per_step[5] / 1e6:.1f} MB", flush=True)
