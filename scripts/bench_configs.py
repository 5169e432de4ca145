#!/usr/bin/env python
"""Kernel-only timings of the three native calls on the shapes of every BASELINE.json config (per-GPU shard sizes),
device-resident inputs, CUDA events.  Not the bench line (bench.py is): a coverage / roofline table for DESIGN.md.
usage: python scripts/bench_configs.py [--scale 1.0] > profiles/rNN_configs.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deseq2_b200 import device as D, pipeline, synth  # noqa: E402

MINLA = float(np.log(1e-9))


def run(name, n, m, x, x_fit=None, lam=None, reps=3, lrt_reduced=None):
    dev = torch.device("cuda")
    d = synth.make_example_counts(n, m, x=x, seed=11, betaSD=0.5)
    counts = d["counts"]
    counts = counts[counts.sum(1) > 0]
    n = len(counts)
    sf = d["sizeFactors"]
    norm = counts / sf
    bm = norm.mean(1)
    alpha = np.clip(0.1 + 4 / bm, 1e-8, max(10, m))
    Q, R = np.linalg.qr(x)
    beta0 = np.linalg.solve(R, Q.T @ np.log(norm + 0.1).T).T
    mu = np.maximum(np.exp(beta0 @ x.T) * sf, 0.5)
    y = D.to_gene_major(counts, dev)
    mud = D.to_gene_major(mu, dev)
    xd = D.x_to_device(x, dev)
    la0 = torch.as_tensor(np.log(alpha * np.exp(np.random.default_rng(0).normal(0, 0.5, n))), device=dev)
    lfit = torch.as_tensor(np.log(alpha), device=dev)
    p = x.shape[1]
    res = {"config": name, "genes": n, "samples": m, "p": p}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            a, b = ev(), ev()
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    out = {}
    res["fit_disp_mle_ms"] = timeit(lambda: out.update(mle=D.fit_disp(y, xd, mud, la0, la0, 1.0, MINLA, 1.0, 1e-6, 100, False)))
    res["fit_disp_map_ms"] = timeit(lambda: out.update(map=D.fit_disp(y, xd, mud, out["mle"]["log_alpha"], lfit, 0.5, MINLA, 1.0, 1e-6, 100, True)))
    xf = x if x_fit is None else x_fit
    pf = xf.shape[1]
    xfd = D.x_to_device(xf, dev)
    if x_fit is None:
        b0 = torch.as_tensor(np.ascontiguousarray(beta0.T), device=dev)
    else:
        bb = np.zeros((pf, n)); bb[0] = np.log(bm)
        b0 = torch.as_tensor(bb, device=dev)
    lamv = np.full(pf, 1e-6) / np.log(2) ** 2 if lam is None else lam
    disp = torch.exp(out["map"]["log_alpha"]).clamp(1e-8, max(10, m))
    sfd = torch.as_tensor(sf, device=dev)
    res["fit_beta_ms"] = timeit(lambda: out.update(beta=D.fit_beta(y, xfd, sfd, disp, np.r_[1.0, np.zeros(pf - 1)], b0, lamv, 1e-8, 100)))
    res["fit_beta_p"] = pf
    res["mean_iter"] = {"disp_mle": float(out["mle"]["iter"].double().mean()), "disp_map": float(out["map"]["iter"].double().mean()),
                        "beta": float(out["beta"]["iter"].mean())}
    tot = res["fit_disp_mle_ms"] + res["fit_disp_map_ms"] + res["fit_beta_ms"]
    res["genes_per_s_3calls"] = n / (tot * 1e-3)
    bytes_disp = n * (12 * m + 88)
    res["fit_disp_hbm_GBs"] = bytes_disp / (res["fit_disp_mle_ms"] * 1e-3) / 1e9
    res["fit_beta_hbm_GBs"] = n * (20 * m + 24 * pf + 40) / (res["fit_beta_ms"] * 1e-3) / 1e9
    # the whole device-resident analysis on this shape (prep, both dispersion fits, trend, grid refits, Wald fit +
    # statistics + Cook's; wall clock incl. the few host syncs).  LRT configs add the reduced fit.
    try:
        from deseq2_b200 import device_pipeline as DP
        import time
        DP.DESeq_device(y, x, sf)
        torch.cuda.synchronize()
        per = []
        for _ in range(5):          # median of 5: the boxes show occasional ~80 ms stalls unrelated to the workload
            t0 = time.perf_counter()
            out_p = DP.DESeq_device(y, x, sf)
            torch.cuda.synchronize()
            per.append((time.perf_counter() - t0) * 1e3)
        res["device_pipeline_wald_ms"] = float(np.median(per))
        res["device_pipeline_wald_ms_all"] = [round(v, 2) for v in per]
        res["device_pipeline_wald_genes_per_s"] = n / (res["device_pipeline_wald_ms"] * 1e-3)
        if lrt_reduced is not None:
            ynz = y[out_p["idx"]].contiguous()
            DP.nbinomLRT_device(ynz, x, lrt_reduced, sf, out_p["dispersion"])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            DP.nbinomLRT_device(ynz, x, lrt_reduced, sf, out_p["dispersion"])
            torch.cuda.synchronize()
            res["device_lrt_ms"] = (time.perf_counter() - t0) * 1e3
    except Exception as ex:
        res["device_pipeline_error"] = repr(ex)[:200]
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    s = a.scale
    run("C2 50k x 100 ~condition p=2", int(50000 * s), 100, synth.design_condition(100))
    run("C3 shard 25k x 500 ~batch+condition p=4", int(25000 * s), 500, synth.design_batch_condition(500, 3))
    run("C4 50k x 1000 10-level factor p=10 (MLE pass)", int(50000 * s), 1000, synth.design_factor(1000, 10))
    run("C4 50k x 1000 expanded p=11 ridge (MAP pass)", int(50000 * s), 1000, synth.design_factor(1000, 10),
        x_fit=synth.design_factor_expanded(1000, 10), lam=np.r_[1e-6, np.full(10, 1 / 0.7)] / np.log(2) ** 2)
    run("C5 shard 125k x 200 ~batch+condition(2x2) p=3", int(125000 * s), 200, synth.design_batch_condition(200, 2),
        lrt_reduced=synth.design_batch_condition(200, 2)[:, :2])
    run("C5 reduced ~batch p=2", int(125000 * s), 200, synth.design_condition(200))
