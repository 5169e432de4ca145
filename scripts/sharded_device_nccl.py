"""Check on 2+ GPUs (NCCL; driven by tests/test_sharded_nccl_gpu.py): the gene-sharded device-resident pipeline equals the single-GPU run.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        scripts/sharded_device_nccl.py
The same logic runs on CPU in tests/test_sharded_gloo.py (gloo + emulated engine)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deseq2_b200 import device as D, device_pipeline as DP, sharded, synth   # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n, m = 20000, 40
x = synth.design_condition(m)
d = synth.make_example_counts(n, m, x=x, seed=5)
lo, hi = sharded.shard_bounds(n, world, rank)
r = sharded.sharded_DESeq_device(D.to_gene_major(d["counts"][lo:hi], dev), x, d["sizeFactors"])
if rank == 0:
    whole = DP.DESeq_device(D.to_gene_major(d["counts"], dev), x, d["sizeFactors"])
    ok = (torch.equal(r["trendCoefs"], whole["trendCoefs"]) and r["dispPriorVar"] == whole["dispPriorVar"]
          and torch.equal(r["gathered"]["dispersion"], whole["dispersion"])
          and torch.allclose(r["gathered"]["betaMatrix"], whole["betaMatrix"], rtol=0, atol=0, equal_nan=True))
    print("sharded == whole:", ok, "| genes", int(whole["idx"].numel()), "| shard of rank 0:", int(r["idx"].numel()),
          "| per-gene collectives:", int(r["collectives"]) - 1, "(+ one 8-byte all-reduce)")
    assert ok
dist.barrier()
dist.destroy_process_group()
