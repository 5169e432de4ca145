#!/bin/bash
# 1-GPU call: general-p kernels at 128 registers (smem vs global rows), C4 pipeline diagnosis
cd "$(dirname "$0")/.."; O=gpurun_out/r02_j; mkdir -p $O
timeout 600 python scripts/c4_ab.py 20000 > $O/c4_ab.txt 2>&1; tail -4 $O/c4_ab.txt
timeout 600 python scripts/c4_diag.py 50000 > $O/c4_diag.txt 2>&1; cat $O/c4_diag.txt | cut -c1-1500
B200NB_GENERIC_ROWS=global timeout 600 python scripts/c4_diag.py 50000 > $O/c4_diag_global.txt 2>&1; tail -3 $O/c4_diag_global.txt | cut -c1-1500
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_device_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4) > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
