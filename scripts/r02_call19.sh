#!/bin/bash
# 1-GPU call: config-4 pipeline stages after the per-group pre-step kernel and the rank-counting Cook's kernel; their tests
cd "$(dirname "$0")/.."; O=gpurun_out/r02_s; mkdir -p $O
C4_SEG_ONLY=1 timeout 300 python scripts/c4_seg_ab.py 20000 > $O/c4_seg_ab.txt 2> $O/c4_seg_ab.err; cat $O/c4_seg_ab.txt; tail -2 $O/c4_seg_ab.err
(timeout 600 python -m pytest tests/test_device_pipeline_gpu.py tests/test_size_factors_gpu.py -m gpu -x -q 2>&1 | tail -4) > $O/pytest_pipe.txt 2>&1; cat $O/pytest_pipe.txt
