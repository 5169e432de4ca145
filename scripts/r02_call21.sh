#!/bin/bash
# 1-GPU call: reciprocal-diagonal Cholesky in the segmented kernels, smaller Cook's scratch: timings + parity subset
cd "$(dirname "$0")/.."; O=gpurun_out/r02_u; mkdir -p $O
C4_SEG_ONLY=1 timeout 300 python scripts/c4_seg_ab.py 20000 > $O/c4_seg_ab.txt 2> $O/c4_seg_ab.err; cat $O/c4_seg_ab.txt; tail -2 $O/c4_seg_ab.err
timeout 250 python scripts/c4_prep_cooks_ab.py 20000 > $O/prep_cooks.txt 2>&1; tail -9 $O/prep_cooks.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_device_pipeline_gpu.py -m gpu -x -q -k "segmented or general_p or config_shapes or beta_prior or cooks or prep or device_pipeline" 2>&1 | tail -4) > $O/pytest.txt 2>&1; cat $O/pytest.txt
