#!/bin/bash
# 1-GPU call: segmented kernels after the 1024-entry 16-bit factor table, the single divergent region before the
# reductions and the log y! table: config-4 timing, parity subset, ncu of both kernels (per-instruction export kept)
cd "$(dirname "$0")/.."; O=gpurun_out/r02_r; mkdir -p $O
C4_SEG_ONLY=1 timeout 300 python scripts/c4_seg_ab.py 20000 > $O/c4_seg_ab.txt 2> $O/c4_seg_ab.err; cat $O/c4_seg_ab.txt; tail -2 $O/c4_seg_ab.err
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "segmented or general_p or config_shapes or beta_prior or small_p_designs or edge_shapes" 2>&1 | tail -4) > $O/pytest_generic.txt 2>&1; cat $O/pytest_generic.txt
NCU_KEEP=src NCU_CMD='python scripts/c4_seg_ab.py 6000' NCU_SKIP=0 timeout 600 scripts/ncu_capture.sh r02k_seg_disp fit_disp_seg_kernel > $O/ncu_sdisp.txt 2>&1; tail -8 $O/ncu_sdisp.txt | cut -c1-160
NCU_KEEP=src NCU_CMD='python scripts/c4_seg_ab.py 6000' NCU_SKIP=0 timeout 600 scripts/ncu_capture.sh r02k_seg_beta fit_beta_seg_kernel > $O/ncu_sbeta.txt 2>&1; tail -8 $O/ncu_sbeta.txt | cut -c1-160
