#!/bin/bash
# 1-GPU call: general-p kernels with one evaluation site + unroll 1 vs the unroll-1 build of the old structure; C4 pipeline host profile
cd "$(dirname "$0")/.."; O=gpurun_out/r02_l; mkdir -p $O
for lib in libb200nb.so libb200nb_exp_gen_u1.so; do
  echo "== $lib"; B200NB_LIB=$PWD/deseq2_b200/$lib C4_AB_MODES=smem timeout 300 python scripts/c4_ab.py 20000 2>&1 | tail -1
done > $O/gen_variants.txt 2>&1
cat $O/gen_variants.txt
timeout 600 python scripts/c4_diag.py 50000 > $O/c4_diag.txt 2>&1; cut -c1-200 $O/c4_diag.txt | head -80
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_device_pipeline_gpu.py tests/test_parity_reference_gpu.py -m gpu -x -q 2>&1 | tail -4) > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
NCU_CMD='python scripts/c4_ab.py 6000' NCU_SKIP=0 scripts/ncu_capture.sh r02l_generic_disp fit_disp_generic_kernel > $O/ncu_gdisp.txt 2>&1; tail -8 $O/ncu_gdisp.txt | cut -c1-200
NCU_CMD='python scripts/c4_ab.py 6000' NCU_SKIP=0 scripts/ncu_capture.sh r02l_generic_beta fit_beta_generic_kernel > $O/ncu_gbeta.txt 2>&1; tail -8 $O/ncu_gbeta.txt | cut -c1-200
