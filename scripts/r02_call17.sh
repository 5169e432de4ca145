#!/bin/bash
# 1-GPU call: loop-shape variants of the segmented general-p kernels (NB_SEG_LOOP = 0 / 1 / 2) on the config-4 shape,
# the failed shape of call 16 again, ncu of the dispersion kernel with the per-instruction source export kept
cd "$(dirname "$0")/.."; O=gpurun_out/r02_q; mkdir -p $O
for lib in libb200nb_exp_segloop0.so libb200nb.so libb200nb_exp_segloop2.so; do
  echo "== $lib"; B200NB_LIB=$PWD/deseq2_b200/$lib C4_SEG_ONLY=1 timeout 300 python scripts/c4_seg_ab.py 20000 2> $O/ab_$lib.err | grep -v "^DESeq_device"
done > $O/c4_seg_loop_ab.txt 2>&1; cat $O/c4_seg_loop_ab.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "segmented or general_p or config_shapes or beta_prior or small_p_designs or edge_shapes" 2>&1 | tail -4) > $O/pytest_generic.txt 2>&1; cat $O/pytest_generic.txt
NCU_KEEP=src NCU_CMD='python scripts/c4_seg_ab.py 6000' NCU_SKIP=0 timeout 600 scripts/ncu_capture.sh r02j_seg_disp fit_disp_seg_kernel > $O/ncu_sdisp.txt 2>&1; tail -8 $O/ncu_sdisp.txt | cut -c1-160
NCU_KEEP=src NCU_CMD='python scripts/c4_seg_ab.py 6000' NCU_SKIP=0 timeout 600 scripts/ncu_capture.sh r02j_seg_disp_loop0 fit_disp_seg_kernel deseq2_b200/libb200nb_exp_segloop0.so > $O/ncu_sdisp0.txt 2>&1; tail -8 $O/ncu_sdisp0.txt | cut -c1-160
du -sh gpurun_out
