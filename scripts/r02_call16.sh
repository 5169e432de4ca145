#!/bin/bash
# 1-GPU call: segmented general-p kernels -- A/B against the kernels of fit_generic.cu on the config-4 shape, the
# general-p parity tests on the GPU, ncu captures of the two segmented kernels
cd "$(dirname "$0")/.."; O=gpurun_out/r02_p; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
(time timeout 600 python scripts/c4_seg_ab.py 20000 > $O/c4_seg_ab.txt 2> $O/c4_seg_ab.err) 2> $O/time_ab.txt; cat $O/c4_seg_ab.txt; tail -3 $O/c4_seg_ab.err
(time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "segmented or general_p or config_shapes or beta_prior or small_p_designs or edge_shapes" 2>&1 | tail -8) > $O/pytest_generic.txt 2>&1; cat $O/pytest_generic.txt
(time timeout 600 python -m pytest tests/test_device_pipeline_gpu.py tests/test_parity_reference_gpu.py -m gpu -x -q 2>&1 | tail -5) > $O/pytest_pipe.txt 2>&1; cat $O/pytest_pipe.txt
NCU_KEEP=0 NCU_CMD='python scripts/c4_seg_ab.py 6000' NCU_SKIP=0 timeout 600 scripts/ncu_capture.sh r02i_seg_disp fit_disp_seg_kernel > $O/ncu_sdisp.txt 2>&1; tail -8 $O/ncu_sdisp.txt | cut -c1-160
NCU_KEEP=0 NCU_CMD='python scripts/c4_seg_ab.py 6000' NCU_SKIP=0 timeout 600 scripts/ncu_capture.sh r02i_seg_beta fit_beta_seg_kernel > $O/ncu_sbeta.txt 2>&1; tail -8 $O/ncu_sbeta.txt | cut -c1-160
du -sh gpurun_out
