#!/bin/bash
# 1-GPU call: register-resident Cox-Reid term for p <= 4 on long rows: C3 shape + config-4 shape (regression check) + tests
cd "$(dirname "$0")/.."; O=gpurun_out/r02_x; mkdir -p $O
timeout 200 python scripts/shape_ab.py C3 2>> $O/err.txt | tail -1 | cut -c1-140 > $O/ab.txt
B200NB_LONG_ROWS=0 timeout 200 python scripts/shape_ab.py C3 2>> $O/err.txt | tail -1 | cut -c1-140 >> $O/ab.txt
C4_SEG_ONLY=1 timeout 300 python scripts/c4_seg_ab.py 20000 2>> $O/err.txt | grep "seg=" >> $O/ab.txt; cat $O/ab.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_device_pipeline_gpu.py -m gpu -x -q -k "long_rows or device_pipeline or segmented or general_p or config_shapes" 2>&1 | tail -3) > $O/pytest.txt 2>&1; cat $O/pytest.txt
