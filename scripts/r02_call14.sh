#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out/r02_n; mkdir -p $O
timeout 600 python scripts/c4_timeline.py 50000 > $O/timeline.txt 2>&1; cut -c1-260 $O/timeline.txt | grep -v "^  "
C4_LIMIT_THREADS=1 timeout 600 python scripts/c4_timeline.py 50000 > $O/timeline_1thread.txt 2>&1; cut -c1-260 $O/timeline_1thread.txt | grep -v "^  "
OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1 timeout 600 python scripts/c4_timeline.py 50000 > $O/timeline_env1.txt 2>&1; cut -c1-260 $O/timeline_env1.txt | grep -v "^  "
