#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out/r02_m; mkdir -p $O
timeout 600 python scripts/sync_probe.py 50000 > $O/sync_probe.txt 2>&1; cut -c1-420 $O/sync_probe.txt
echo "== CUDA_DEVICE_MAX_CONNECTIONS / spin flags"; CUDA_DEVICE_SCHEDULE=spin timeout 600 python scripts/c4_diag.py 50000 2>&1 | grep "^run\|^stage\|cProfile run" | cut -c1-400 > $O/c4_diag2.txt; cat $O/c4_diag2.txt
