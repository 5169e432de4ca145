#!/bin/bash
# Round-2 host-path sweep on the GPU box: scripts/e2e_probe2.py once per knob setting (the knobs are read once per process).
cd "$(dirname "$0")/.."
run() { env "$@" timeout 300 python scripts/e2e_probe2.py 2>&1 | tail -1; }
cat /sys/kernel/mm/transparent_hugepage/enabled; nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"
run B200NB_X=0
B200NB_HOST_TIMING=1 timeout 300 python scripts/e2e_probe2.py 2>&1 | grep "b200nb timing" | tail -6
run B200NB_SPECULATE=0
for t in 4 12 16; do run B200NB_HOST_THREADS=$t; done
run B200NB_NUMA_BIND=0
run B200NB_CACHE_MB=0 B200NB_SF_DETECT=0
