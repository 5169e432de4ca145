#!/bin/bash
# Round-2 host-path sweep on the GPU box: scripts/e2e_probe2.py once per knob setting (the knobs are read once per process).
cd "$(dirname "$0")/.."
run() { env "$@" timeout 300 python scripts/e2e_probe2.py 2>&1 | tail -1; }
run B200NB_X=0
B200NB_HOST_TIMING=1 timeout 300 python scripts/e2e_probe2.py 2>&1 | grep "b200nb timing" | tail -6
run B200NB_SPECULATE=0
run B200NB_POPULATE=0
run B200NB_POPULATE_THREADS=16
for t in 8 24 32; do run B200NB_HOST_THREADS=$t; done
run B200NB_CACHE_MB=0 B200NB_SF_DETECT=0 B200NB_POPULATE=0
