#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" timeout 300 python scripts/e2e_probe2.py 2>&1 | tail -1; }
run B200NB_X=0
B200NB_HOST_TIMING=1 timeout 300 python scripts/e2e_probe2.py 2>&1 | grep "b200nb timing" | tail -3
run B200NB_CHUNKS=1
run B200NB_CHUNKS=2
run B200NB_CHUNKS=3
run B200NB_CHUNKS=6
run B200NB_CHUNKS=8
