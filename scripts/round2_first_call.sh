#!/bin/bash
# The first GPU call of round 2, in one piece.  On the build machine first:
#     scripts/ab_experiments.sh build
# then:  gpurun --timeout 1500 -- 'scripts/round2_first_call.sh'
# Everything lands in gpurun_out/r02_first/.  Each stage has its own timeout so that a surprise in one of the paths that
# have never run on hardware (size factors, refit, chunked host path, experiment kernels) cannot eat the call.
cd "$(dirname "$0")/.."
out=gpurun_out/r02_first
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$out/gpu.txt" 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.txt" 2>&1; echo "smoke exit $?" >> "$out/smoke.txt"
timeout 600 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.txt" 2>&1; echo "pytest exit $?" >> "$out/pytest_gpu.txt"
timeout 120 scripts/ab_experiments.sh quick > "$out/ab_quick.txt" 2>&1
timeout 240 scripts/ab_experiments.sh host > "$out/ab_host.txt" 2>&1
timeout 200 python bench.py > "$out/bench_default.json" 2> "$out/bench_default.err"
timeout 600 scripts/ab_experiments.sh run > "$out/ab_run.txt" 2>&1
tail -n 3 "$out"/smoke.txt "$out"/pytest_gpu.txt
cat "$out/ab_quick.txt" "$out/ab_host.txt"
