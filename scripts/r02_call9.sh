#!/bin/bash
# 1-GPU call: GPU suite, host-path sweep (staging ring size), C4 kernels + ncu of the general-p kernels, bench line
cd "$(dirname "$0")/.."; O=gpurun_out/r02_i; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
run() { env "$@" timeout 300 python scripts/e2e_probe2.py 2>&1 | tail -1 | cut -c1-400; }
{ run B200NB_X=0; run B200NB_STAGE_KB=2048 B200NB_STAGE_BLOCK_KB=128; run B200NB_STAGE_KB=1024 B200NB_STAGE_BLOCK_KB=64
  run B200NB_STAGE_KB=4096 B200NB_STAGE_BLOCK_KB=256; run B200NB_HOST_THREADS=8; run B200NB_HOST_THREADS=8 B200NB_STAGE_KB=2048 B200NB_STAGE_BLOCK_KB=128
  B200NB_HOST_TIMING=1 timeout 300 python scripts/e2e_probe2.py 2>&1 | grep "b200nb timing" | tail -3; } > $O/e2e_sweep.txt 2>&1
cat $O/e2e_sweep.txt
timeout 600 python scripts/c4_ab.py 20000 > $O/c4_ab.txt 2>&1; tail -4 $O/c4_ab.txt
NCU_CMD='python scripts/c4_ab.py 6000' NCU_SKIP=0 scripts/ncu_capture.sh r02i_generic_disp fit_disp_generic_kernel > $O/ncu_gdisp.txt 2>&1; tail -8 $O/ncu_gdisp.txt | cut -c1-200
NCU_CMD='python scripts/c4_ab.py 6000' NCU_SKIP=0 scripts/ncu_capture.sh r02i_generic_beta fit_beta_generic_kernel > $O/ncu_gbeta.txt 2>&1; tail -8 $O/ncu_gbeta.txt | cut -c1-200
(time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err) 2> $O/bench_time.txt; cat $O/bench_time.txt
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][0])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "kernels", d["roofline"]["kernel_ms"])
print(json.dumps(d["configs"], indent=1)[:2600])
PY
