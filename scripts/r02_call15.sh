#!/bin/bash
# 1-GPU call: GPU suite, bench line + reference arm, launch list, ncu captures of the final kernels
cd "$(dirname "$0")/.."; O=gpurun_out/r02_o; mkdir -p $O
cat /sys/fs/cgroup/cpu.max > $O/cpu_max.txt 2>&1; nproc >> $O/cpu_max.txt
(time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err) 2> $O/bench_time.txt; cat $O/bench_time.txt
(time timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err) 2>> $O/bench_time.txt
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][0])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "kernels", d["roofline"]["kernel_ms"])
print({k: (v.get("deseq_device_ms"), v.get("stage_ms")) for k, v in d["configs"].items()})
r=json.loads([l for l in open("$O/bench_ref.json") if l.startswith("{")][0]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"].get("single_thread"))
PY
run() { env "$@" timeout 300 python scripts/e2e_probe2.py 2>&1 | tail -1 | cut -c1-400; }
{ run B200NB_X=0; B200NB_HOST_TIMING=1 timeout 300 python scripts/e2e_probe2.py 2>&1 | grep "b200nb timing" | tail -3; } > $O/e2e_probe.txt 2>&1; cat $O/e2e_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > $O/b_ncu.log 2>&1
NCU_KEEP=0 scripts/ncu_capture.sh r02h_fit_disp fit_disp_grp_kernel > $O/ncu_disp.txt 2>&1; tail -7 $O/ncu_disp.txt | cut -c1-160
NCU_KEEP=0 scripts/ncu_capture.sh r02h_fit_beta fit_beta_grp_kernel > $O/ncu_beta.txt 2>&1; tail -7 $O/ncu_beta.txt | cut -c1-160
NCU_KEEP=0 NCU_CMD='python scripts/c4_ab.py 6000' NCU_SKIP=0 scripts/ncu_capture.sh r02h_generic_disp fit_disp_generic_kernel > $O/ncu_gdisp.txt 2>&1; tail -7 $O/ncu_gdisp.txt | cut -c1-160
NCU_KEEP=0 NCU_CMD='python scripts/c4_ab.py 6000' NCU_SKIP=0 scripts/ncu_capture.sh r02h_generic_beta fit_beta_generic_kernel > $O/ncu_gbeta.txt 2>&1; tail -7 $O/ncu_gbeta.txt | cut -c1-160
du -sh gpurun_out
