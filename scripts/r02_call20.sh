#!/bin/bash
# 1-GPU call: the whole GPU suite, the bench line + the reference arm, the launch list of one bench command
cd "$(dirname "$0")/.."; O=gpurun_out/r02_t; mkdir -p $O
cat /sys/fs/cgroup/cpu.max > $O/cpu_max.txt 2>&1; nproc >> $O/cpu_max.txt
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/pytest_gpu.txt 2>&1; cat $O/pytest_gpu.txt
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > $O/smoke.txt 2>&1; cat $O/smoke.txt
(time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err) 2> $O/bench_time.txt; cat $O/bench_time.txt; tail -2 $O/bench.err
(time timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err) 2>> $O/bench_time.txt
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][0])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "kernels", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
print({k: (v.get("deseq_device_ms"), v.get("hot_path_ms"), v.get("stage_ms")) for k, v in d["configs"].items()})
r=json.loads([l for l in open("$O/bench_ref.json") if l.startswith("{")][0]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"].get("single_thread"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/b_ncu.log 2>&1; wc -l $O/launches.csv
du -sh gpurun_out
