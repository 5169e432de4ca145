#!/bin/bash
# One full ncu capture of one kernel of the bench step (recipe of /opt/skills/guides/B200_PROFILING.md), to be run on
# the GPU box under gpurun with ONE GPU; writes gpurun_out/<tag>.ncu-rep plus the raw / source CSV exports that
# profiles/src_hot.py and profiles/sass_hist.py read.
#   scripts/ncu_capture.sh <tag> <kernel regex> [library]      e.g.  scripts/ncu_capture.sh r02a_fit_disp fit_disp_kernel
#   NCU_CMD='python scripts/c4_ab.py 8000' NCU_SKIP=0 scripts/ncu_capture.sh r02h_generic_disp fit_disp_generic_kernel
#   scripts/ncu_capture.sh r02a_half fit_disp_grp_kernel deseq2_b200/libb200nb_exp_half_warp.so
set -e
cd "$(dirname "$0")/.."
tag=$1; kern=$2; lib=${3:-}; skip=${NCU_SKIP:-8}   # skip the (row-chunked) launches of the workload build: land in the device-resident steps
mkdir -p gpurun_out
[ -n "$lib" ] && export B200NB_LIB="$PWD/$lib"
ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f -o "gpurun_out/$tag" \
    ${NCU_CMD:-python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs} > "gpurun_out/$tag.log" 2>&1
ncu -i "gpurun_out/$tag.ncu-rep" --page raw --csv > "gpurun_out/${tag}_raw.csv"
ncu -i "gpurun_out/$tag.ncu-rep" --page source --csv --print-source cuda,sass > "gpurun_out/${tag}_src.csv" || true
python profiles/src_hot.py "gpurun_out/${tag}_src.csv" > "gpurun_out/${tag}_hot_lines_all.txt"; head -40 "gpurun_out/${tag}_hot_lines_all.txt" > "gpurun_out/${tag}_hot_lines.txt" || true
grep -E "gpu__time_duration.sum|sm__pipe_fp64_cycles_active.avg.pct|smsp__issue_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread|smsp__inst_executed.sum|dram__bytes_(read|write).sum" \
    "gpurun_out/${tag}_raw.csv" | head -20 || true
python profiles/ncu_summary.py "gpurun_out/${tag}_raw.csv" > "gpurun_out/${tag}_ncu_summary.txt" 2>&1 || true
python profiles/sass_hist.py "gpurun_out/${tag}_src.csv" > "gpurun_out/${tag}_sass_hist.txt" 2>&1 || true
# gpurun copies at most 64 MiB back: NCU_KEEP=0 drops the report and the source export once the summaries exist
# (NCU_KEEP=src keeps the per-instruction source export, gzipped, but not the report)
if [ "${NCU_KEEP:-1}" = "0" ]; then rm -f "gpurun_out/$tag.ncu-rep" "gpurun_out/${tag}_src.csv" "gpurun_out/${tag}_hot_lines_all.txt"; fi
if [ "${NCU_KEEP:-1}" = "src" ]; then rm -f "gpurun_out/$tag.ncu-rep"; gzip -f "gpurun_out/${tag}_src.csv"; fi
