"""Turn an `ncu -i X.ncu-rep --page raw --csv` export into the short summary kept under profiles/ (the metrics the
round verdicts quote + the warp-stall breakdown).  usage: python profiles/ncu_summary.py <raw.csv> "<title line>" """
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[-1]
col = {h: i for i, h in enumerate(hdr)}
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("kernel =", vals[col["Kernel Name"]], "| grid", vals[col["Grid Size"]], "block", vals[col["Block Size"]])
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dmul_pred_on.sum"]
for w in WANT:
    if w in col:
        print(f"{w} [{units[col[w]]}] = {vals[col[w]]}")
stalls = [(h, float(vals[i])) for h, i in col.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and
          not h.endswith("_not_issued") and vals[i] not in ("", "n/a")]
tot = sum(v for _, v in stalls) or 1.0
for h, v in sorted(stalls, key=lambda t: -t[1])[:9]:
    print(f"stall {100 * v / tot:6.2f}%  {h}")
