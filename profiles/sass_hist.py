"""Opcode histogram (weighted by executed warp-instructions) of an `ncu --page source --csv --print-source sass` export.
usage: python profiles/sass_hist.py <csv> [kernel_index]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
lo = hdr_idx[which]; hi = hdr_idx[which + 1] if which + 1 < len(hdr_idx) else len(rows)
hdr = rows[lo]
ie = hdr.index("Instructions Executed"); si = hdr.index("Source"); ss = hdr.index("Warp Stall Sampling (All Samples)")
te = hdr.index("Thread Instructions Executed")
ops = collections.Counter(); st = collections.Counter(); total = 0; thr = 0; n = 0
for r in rows[lo + 1:hi]:
    if len(r) <= ie: continue
    try: c = int(r[ie] or 0)
    except ValueError: continue
    ins = r[si].strip()
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    op = ins.split()[0].split(".")[0] if ins else "?"
    ops[op] += c; st[op] += int(r[ss] or 0); total += c; thr += int(r[te] or 0); n += 1
print("static SASS instructions:", n, " executed warp-instructions:", total, " avg active threads: %.1f" % (thr / max(1, total)))
for op, c in ops.most_common(30):
    print("%-10s %6.2f%%  stall %6.2f%%" % (op, 100.0 * c / total, 100.0 * st[op] / max(1, sum(st.values()))))
