"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: python profiles/src_hot.py <csv> [kernel_index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tables = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fp = rows[i][1]; hdr = rows[i + 2]; j = i + 3; body = []
        while j < len(rows) and rows[j] and rows[j][0] != "File Path":
            body.append(rows[j]); j += 1
        tables.append((fp, hdr, body)); i = j
    else:
        i += 1
# tables repeat per profiled kernel instance; group by first file occurrence
first = tables[0][0]
starts = [k for k, t in enumerate(tables) if t[0] == first]
lo = starts[which]; hi = starts[which + 1] if which + 1 < len(starts) else len(tables)
agg = collections.Counter(); stall = collections.Counter(); src = {}; total = 0; nsass = 0
for fp, hdr, body in tables[lo:hi]:
    ie = hdr.index("Instructions Executed"); ss = hdr.index("Warp Stall Sampling (All Samples)")
    cur = None
    ai = hdr.index("Address")
    for r in body:
        if r[0]:
            cur = (fp.split("/")[-1], r[0]); src[cur] = r[1].strip()
            continue            # the source-line row itself only repeats the sum of its SASS rows
        if not r[ai]:
            continue
        try:
            n = int(r[ie] or 0); s = int(r[ss] or 0)
        except ValueError:
            continue
        key = cur; agg[key] += n; stall[key] += s; total += n; nsass += 1
print("SASS instructions:", nsass, " warp-instructions executed:", total, " stall samples:", sum(stall.values()))
print("%-16s %6s %7s %7s  %s" % ("file", "line", "inst%", "stall%", "source"))
for key, n in agg.most_common(45):
    print("%-16s %6s %6.2f%% %6.2f%%  %s" % (key[0][:16], key[1], 100.0 * n / total, 100.0 * stall[key] / max(1, sum(stall.values())), src.get(key, "")[:110]))
