"""Summarise an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`) per
source line: warp instructions executed and stall samples, biggest first."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2] if rows[1][0] == "Function Name" else rows[1]
start = 3 if rows[1][0] == "Function Name" else 2
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = []
tot_i = tot_s = 0
for r in rows[start:]:
    if len(r) < 10 or r[0] in ("", "Line No"):
        continue
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    stall = {h[6:]: int(r[i]) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i]) > 0}
    lines.append((n, s, r[0], r[1][:100], stall))
    tot_i += n; tot_s += s
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("--- by instructions"); 
for n, s, l, src, st in sorted(lines, reverse=True)[:top]:
    print("%10d (%4.1f%%) smp %6d  L%-4s %s" % (n, 100.0 * n / tot_i, s, l, src))
print("--- by stall samples")
for n, s, l, src, st in sorted(lines, key=lambda x: -x[1])[:top]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%6d (%4.1f%%) inst %10d  L%-4s %-70s %s" % (s, 100.0 * s / tot_s, n, l, src[:70], tops))
