"""Aggregate the stall reasons of an ncu source page over ranges of source lines.
usage: ncu_stalls.py both.csv 255-281 641-720 ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2] if rows[1][0] == "Function Name" else rows[1]
start = 3 if rows[1][0] == "Function Name" else 2
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
stall_cols = [(h[6:], i) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h]
for rng in sys.argv[2:]:
    lo, hi = [int(v) for v in rng.split("-")]
    agg = {}; inst = 0; smp = 0
    for r in rows[start:]:
        if len(r) < 10 or not r[0].isdigit():
            continue
        if lo <= int(r[0]) <= hi:
            if not r[ix["Instructions Executed"]].isdigit():
                continue
            inst += int(r[ix["Instructions Executed"]]); smp += int(r[ix["# Samples"]])
            for name, i in stall_cols:
                if r[i].isdigit():
                    agg[name] = agg.get(name, 0) + int(r[i])
    tops = sorted(agg.items(), key=lambda kv: -kv[1])[:8]
    print("lines %s: inst %d samples %d : %s" % (rng, inst, smp, ", ".join("%s %d" % kv for kv in tops)))
