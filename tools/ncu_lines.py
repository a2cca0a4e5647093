#!/usr/bin/env python3
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on).

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv
  python tools/ncu_lines.py x.csv [top]
"""
import collections
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Function Name":
        cur["func"] = r[1]
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
agg = collections.OrderedDict()
for s in sections:
    h = s["hdr"]
    iI, iT, iS = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    for r in s["rows"]:
        if not r[0].strip().isdigit() or len(r) <= iT:
            continue
        key = (s["file"].split("/")[-1], int(r[0]), r[1][:90])
        a = agg.setdefault(key, [0, 0, 0])
        try:
            a[0] += int(r[iI]); a[1] += int(r[iT]); a[2] += int(r[iS])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print("total warp instructions", tot, "samples", tots, "mean lanes %.2f" % (sum(a[1] for a in agg.values()) / tot))
for (f, l, src), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top_n]:
    print(f"{f}:{l:4d} inst {a[0]/tot*100:5.1f}% lanes {a[1]/max(a[0],1):5.1f} samp {a[2]/tots*100:5.1f}% | {src}")
