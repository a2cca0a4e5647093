#!/usr/bin/env python3
"""Turns one tools/gpu_round.sh pass (gpurun_out/) into the tracked summaries under profiles/.

  python tools/summarise_profiles.py <tag>        e.g. r01_s4

Writes profiles/<tag>_bench_c3.json, _bench_ref.json, _pytest_gpu.log, _launches_summary.csv,
_k_intersect_ncu_full.json (every metric of the --set full capture) and _k_intersect_lines.txt
(per source line: share of executed warp instructions, mean active lanes, share of stall samples).
"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "k_intersect"

for src, dst in [("bench_c3.json", "bench_c3.json"), ("bench_ref.json", "bench_ref.json"), ("pytest_gpu.log", "pytest_gpu.log")]:
    if (OUT / src).exists():
        shutil.copy(OUT / src, PROF / f"{tag}_{dst}")

# launch list -> per-kernel totals
if (OUT / "launches.csv").exists():
    rows = [r for r in csv.reader(open(OUT / "launches.csv")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    iK, iV = h.index("Kernel Name"), h.index("Metric Value")
    iU = h.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iV:
            continue
        v = float(r[iV].replace(",", ""))
        v *= {"us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(r[iU], 1.0)
        a = tot.setdefault(r[iK][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in tot.values())
    with open(PROF / f"{tag}_launches_summary.csv", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 700, python bench.py --steps 1 --warmup 1 "
                "--spp-step 4 (c3); ns, cold-cache/serialised\n")
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ns", "share"])
        for k, a in sorted(tot.items(), key=lambda x: -x[1][1]):
            w.writerow([k, a[0], int(a[1]), "%.4f" % (a[1] / total)])

rep = OUT / f"{kernel}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) >= 3:
        names, units, vals = rows[0], rows[1], rows[2]
        d = {n: [u, v] for n, u, v in zip(names, units, vals)}
        json.dump(d, open(PROF / f"{tag}_{kernel}_ncu_full.json", "w"), indent=1)
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    tmp = OUT / f"{kernel}_source.csv"
    tmp.write_text(src)
    lines = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_lines.py"), str(tmp), "40"], capture_output=True, text=True)
    (PROF / f"{tag}_{kernel}_lines.txt").write_text(lines.stdout + lines.stderr[-2000:])
print("wrote", sorted(p.name for p in PROF.glob(tag + "_*")))
