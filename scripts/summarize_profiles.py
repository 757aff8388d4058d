"""Turns one gpurun_out/<tag>/ directory (written by scripts/gpu_round.sh) into the tracked summaries under
profiles/<tag>/: bench JSON lines, the ncu launch list reduced to per-kernel totals and shares, and the key counters of
the two `ncu --set full` captures.   usage: python scripts/summarize_profiles.py gpurun_out/r01a profiles/r01"""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

src, dst = sys.argv[1], sys.argv[2]
os.makedirs(dst, exist_ok=True)
here = os.path.dirname(os.path.abspath(__file__))

for f in os.listdir(src):
    if f.endswith(".json") or f in ("smi.txt", "pytest_gpu.log") or (f.endswith(".log") and "time" in f):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))


def short(name):
    for key in ("lk_fast_kernel", "lk_kernel", "pyr_down_ring_kernel", "pyr_down_tma_kernel", "pyr_down_kernel", "repitch_kernel",
                "track_filter_kernel", "min_eig", "corner"):
        if key in name:
            tail = name.split(key, 1)[1]
            targs = tail.split("(", 1)[0]
            return key + targs
    return name.split("(")[0][-90:]


launches = os.path.join(src, "launches.csv")
if os.path.exists(launches):
    rows = []
    with open(launches) as f:
        text = f.read()
    start = text.find('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
            rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], us))
    agg = collections.OrderedDict()
    for k, g, b, us in rows:
        a = agg.setdefault(k, [0, 0.0, g, b])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(dst, "launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 40 --warmup 3 --no-cpu-baseline`\n")
        f.write("(-s 60 -c 400: setup copies skipped; per-launch times are cold-cache and serialised -> shares only)\n")
        f.write("%d launches captured, %.1f us total\n\n" % (len(rows), tot))
        f.write("%-60s %8s %12s %10s %7s  %s\n" % ("kernel", "launches", "total us", "avg us", "share", "grid x block (first)"))
        for k, (n, us, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.1f %10.2f %6.1f%%  %s x %s\n" % (k, n, us, us / n, 100 * us / tot, g, b))
        f.write("\nfirst 40 launches in order:\n")
        for k, g, b, us in rows[:40]:
            f.write("  %-60s %-14s %-12s %9.2f us\n" % (k, g, b, us))

for rep in sorted(os.listdir(src)):
    if rep.endswith(".ncu-rep"):
        out = subprocess.run([sys.executable, os.path.join(here, "ncu_key.py"), os.path.join(src, rep)], capture_output=True, text=True).stdout
        with open(os.path.join(dst, rep.replace(".ncu-rep", "_key_metrics.txt")), "w") as f:
            f.write("ncu --set full --clock-control none --import-source on  (%s; replayed passes, not a timing source)\n" % rep)
            f.write(out)
        out = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), os.path.join(src, rep), "", "25"], capture_output=True, text=True).stdout
        with open(os.path.join(dst, rep.replace(".ncu-rep", "_hot_lines.txt")), "w") as f:
            f.write("per-source-line share of executed warp-instructions and of stall samples (top 25 lines per kernel)\n")
            f.write(out)
print("wrote", sorted(os.listdir(dst)))
