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
    for key in ("lk_fast_kernel", "lk_kernel", "pyr_build_fused_kernel", "pyr_down_ring_kernel", "pyr_down_tma_kernel", "pyr_down_kernel", "repitch_kernel",
                "track_filter_kernel", "cov_kernel", "row_scan_kernel", "row_sum_small_kernel", "col_scan_kernel", "candidates_kernel",
                "rank_keys_kernel", "scatter_keys_kernel"):
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
        f.write("(-s 466 -c 400: the pool-setup copies are skipped; per-launch times are cold-cache and serialised -> shares only)\n")
        f.write("%d launches captured, %.1f us total\n\n" % (len(rows), tot))
        f.write("%-60s %8s %12s %10s %7s  %s\n" % ("kernel", "launches", "total us", "avg us", "share", "grid x block (first)"))
        for k, (n, us, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.1f %10.2f %6.1f%%  %s x %s\n" % (k, n, us, us / n, 100 * us / tot, g, b))
        f.write("\nfirst 40 launches in order:\n")
        for k, g, b, us in rows[:40]:
            f.write("  %-60s %-14s %-12s %9.2f us\n" % (k, g, b, us))

corners = os.path.join(src, "corners_launches.csv")
if os.path.exists(corners):
    with open(corners) as f:
        text = f.read()
    start = text.find('"ID"')
    rows = []
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
    with open(os.path.join(dst, "corners_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none over `python scripts/corners_time.py` (goodFeaturesToTrack on a\n")
        f.write("KITTI-shape frame with the reference's parameters; per-launch times are cold-cache and serialised -> shares only)\n")
        f.write("%d launches captured, %.1f us total\n\n" % (len(rows), tot))
        f.write("%-40s %8s %12s %10s %7s  %s\n" % ("kernel", "launches", "total us", "avg us", "share", "grid x block (first)"))
        for k, (n, us, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s %8d %12.1f %10.2f %6.1f%%  %s x %s\n" % (k, n, us, us / n, 100 * us / tot, g, b))

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
# machine-readable counters of the captured kernels (bench.py copies `traffic` and the issue / shared-memory utilisation
# from here into its JSON line, labelled as ncu figures)
import json


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [{h.strip(): (v, u) for h, u, v in zip(hdr, units, r)} for r in rows[2:]]


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,      # to bytes
         "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}   # to us


def num(d, key):
    if key not in d:
        return None
    v, u = d[key]
    try:
        return float(v.replace(",", "")) * SCALE.get(u.strip(), 1.0)
    except ValueError:
        return None


counters = {}
for rep in sorted(os.listdir(src)):
    if not rep.endswith(".ncu-rep"):
        continue
    for d in raw_rows(os.path.join(src, rep)):
        name = short(d.get("Kernel Name", ("", ""))[0])
        rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
        entry = {
            "report": rep, "kernel": name, "grid": d.get("launch__grid_size", ("", ""))[0].strip(),
            "duration_us_under_ncu": num(d, "gpu__time_duration.sum"),
            "dram_bytes_read": rd, "dram_bytes_write": wr,
            "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "smem_wavefronts_pct_of_peak": num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers_per_thread": num(d, "launch__registers_per_thread"),
            "sm_active_cycles_avg": num(d, "smsp__cycles_active.avg"), "sm_elapsed_cycles_max": num(d, "sm__cycles_elapsed.max"),
        }
        counters.setdefault(rep.replace(".ncu-rep", ""), []).append(entry)
# bench.py quotes these figures only for the kernel sources they were captured on
sys.path.insert(0, os.path.dirname(here))
try:
    import bench
    counters["csrc_sha16"] = bench.csrc_sha16()
except Exception as ex:   # pragma: no cover
    counters["csrc_sha16"] = "unknown (%r)" % ex
with open(os.path.join(dst, "counters.json"), "w") as f:
    json.dump(counters, f, indent=1)
print("wrote", sorted(os.listdir(dst)))
