"""Per-source-line summary of an ncu report's source page (instructions executed, stall samples).
usage: python scripts/ncu_lines.py <report.ncu-rep> [kernel-regex] [top N]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; func = None
agg = collections.OrderedDict()
seen_func = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if len(sys.argv) > 2 and sys.argv[2] not in (func or ""): continue
    d = dict(zip(hdr, r))
    if r[0] != "":   # source line row
        key = (func, cur_file, int(r[0]))
        def num(k):
            try: return int(d.get(k, "0").replace(",", ""))
            except ValueError: return 0
        inst, samp = num("Instructions Executed"), num("# Samples")
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += inst; a[1] += samp
funcs = collections.OrderedDict()
for (f, fl, ln), (i, s, src) in agg.items():
    funcs.setdefault(f, []).append((i, s, fl, ln, src))
for f, lst in funcs.items():
    ti = sum(x[0] for x in lst); ts = sum(x[1] for x in lst)
    print("== %s: %d warp-instructions, %d samples" % (f, ti, ts))
    for i, s, fl, ln, src in sorted(lst, key=lambda x: -x[0])[:top]:
        print("%6.2f%% inst %5.2f%% samp  %s:%d  %s" % (100.0 * i / max(ti, 1), 100.0 * s / max(ts, 1), fl, ln, src))
