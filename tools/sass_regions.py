"""Group consecutive SASS lines with equal execution counts: shows which loops the instructions are spent in.
usage: python tools/sass_regions.py sass.csv [min_share_percent]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = [r for r in rows if r and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break
        continue
    if len(r) >= len(hdr) and r[0] != "Address":
        data.append(r)
tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
groups = []
for k, r in enumerate(data):
    ex = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    th = float(r[ix["Avg. Threads Executed"]] or 0)
    if groups and groups[-1]["ex"] == ex:
        g = groups[-1]
        g["n"] += 1
        g["s"] += s
        g["th"] += th
        g["end"] = k
    else:
        groups.append({"ex": ex, "n": 1, "s": s, "th": th, "start": k, "end": k, "first": r[ix["Source"]]})
ts = sum(g["s"] for g in groups)
print("total warp instr %d, samples %d" % (tot, ts))
for g in groups:
    share = 100.0 * g["ex"] * g["n"] / tot
    if share >= minshare or 100.0 * g["s"] / ts >= minshare:
        print("#%5d-%5d  exec %9d x %3d lines = %5.1f%% instr  %5.1f%% samples  thr %4.1f  | %s"
              % (g["start"], g["end"], g["ex"], g["n"], share, 100.0 * g["s"] / ts, g["th"] / g["n"], g["first"][:60]))
