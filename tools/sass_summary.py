"""Summarise `ncu --page source --csv --print-source sass` output: instructions by opcode, stall reasons, hottest lines.
usage: python tools/sass_summary.py sass.csv [kernel_index] [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sections, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Address":
        hdr = r
    elif cur is not None and hdr and len(r) >= len(hdr):
        cur["rows"].append(r)
ix = {h: i for i, h in enumerate(hdr)}
sec = sections[want]
print(sec["name"], "sass lines", len(sec["rows"]))
tot, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
thr = collections.Counter()
lines = []
for k, r in enumerate(sec["rows"]):
    toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
    opc = toks[0].split(".")[0] if toks else "?"
    ex = int(r[ix["Instructions Executed"]] or 0)
    tot[opc] += ex
    thr[opc] += int(r[ix["Thread Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    samp[opc] += s
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]] or 0)
    lines.append((s, ex, k, r[ix["Source"]]))
te = sum(tot.values())
print("total warp instructions", te, "samples", sum(samp.values()))
for k, v in tot.most_common(28):
    print("%-10s %12d %5.1f%%  avg thr %4.1f  samples %7d %5.1f%%" % (k, v, 100 * v / te, thr[k] / max(v, 1), samp[k],
                                                              100 * samp[k] / max(1, sum(samp.values()))))
print()
ts = sum(stalls.values())
for k, v in stalls.most_common(12):
    print("%-28s %9d %5.1f%%" % (k, v, 100 * v / ts))
print()
for s, ex, k, src in sorted(lines, reverse=True)[:topn]:
    print("%7d samples %10d exec  #%5d  %s" % (s, ex, k, src))
