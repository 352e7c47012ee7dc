#!/usr/bin/env python
"""Per-source-line stall summary from an ncu report captured with -lineinfo / --import-source on:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [N]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.OrderedDict()
fname, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ci = hdr.index("# Samples")
        ie = hdr.index("Instructions Executed")
        stall = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= ci:
        continue
    line, src = r[0], r[1]
    key = (fname, line)
    if line:
        cur = agg.setdefault(key, {"src": src.strip(), "samples": 0, "inst": 0, "stall": collections.Counter()})
        last = key
    else:
        cur = agg[last]
    try:
        cur["samples"] += int(r[ci] or 0)
        cur["inst"] += int(r[ie] or 0)
        for j, h in stall:
            cur["stall"][h] += int(r[j] or 0)
    except ValueError:
        pass
tot = sum(v["samples"] for v in agg.values())
toti = sum(v["inst"] for v in agg.values())
print("total samples", tot, "total warp-instructions", toti)
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    st = ", ".join(f"{h[6:]}={c}" for h, c in v["stall"].most_common(3))
    print(f"{f}:{l:>4} {100*v['samples']/max(tot,1):5.1f}% inst {100*v['inst']/max(toti,1):5.1f}%  {v['src'][:70]:70s} [{st}]")
