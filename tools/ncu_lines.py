#!/usr/bin/env python
"""Hot CUDA-C source lines of an .ncu-rep (needs -lineinfo and --import-source on):  python tools/ncu_lines.py rep [N]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("File Name", "Function Name") or hdr is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    if r[2] != "-":  # SASS rows carry an address; line rows have '-'
        continue
    wi, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    try:
        agg[(fname, ln)] = (int(r[wi] or 0), int(r[ii] or 0), r[1].strip())
    except ValueError:
        pass
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print(f"samples {ts}, warp instructions {ti}")
byfile = collections.Counter()
for (f, ln), v in agg.items():
    byfile[f] += v[1]
print({f: round(n / ti, 3) for f, n in byfile.most_common()})
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    print(f"{v[1] / ti:6.3f} inst {v[0] / ts:6.3f} smp  {f}:{ln}  {v[2][:110]}")
