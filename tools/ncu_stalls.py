#!/usr/bin/env python
"""Per-file / per-line stall reasons and shared-memory excess wavefronts from an .ncu-rep source page.
   python tools/ncu_stalls.py rep [N]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
fname = "?"
stall_cols = None
tot = collections.Counter()
byline = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        i_exc = hdr.index("L1 Wavefronts Shared Excessive")
        i_smp = hdr.index("Warp Stall Sampling (All Samples)")
        i_ins = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = {}
    for i, c in stall_cols:
        try:
            d[c] = int(r[i] or 0)
        except ValueError:
            d[c] = 0
    try:
        exc = int(r[i_exc] or 0)
    except ValueError:
        exc = 0
    byline[(fname, ln)] = (d, exc, int(r[i_smp] or 0), int(r[i_ins] or 0), r[1].strip())
    for k, v in d.items():
        tot[k] += v
T = sum(tot.values()) or 1
print("stall totals:", {k: round(v / T, 3) for k, v in tot.most_common(8)})
print("-- lines by samples")
for (f, ln), (d, exc, smp, ins, src) in sorted(byline.items(), key=lambda kv: -kv[1][2])[:N]:
    top = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    print(f"{smp / T:6.3f} {f}:{ln} " + " ".join(f"{k[6:]}={v / max(1, smp):.2f}" for k, v in top) + "  | " + src[:80])
print("-- lines by excessive shared wavefronts")
E = sum(v[1] for v in byline.values()) or 1
for (f, ln), (d, exc, smp, ins, src) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{exc / E:6.3f} ({exc}) {f}:{ln} | {src[:90]}")
