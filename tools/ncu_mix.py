#!/usr/bin/env python
"""Executed-instruction mix and stall samples by opcode from an .ncu-rep source page:  python tools/ncu_mix.py rep [launch_index]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
b = blocks[which]
h = b[0]
si, wi, ii = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
st = {k: h.index(k) for k in h if k.startswith("stall_") and "Not Issued" not in k}
mix, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
nstatic = 0
hot = []
for r in b[1:]:
    if len(r) <= ii:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
    op = m.group(2).split(".")[0] if m else "?"
    n = int(r[ii] or 0)
    mix[op] += n
    samp[op] += int(r[wi] or 0)
    nstatic += 1
    for k, i in st.items():
        stalls[k] += int(r[i] or 0)
ti, ts = sum(mix.values()), sum(samp.values())
print(f"static SASS instructions {nstatic}, executed warp instr {ti}, samples {ts}")
for op, n in mix.most_common(22):
    print(f"{op:10s} inst {n / ti:6.3f}  stall-samples {samp[op] / ts:6.3f}")
print({k: round(v / ts, 3) for k, v in stalls.most_common(9)})
