#!/usr/bin/env python
"""Code footprint of a kernel from an .ncu-rep (SASS page): static instruction count, and how many instructions carry
which share of the executed warp instructions -- i.e. the size of the hot code that has to stay in the instruction caches.
   python tools/ncu_codesize.py rep"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
ins = []
for r in rows:
    if len(r) > 6 and r[0].startswith("0x"):
        try:
            ins.append((int(r[0], 16), int(r[5] or 0), int(r[2] or 0), r[1].strip()))
        except ValueError:
            pass
base = ins[0][0]
tot = sum(i[1] for i in ins)
print(f"static instructions {len(ins)} ({len(ins) * 16 / 1024:.0f} KiB), executed warp instructions {tot}")
srt = sorted(ins, key=lambda i: -i[1])
acc = 0
marks = [0.5, 0.8, 0.9, 0.95, 0.99]
mi = 0
for n, i in enumerate(srt, 1):
    acc += i[1]
    while mi < len(marks) and acc >= marks[mi] * tot:
        print(f"  {marks[mi]:.0%} of the executed instructions come from {n} static instructions ({n * 16 / 1024:.1f} KiB)")
        mi += 1
# contiguous hot regions (executed count > 2% of the maximum)
thr = 0.02 * srt[0][1]
regions, cur = [], None
for a, c, s, t in ins:
    if c > thr:
        if cur and a - cur[1] <= 16 * 8:
            cur[1] = a
            cur[2] += c
        else:
            cur = [a, a, c]
            regions.append(cur)
print("hot regions (offset KiB, size KiB, share of executed):")
for r0, r1, c in regions:
    if c > 0.01 * tot:
        print(f"  +{(r0 - base) / 1024:7.1f}  {(r1 - r0 + 16) / 1024:6.1f}  {c / tot:.3f}")
