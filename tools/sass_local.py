#!/usr/bin/env python
"""Local-memory instructions (LDL / STL: register spills and demoted arrays) per kernel of an object file or library:
   python tools/sass_local.py jqmc_b200/lib/obj/qe_walker.o [name filter]
The fused walker kernel sits at the 128-register cap; builds whose hot sweep spills show 700+ LDL and run 4-6 % slower
(profiles/r02_walker_history.md), so this is checked before a build is benchmarked."""
import collections, re, subprocess, sys

obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        cnt[name] = [0, 0, 0]
        continue
    if name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", line)
    if m:
        op = m.group(1)
        cnt[name][0] += 1
        if op.startswith("LDL"):
            cnt[name][1] += 1
        elif op.startswith("STL"):
            cnt[name][2] += 1
for n, (tot, ldl, stl) in cnt.items():
    if flt in n and tot > 2000:
        short = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0]
        print(f"{tot:7d} instr  LDL {ldl:5d}  STL {stl:5d}  {short}")
