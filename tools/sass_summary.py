#!/usr/bin/env python
"""Per-kernel-family SASS opcode summary of libdavf_sm100.so (cuobjdump -sass): which kernels issue tcgen05 MMAs
(UTCHMMA / .2CTA), TMA loads / stores / reductions (UTMALDG / UTMASTG / UTMAREDG), TMEM loads (LDTM) and which are
mma.sync / ldmatrix (HMMA / LDSM) or plain CUDA-core kernels.  Makes the "Blackwell-native" claim reviewable without a
rebuild:  python tools/sass_summary.py > profiles/r2/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deepavfusion_b200", "lib", "libdavf_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "LDSM", "MUFU.EX2", "REDG", "ATOMG", "ACQBULK"]
fam = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("davf::", "")
        family = re.sub(r"<.*", "", name)
        cur = fam.setdefault(family, dict(n=0, instr=0, ops=collections.Counter(), variants=[]))
        cur["n"] += 1
        cur["variants"].append(name)
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cur["instr"] += 1
    for o in OPS:
        if o == "UTCHMMA.2CTA":
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["ops"][o] += 1
        elif op.startswith(o):
            cur["ops"][o] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(fam)} kernel families, {sum(f['n'] for f in fam.values())} kernels")
print(f"{'kernel family':34s} {'inst.':>5s} {'SASS':>8s}  " + " ".join(f"{o:>12s}" for o in OPS))
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["instr"]):
    print(f"{k:34s} {f['n']:5d} {f['instr']:8d}  " + " ".join(f"{f['ops'].get(o, 0):12d}" for o in OPS))
