"""SASS opcode summary per kernel of the shipped library (no GPU needed):
python tools/sass_summary.py [medgp_b200/libmedgp_cuda.so] > profiles/r02_sass_opcodes.txt
Shows, per kernel, the instruction count and the opcodes that identify the hardware path: DMMA
(FP64 tensor core), UBLKCP (cp.async.bulk, TMA engine), SYNCS (mbarrier), LDGSTS (cp.async), DFMA/DMUL/DADD,
MUFU.RSQ64H, SHFL, ATOM/RED, and the absence of UTMALDG / UTCMMA / LDTM (tcgen05 has no FP64 kind)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "medgp_b200", "libmedgp_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, ops = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        ops[kern][m.group(1).split(".")[0] if not m.group(1).startswith("MUFU") else m.group(1)] += 1
KEY = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "SHFL", "ATOMS", "ATOMG", "RED", "LDS", "STS",
       "LDG", "STG", "BAR", "UTMALDG", "UTCMMA", "LDTM"]
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.strip().splitlines()
print("library:", os.path.relpath(lib, ROOT), "| embedded ELF:", ", ".join(a.split()[-1] for a in arch))
print(f"{'kernel':34s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEY))
tot = collections.Counter()
for k, c in ops.items():
    print(f"{k[:34]:34s} {sum(c.values()):7d} " + " ".join(f"{c.get(x, 0):7d}" for x in KEY))
    tot.update(c)
print(f"{'TOTAL':34s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(x, 0):7d}" for x in KEY))
