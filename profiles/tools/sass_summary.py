"""SASS mnemonic counts per kernel of genes2genes_b200/libg2g.so (the evidence table of B200_PROFILING.md:
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA, FFMA2 = packed f32 FMA).

    python profiles/tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "genes2genes_b200", "libg2g.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DADD",
        "F2FP", "HADD2", "FHADD", "SHFL", "USETMAXREG", "BAR", "ATOMS", "REDG", "LDS", "STS", "LDG", "STG", "LDL", "STL"]
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\((anonymous namespace::)?\w*Params\)|\(.*\)$", "", name)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        counts[name][m.group(1)] += 1
print("kernel".ljust(58) + " ".join(w.rjust(7) for w in want) + "   total")
for k, c in counts.items():
    tot = sum(c.values())
    if tot < 50:
        continue
    print(k[:57].ljust(58) + " ".join(str(c.get(w, 0)).rjust(7) for w in want) + "  %6d" % tot)
