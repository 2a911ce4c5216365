#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the shipped library (cuobjdump -sass p64_b200/libp64b200.so): the mnemonics that prove the
design claims (UTMALDG = TMA tile loads, VABSDIFF4 = packed-byte SAD, IDP.4A = byte dot products, REDUX/VOTE = warp reductions,
ATOMG/ATOMS, LDL/STL = spills) plus the static instruction count and code size of every kernel.
    python tools/sass_counts.py > profiles/rNN_sass_counts.txt"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "p64_b200", "libp64b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTMALDG", "VABSDIFF4", "IDP", "IMAD", "SHF", "PRMT", "LDS", "STS", "LDG", "STG", "REDUX", "VOTE", "SHFL", "ATOMG", "ATOMS", "BAR", "SYNCS", "LDL", "STL"]
print(f"# {os.path.relpath(lib, ROOT)}: arch " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", txt)))))
print(f"{'kernel':58s} {'instr':>6s} {'KiB':>6s}  " + " ".join(f"{k:>9s}" for k in KEYS))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    ops = Counter()
    n = 0
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(1)] += 1
            n += 1
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem).replace("p64b::", "").replace("void ", "")
    print(f"{short[:58]:58s} {n:6d} {n * 16 / 1024:6.1f}  " + " ".join(f"{ops[k]:9d}" for k in KEYS))
