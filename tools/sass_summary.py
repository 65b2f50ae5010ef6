"""Opcode histogram per kernel of the shipped library (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that prove
which hardware path each kernel takes.  usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "jatts_b200", "lib", "libjatts_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "IMMA", "MUFU", "LDG", "STG",
        "LDS", "STS", "SHFL", "BAR", "FFMA", "F2FP", "F2F", "ELECT", "UCGABAR")
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("jb::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEYS:
            if op.startswith(k):
                full = op if k in ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "HMMA") else k
                hist[kern][full] += 1
                break
print(f"# SASS opcode summary of {os.path.relpath(lib, ROOT)} (sm_100a), cuobjdump -sass; one line per kernel\n")
for k, h in hist.items():
    if total[k] == 0:
        continue
    tc = sum(v for o, v in h.items() if o.startswith(("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG")))
    print(f"{k}  [{total[k]} instructions]")
    print("    " + ", ".join(f"{o} x{v}" for o, v in sorted(h.items(), key=lambda x: (-x[1], x[0]))))
    if tc == 0 and h.get("HMMA.1688.F32.TF32", 0) == 0:
        print("    (CUDA-core kernel: no tensor-core / TMA instructions)")
    print()
