"""Opcode histogram of every kernel in libhdgpu.so (cuobjdump -sass): instruction count, the FP64/FP32 FMA count, the memory
and asynchronous-copy mnemonics that prove what the kernel uses (UTMALDG = TMA tensor load, UBLKCP/UBLKPF = bulk copy /
prefetch, SYNCS = mbarrier, LDG.E.ENL2.256 = 32-byte global loads, USETMAXREG = setmaxnreg), and local-memory traffic (LDL/STL,
0 = no spills and no dynamically indexed register arrays).   python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hyperdeal_b200", "lib", "libhdgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, hist = None, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
def shorten(name):
    name = name.replace("(anonymous namespace)::", "")
    if name.startswith("void "):
        name = name[5:]
    depth = 0
    for i, ch in enumerate(name):  # cut the parameter list: the first "(" outside template brackets
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            return name[:i]
    return name


KEYS = ["DFMA", "DADD", "DMUL", "FFMA", "LDG", "STG", "LDS", "STS", "LDL", "STL", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTMAPF", "SYNCS", "USETMAXREG", "BAR", "RED", "ATOM", "MUFU"]
print("library: %s" % os.path.relpath(lib, ROOT))
print("%-88s %7s " % ("kernel", "instr") + " ".join("%6s" % k[:6] for k in KEYS))
rows = []
for k, h in hist.items():
    fam = collections.Counter()
    for op, c in h.items():
        fam[op.split(".")[0]] += c
    rows.append((demangle(k), sum(h.values()), fam, h))
rows.sort(key=lambda r: r[0])
for name, total, fam, h in rows:
    short = shorten(name)
    print("%-88s %7d " % (short[:88], total) + " ".join("%6d" % fam.get(k, 0) for k in KEYS))
print()
print("wide / special variants per kernel (only kernels that have them):")
for name, total, fam, h in rows:
    special = {op: c for op, c in h.items() if re.search(r"ENL2|\.256|\.128|UTMA|UBLK|SYNCS|USETMAXREG|MULTICAST", op)}
    if special:
        print("  %s" % shorten(name)[:110])
        print("      " + ", ".join("%s x%d" % kv for kv in sorted(special.items())))
