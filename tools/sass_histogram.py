"""SASS opcode histogram of every translation unit of libpnol_b200.so (cuobjdump -sass build/*.o): per object and per kernel the
counts of the opcodes that prove which hardware path a kernel takes -- DMMA (FP64 tensor), UTMALDG / UBLKCP (TMA), SYNCS (mbarrier),
LDGSTS (cp.async), DFMA / DADD / DMUL / MUFU (FP64 ALU), SHFL, ATOM / RED, BAR, and the totals. Run on the build container (no GPU):
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "UCGABAR", "USETMAXREG", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "SHFL", "MATCH", "ATOM", "ATOMS", "RED",
         "BAR", "LDS", "STS", "LDG", "STG", "LDL", "STL"]


def main():
    for obj in sorted(glob.glob(os.path.join(ROOT, "build", "*.o"))):
        if os.path.basename(obj).startswith("host_"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
                cur = kernels.setdefault(name[:100], collections.Counter())
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and cur is not None:
                cur[m.group(1)] += 1
                cur["_total"] += 1
        tot = collections.Counter()
        for c in kernels.values():
            tot.update(c)
        print("== %s: %d kernels, %d instructions" % (os.path.basename(obj), len(kernels), tot["_total"]))
        print("   " + "  ".join("%s=%d" % (op, tot[op]) for op in WATCH if tot[op]))
        for name, c in kernels.items():
            sel = "  ".join("%s=%d" % (op, c[op]) for op in WATCH if c[op])
            print("   %-100s total=%-6d %s" % (name, c["_total"], sel))
        print()


if __name__ == "__main__":
    main()
