"""Opcode histogram of every kernel in libdcbf.so (cuobjdump -sass), the Blackwell-specific ones first.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt

The .so is git-ignored, so this text file is the repository's own record that the tcgen05 / TMEM / TMA instructions are
in the shipped kernels: UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UBLKPF = TMA tensor load / store /
bulk prefetch, UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg, SYNCS = mbarrier.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dpdk_dc_sand_b200", "lib", "libdcbf.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTCBAR", "UTCATOMSWS",
       "USETMAXREG", "SYNCS", "ACQBULK", "HMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "STL", "LDL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = cur.replace("(anonymous namespace)::", "").replace("(bool)", "")
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (static instruction counts)")
    for name, ops in kernels.items():
        total = sum(ops.values())
        key = ", ".join(f"{k} {ops[k]}" for k in KEY if ops.get(k))
        print(f"{name}\n    {total} instructions; {key}")


if __name__ == "__main__":
    sys.exit(main())
