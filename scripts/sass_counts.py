"""Per-kernel SASS instruction counts of libgpmp_b200.so (cuobjdump -sass): the mnemonics that show which
hardware paths a kernel uses -- DMMA (FP64 tensor pipe), UTMALDG (TMA bulk tensor load), SYNCS (mbarrier),
LDGSTS (cp.async), MUFU (special function unit) -- plus registers per thread from -res-usage.

    python scripts/sass_counts.py > profiles/r02_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpmp_b200", "libgpmp_b200.so")
WATCH = ["DMMA", "DFMA", "DMUL", "DADD", "UTMALDG", "SYNCS", "LDGSTS", "LDS", "STS", "MUFU", "BAR", "SHFL", "F2I", "FRND"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o).replace("void ", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = tuple(int(v) for v in m.groups())
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["_total"] += 1
    names = list(counts)
    pretty = dict(zip(names, demangle(names)))
    print("SASS instruction counts per kernel, libgpmp_b200.so (sm_100a), static code\n")
    print("| kernel | instr | regs | stack | " + " | ".join(WATCH) + " |")
    print("|---|---|---|---|" + "---|" * len(WATCH))
    for n in sorted(names, key=lambda k: pretty[k]):
        c = counts[n]
        r = regs.get(n, (0, 0, 0))
        print(f"| `{pretty[n]}` | {c['_total']} | {r[0]} | {r[1]} | " + " | ".join(str(c.get(w, 0)) for w in WATCH) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\nlibrary totals: " + ", ".join(f"{w} {tot.get(w, 0)}" for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
