#!/usr/bin/env python
"""Opcode counts per kernel of libpherofield.so (cuobjdump -sass): the evidence for what the kernels are made of —
TMA bulk copies (UBLKCP), mbarrier traffic (SYNCS), 128-bit shared / global accesses, reductions performed by the
L2 (RED / REDG), and the absence of FFMA contraction on the parity path.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "swarm_robotics_pheromones_b200" / "lib" / "libpherofield.so"
WATCH = ("UBLKCP", "UTMALDG", "SYNCS", "LDS.128", "LDS", "STS", "LDG.E.128", "LDG", "STG.E.128", "STG", "RED", "REDG",
         "ATOMG", "ATOMS", "FADD", "FMUL", "FFMA", "MUFU", "SHFL", "BAR", "NANOSLEEP", "MATCH", "VOTE")


def demangle(names):
    out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    print(f"# SASS summary of {LIB.relative_to(ROOT)} — cubin arch: {', '.join(arch)}; {len(kernels)} kernels")
    print("# columns: total instructions, then counts of opcodes starting with the given mnemonic\n")
    for k, c in kernels.items():
        total = sum(c.values())
        name = re.sub(r"\(.*", "", names[k])[:110]
        parts = []
        for w in WATCH:
            n = sum(v for op, v in c.items() if op == w or op.startswith(w + "."))
            if n:
                parts.append(f"{w} {n}")
        print(f"{name}\n    total {total}; " + ", ".join(parts))
    return 0


if __name__ == "__main__":
    sys.exit(main())
