#!/usr/bin/env python3
"""Opcode histogram of the built library, per kernel (cuobjdump -sass; no GPU needed): the evidence that the hot kernels
use TMA (UTMALDG), FP64 tensor cores (DMMA) and no local-memory spills (STL/LDL).  Writes profiles/r02_sass_summary.md.

    python tools/sass_summary.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantum-simulation_b200", "lib", "libqsb200.so")
KEYS = ["DMMA", "UTMALDG", "UTMAPF", "UTMACCTL", "UTMASTG", "LDGSTS", "LDG", "STG", "LDS", "STS", "SYNCS", "STL", "LDL", "BAR",
        "ATOM", "RED", "DFMA", "DADD", "DMUL", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    arch = set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            per[cur][op] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    for k, n in zip(per, names):
        demangle[k] = n
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    lines = [f"# SASS opcode summary of `lib/libqsb200.so` (cuobjdump -sass; archs: {', '.join(sorted(arch))})", "",
             f"{len(per)} kernels.  Whole library: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]) + ".", "",
             "tcgen05 (`UTC*MMA`, `LDTM`/`STTM`) does not appear by design: tcgen05 has no FP64 kind, the FP64 tensor path of "
             "sm_100a is the warp-level `DMMA.8x8x4` (SURVEY headline 5).  `STL`/`LDL` = local-memory traffic (spills / stack).",
             "", "| kernel | instr | " + " | ".join(KEYS[:16]) + " |", "|---|---|" + "---|" * 16]
    for k, c in per.items():
        n = sum(c.values())
        name = demangle[k]
        name = re.sub(r"\(.*", "", name)[-90:]
        lines.append(f"| `{name}` | {n} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS[:16]) + " |")
    out = os.path.join(ROOT, "profiles", "r02_sass_summary.md")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:4]))
    print("wrote", out)


if __name__ == "__main__":
    main()
