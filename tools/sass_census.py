#!/usr/bin/env python
"""SASS opcode census of libb200blas.so per kernel (cuobjdump -sass): the mnemonics that prove which hardware path a
kernel uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, mma.sync f64 -> DMMA,
cp.async -> LDGSTS).  Writes a markdown table; run on the CPU box (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "STTM", "DMMA", "LDGSTS", "SYNCS", "FFMA", "DFMA", "HMMA"]


def main():
    so = os.path.join(ROOT, "eigen_b200", "libb200blas.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("b200::", "")
            name = re.sub(r"\(.*", "", re.sub(r"^void ", "", name))
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            cur["_total"] += 1
            if op in OPS:
                cur[op] += 1
    out = ["| kernel | instructions | " + " | ".join(OPS) + " |", "|---|---|" + "---|" * len(OPS)]
    tot = collections.Counter()
    for name, c in counts.items():
        tot.update(c)
        out.append("| `%s` | %d | %s |" % (name[:150], c["_total"], " | ".join(str(c[o]) if c[o] else "" for o in OPS)))
    out.append("| **all kernels** | %d | %s |" % (tot["_total"], " | ".join(str(tot[o]) for o in OPS)))
    text = ("# SASS opcode census of `eigen_b200/libb200blas.so` (sm_100a)\n\n`python tools/sass_census.py` = `cuobjdump -sass` "
            "per kernel; UTCHMMA = tcgen05.mma, UTMALDG = TMA loads, LDTM/STTM = tcgen05.ld/st, DMMA = mma.sync.m8n8k4.f64, LDGSTS = cp.async, "
            "SYNCS = mbarrier ops.  No HMMA (legacy mma.sync half path) anywhere.\n\n" + "\n".join(out) + "\n")
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_census_r02.md")
    open(dst, "w").write(text)
    print(text[-1500:])


if __name__ == "__main__":
    main()
