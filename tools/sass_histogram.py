"""Per-kernel SASS opcode histogram of libjudi_b200.so (cuobjdump -sass): the evidence that the hot
kernels are sm_100a code using TMA (UTMALDG), packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2),
128-bit shared loads (LDS.128) and mbarriers (SYNCS), without running anything on a GPU.

    python tools/sass_histogram.py [pattern ...] > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "judi.jl_b200", "libjudi_b200.so")
KEY = ("UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "LDS.128",
       "LDS", "STS.128", "STS", "LDG.E.128", "LDG", "LD.E.128", "STG.E.128", "STG", "ST.E.128", "BAR", "MOV", "BRA",
       "IMAD", "MUFU.RCP")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return out.splitlines()


def main(patterns):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur, arch = collections.OrderedDict(), None, None
    for line in txt.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = [arch, collections.Counter()]
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur][1][m.group(1)] += 1
    names = list(kernels)
    pretty = dict(zip(names, demangle(names)))
    print("# SASS opcode histogram of libjudi_b200.so (static instruction counts per kernel)\n")
    print("`cuobjdump -sass judi.jl_b200/libjudi_b200.so`, %d kernels; counts are static (instructions "
          "in the binary, not executed).\n" % len(names))
    print("| kernel | arch | total | " + " | ".join(KEY) + " |")
    print("|---|---|---:|" + "---:|" * len(KEY))
    tot = collections.Counter()
    for n in names:
        p = pretty[n]
        if patterns and not any(q in p for q in patterns):
            continue
        arch, c = kernels[n]

        def cnt(k):
            # exact mnemonic or mnemonic with modifiers; LDS / STS / LDG / STG exclude the .128 rows
            v = sum(x for op, x in c.items() if op == k or op.startswith(k + "."))
            if k in ("LDS", "STS"):
                v -= sum(x for op, x in c.items() if op.startswith(k + ".128"))
            if k in ("LDG", "STG"):
                v -= sum(x for op, x in c.items() if op.startswith(k + ".E.128"))
            if k in ("FFMA", "FADD", "FMUL"):
                v = sum(x for op, x in c.items() if op == k or (op.startswith(k + ".") and not op.startswith(k + "2")))
            return v
        row = [cnt(k) for k in KEY]
        for k, v in zip(KEY, row):
            tot[k] += v
        short = re.sub(r"\(.*", "", p)
        print("| `%s` | %s | %d | %s |" % (short[:110], arch, sum(c.values()), " | ".join(str(v) for v in row)))
    print("\nTotals over the listed kernels: " + ", ".join("%s %d" % (k, tot[k]) for k in KEY if tot[k]))


if __name__ == "__main__":
    main(sys.argv[1:])
