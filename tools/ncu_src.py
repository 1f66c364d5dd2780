"""Instruction mix and stall samples of one kernel of an ncu report (source page).
usage: python tools/ncu_src.py report.ncu-rep [launch index] [--lines N]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
print(rows[0][:2])
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
ops = collections.Counter(); smp = collections.Counter(); st = collections.Counter(); tot = tots = 0
lines = []
for r in rows[h + 1:]:
    if len(r) <= iex or not r[iex].isdigit(): continue
    t = r[isrc].split()
    if not t: continue
    op = t[1] if t[0].startswith("@") else t[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "BAR", "SYNCS")) else op.split(".")[0]
    n = int(r[iex]); k = int(r[ismp] or 0)
    ops[op] += n; tot += n; smp[op] += k; tots += k
    for i in stalls:
        if r[i].isdigit(): st[hdr[i]] += int(r[i])
    lines.append((k, n, r[0], r[isrc]))
print("warp instructions %.1f M, samples %d" % (tot / 1e6, tots))
for op, n in ops.most_common(28):
    print("%-16s %8.1fM %5.1f%%  samples %5.1f%%" % (op, n / 1e6, 100 * n / tot, 100 * smp[op] / max(tots, 1)))
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tots, 1)) for k, v in st.most_common(10)))
if "--lines" in sys.argv:
    nl = int(sys.argv[sys.argv.index("--lines") + 1])
    for k, n, a, s in sorted(lines, reverse=True)[:nl]:
        print("%5.2f%% %8.2fM  %s  %s" % (100 * k / max(tots, 1), n / 1e6, a, s[:110]))
