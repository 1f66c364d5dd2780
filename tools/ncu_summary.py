"""Turn the ncu artefacts brought back in gpurun_out/ into the markdown summaries kept under profiles/.
usage: python tools/ncu_summary.py launches <launch_list.csv>       -> per-kernel share table
       python tools/ncu_summary.py full <report.ncu-rep>            -> per-launch metric table"""
import collections, csv, subprocess, sys, io


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM rd"),
            ("dram__bytes_write.sum", "DRAM wr"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %pk"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
            ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
            ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math")]
    cols = [(c, n) for c, n in cols if c in hdr]
    print("| # | kernel | " + " | ".join("%s%s" % (n, (" " + units[hdr.index(c)]) if n.startswith("DRAM r") or n.startswith("DRAM w") else "") for c, n in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for k, r in enumerate(data):
        name = r[hdr.index("Kernel Name")]
        vals = []
        for c, n in cols:
            v = r[hdr.index(c)].replace(",", "")
            try:
                f = float(v)
                vals.append("%.4g" % f)
            except ValueError:
                vals.append(v)
        print("| %d | `%s` | %s |" % (k, name[:58], " | ".join(vals)))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
