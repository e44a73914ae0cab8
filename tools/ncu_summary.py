#!/usr/bin/env python
"""Summarise ncu output into profiles/: tools/ncu_summary.py <report.ncu-rep>[,<report2>...] <launches.csv> <out.md> [title]

  <report.ncu-rep>  from `ncu --set full --clock-control none --import-source on ...` (read with ncu -i, no GPU);
                    or the `ncu -i <rep> --page raw --csv` export of one (round 2: the .ncu-rep files are exported to
                    CSV on the GPU box, they are too big to bring back); "-" = no launch list
  <launches.csv>    from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall / issue"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]


def short(name):
    name = name.replace("void ", "").replace("invpref::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0][:70]


def main():
    rep, launches, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else rep
    lines = [f"# {title}", "", "## Kernels captured with `ncu --set full --clock-control none` (one launch each)", ""]
    captured = []
    for one in rep.split(","):
        if one.endswith(".csv"):
            raw = open(one).read()
        else:
            raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        captured += [(rows[0], rows[1], r, one) for r in rows[2:]]
    for hdr, units, r, one in captured:
        idx = {h: i for i, h in enumerate(hdr)}
        lines.append(f"### `{short(r[idx['Kernel Name']])}`  ({one.split('/')[-1]})")
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                lines.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        if "dram__bytes_read.sum" in idx:
            def gb(v, u):
                v = float(v.replace(",", ""))
                return v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1)
            tr = gb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                gb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            d = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            du = units[idx["gpu__time_duration.sum"]]
            d_ms = d * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(du, 1)
            lines.append(f"| **traffic** (dram read + write) | {tr:.3f} GB -> {tr / d_ms * 1e3:.0f} GB/s under ncu |")
        lines.append("")
    # launch list: share of each kernel in the step
    agg = OrderedDict()
    total = 0.0
    if launches == "-":
        open(out, "w").write("\n".join(lines))
        print("wrote", out)
        return
    with open(launches) as f:
        rd = [r for r in csv.reader(l for l in f if not l.startswith("==")) if r]
    h = rd[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    for r in rd[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum`, whole bench process: cold-cache, serialised; "
              "compare SHARES)", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {v:.3f} | {100 * v / total:.1f} % |")
    lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main()
