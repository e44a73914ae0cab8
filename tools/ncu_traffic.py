#!/usr/bin/env python
"""Write profiles/ncu_traffic.json (read by bench.py for roofline.traffic) from `ncu --set full` reports.

    tools/ncu_traffic.py key=report.ncu-rep[:kernel-regex][:samples] ...

For each key the first kernel of the report whose name matches the regex gives
dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) and its duration under ncu.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TSCALE = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        out = json.load(open(out_path))
    except Exception:
        out = {}
    for spec in sys.argv[1:]:
        key, rest = spec.split("=", 1)
        parts = rest.split(":")
        rep, rx = parts[0], (parts[1] if len(parts) > 1 and parts[1] else ".")
        samples = int(parts[2]) if len(parts) > 2 else None
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if re.search(rx, r[ix["Kernel Name"]]):
                def val(name, scale):
                    return float(r[ix[name]].replace(",", "")) * scale[units[ix[name]]]
                rd, wr = val("dram__bytes_read.sum", SCALE), val("dram__bytes_write.sum", SCALE)
                e = {"bytes": rd + wr, "read": rd, "write": wr, "ms_under_ncu": val("gpu__time_duration.sum", TSCALE),
                     "kernel": r[ix["Kernel Name"]].split("(")[0][-80:], "source": os.path.basename(rep)}
                if samples:
                    e["samples"] = samples
                out[key] = e
                break
        else:
            print("no kernel matching", rx, "in", rep)
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
