#!/usr/bin/env python
"""Busiest hardware units of one profiled kernel: every `...pct_of_peak_sustained_elapsed` metric of an
`ncu --set full` report, sorted.  Complements scripts/ncu_summary.py (which prints a fixed list).

    python scripts/ncu_busiest.py gpurun_out/r01n_warp.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, vals = rows[0], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"kernel: {name}")
    seen = []
    for h, v in zip(hdr, vals):
        if h.endswith(".avg.pct_of_peak_sustained_elapsed") or \
                h.endswith(".sum.pct_of_peak_sustained_elapsed"):
            try:
                seen.append((float(v.replace(",", "")), h))
            except ValueError:
                pass
    best = {}
    for v, h in seen:     # one line per metric (avg and sum variants repeat)
        key = h.rsplit(".", 2)[0]
        best[key] = max(best.get(key, 0.0), v)
    for key, v in sorted(best.items(), key=lambda kv: -kv[1])[:top]:
        print(f"  {v:6.1f} %  {key}")


if __name__ == "__main__":
    main()
