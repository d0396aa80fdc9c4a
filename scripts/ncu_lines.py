"""Per-CUDA-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo
and --import-source on).   python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) >= 8 and r[0].strip().isdigit() and r[2] == "-":
        try:
            ex, sm = int(r[7] or 0), int(r[6] or 0)
        except ValueError:
            continue
        a = agg.setdefault((fname, int(r[0])), [r[1].strip(), 0, 0])
        a[1] += ex
        a[2] += sm
tot = sum(a[1] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print(f"total warp instructions {tot}, stall samples {tots}")
for (f, ln), (src, ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ex / tot * 100:5.1f}% inst {sm / tots * 100:5.1f}% smp  {f}:{ln:<4d} {src[:100]}")
