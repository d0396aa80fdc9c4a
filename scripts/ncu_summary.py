"""Text summary of one `ncu --set full` report (the files under profiles/):
   python scripts/ncu_summary.py report.ncu-rep > profiles/<tag>_<kernel>_ncu_summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
print(f"# ncu --set full --clock-control none, one launch; source: {rep} (scratch, not committed)")
print("kernel:", vals[col["Kernel Name"]])
for w in want:
    if w in col:
        print(f"  {w} = {vals[col[w]]} {units[col[w]]}")
print("  warp stall reasons (cycles per issued instruction):")
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        try:
            v = float(vals[col[h]])
        except ValueError:
            continue
        if v >= 0.1:
            name = h.split("issue_stalled_")[1].split("_per_issue")[0]
            print(f"    {name} = {v:.2f}")
