"""Record the DRAM traffic of one kernel from an `ncu --set full` report into
profiles/ncu_traffic.json, stamped with the SHA-256 of the kernel's source file so that
bench.py reports it only while the source is unchanged.

    python scripts/ncu_traffic.py <report.ncu-rep> <key: warp|decode|...> <path/to/kernel.cu> <source label>
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key, cu, label = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
total = 0.0
for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    total += float(vals[col[m]]) * scale[units[col[m]]]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
with open(os.path.join(ROOT, cu), "rb") as f:
    sha = hashlib.sha256(f.read()).hexdigest()
data[key] = {"kernel": vals[col["Kernel Name"]], "dram_bytes": int(total), "cu": cu,
             "cu_sha256": sha, "source": label,
             "duration_us": float(vals[col["gpu__time_duration.sum"]])}
with open(path, "w") as f:
    json.dump(data, f, indent=1)
print(key, data[key])
