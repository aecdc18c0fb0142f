#!/usr/bin/env python
"""(runs here, no GPU) Rewrites profiles/sweep_traffic.json from `ncu --set full` captures of the CURRENT build:
  tools/refresh_traffic.py gpurun_out/sweep_r02.ncu-rep [gpurun_out/sweep_q27_r02.ncu-rep ...]
bench.py copies the per-launch DRAM bytes into roofline.traffic together with `_source` (capture file, date,
git revision of the build that was profiled)."""
import csv
import datetime
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "sweep_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
used = []
for rep in sys.argv[1:]:
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U, V = rows[0], rows[1], rows[2]
    name = V[H.index("Kernel Name")]
    grid = [int(v) for v in re.findall(r"\d+", V[H.index("Grid Size")])]
    block = [int(v) for v in re.findall(r"\d+", V[H.index("Block Size")])]
    Q = int(re.search(r"sweep_kernel<\(?(?:int\))?(\d+)", name).group(1))

    def val(key):
        i = H.index(key)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[U[i]]
        return float(V[i].replace(",", "")) * scale
    n = round((grid[0] * grid[1] * grid[2] * block[0]) ** (1 / 3))
    d["D3Q%d_%d" % (Q, n)] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    used.append(os.path.basename(rep))
rev = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
d["_source"] = "ncu --set full --clock-control none, captures %s, %s, build at git %s" % (
    ", ".join(used), datetime.date.today().isoformat(), rev)
d["_comment"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch of sweep_kernel; rewritten by tools/refresh_traffic.py"
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
