"""`ncu -i X.ncu-rep --page raw --csv` -> small JSON of the metrics the profiles/ summaries quote, one record per
captured launch.  usage: python tools/ncu_extract.py gpurun_out/X.ncu-rep [extra-metric-regex] > profiles/X.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__cluster_size", "launch__grid_size", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg"]

import re
EXTRA = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None  # optional: regex of further metric names to keep
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
recs = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    rec = {"name": d["Kernel Name"][:120], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in KEEP + ([h for h in hdr if EXTRA.search(h) and h not in KEEP] if EXTRA else []):
        if k in d and d[k] != "":
            u = units[hdr.index(k)]
            try:
                rec[f"{k} [{u}]"] = float(d[k].replace(",", ""))
            except ValueError:
                rec[f"{k} [{u}]"] = d[k]
    def get(k):
        u = units[hdr.index(k)] if k in hdr else ""
        v = rec.get(f"{k} [{u}]")
        if v is None:
            return None
        scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, None)
        return v * scale if scale else None
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        rec["dram_traffic_GB"] = round(rd + wr, 4)
    recs.append(rec)
print(json.dumps(recs, indent=1))
