#!/usr/bin/env python
"""ncu_summary.py <report.ncu-rep> <out_prefix>

Turns an `ncu --set full` report into the two small files profiles/ keeps:
  <out_prefix>_summary.csv   one row per captured launch: duration, DRAM bytes, throughputs, occupancy, issue-active, top stalls
  <out_prefix>_traffic.json  DRAM bytes per launch (read + write) of the a-trous kernels, averaged over the captured iterations —
                             bench.py reads the newest of these for `roofline.traffic` (no constant in the code)
Runs where ncu is installed (the build container can read reports captured on the GPU box).
"""
import csv
import json
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
cols = [c for c in cols if c in idx]
stall = [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


out_rows, traffic = [], {}
for r in rows[2:]:
    st = []
    for k in stall:
        try:
            st.append((float(r[idx[k]]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    st.sort(reverse=True)
    name = r[idx["Kernel Name"]]
    out_rows.append([r[idx[c]] for c in cols] + ["; ".join(f"{n} {v:.2f}" for v, n in st[:4])])
    if "atrous" in name:
        key = "reflections" if "refl" in name else "shadows"
        b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic.setdefault(key, []).append(b)
with open(prefix + "_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([c + (f" [{units[idx[c]]}]" if units[idx[c]] else "") for c in cols] + ["top stalls (warps per issue)"])
    w.writerows(out_rows)
if traffic:
    js = {k: sum(v) / len(v) for k, v in traffic.items()}
    js["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of {', '.join(str(len(v)) for v in traffic.values())} captured a-trous launches, ncu --set full, {rep.split('/')[-1]}"
    js["_per_launch"] = traffic
    json.dump(js, open(prefix + "_traffic.json", "w"), indent=1)
print(f"{len(out_rows)} launches -> {prefix}_summary.csv" + (f", {prefix}_traffic.json" if traffic else ""))
