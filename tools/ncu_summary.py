#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here on the CPU box with `ncu -i`) into the small text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_k1_f32.ncu-rep --tag r01_k1_f32 [--traffic-json profiles/k1_traffic_f32.json]

Writes profiles/<tag>.details.txt (ncu --page details), profiles/<tag>.raw_selected.csv (the metrics the roofline
uses) and, with --traffic-json, the per-launch DRAM traffic that bench.py copies into `roofline.traffic`.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"dram__bytes_(read|write)\.sum$|dram__cycles_active|gpu__dram_throughput|gpu__time_duration\.sum|"
                  r"sm__warps_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
                  r"launch__shared_mem_per_block|launch__occupancy_limit|sm__throughput\.avg\.pct|"
                  r"smsp__inst_executed\.sum$|sm__inst_executed_pipe_fma|smsp__issue_active\.avg\.pct|"
                  r"lts__t_bytes\.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|sm__cycles_active\.avg$|"
                  r"smsp__average_warp_latency|smsp__warp_issue_stalled.*_per_warp_active\.pct$")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--traffic-json", default="")
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    details = subprocess.run(["ncu", "-i", a.rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(out_dir, a.tag + ".details.txt"), "w").write(details)
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        sys.exit("no rows in " + a.rep)
    header, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(header) if h in ("ID", "Kernel Name", "Grid Size", "Block Size") or KEEP.search(h)]
    with open(os.path.join(out_dir, a.tag + ".raw_selected.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([header[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i in cols])
    idx = {h: i for i, h in enumerate(header)}

    def num(r, key):
        v = r[idx[key]].replace(",", "")
        u = units[idx[key]]
        x = float(v)
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

    per = []
    for r in rows[2:]:
        per.append({"kernel": r[idx["Kernel Name"]][:80], "dram_read": num(r, "dram__bytes_read.sum"),
                    "dram_write": num(r, "dram__bytes_write.sum"), "duration_ns": float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
                    * {"us": 1e3, "ns": 1.0, "ms": 1e6}.get(units[idx["gpu__time_duration.sum"]], 1.0)})
    for p in per:
        print(p)
    if a.traffic_json:
        rd = sum(p["dram_read"] for p in per) / len(per)
        wr = sum(p["dram_write"] for p in per) / len(per)
        json.dump({"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "launches": len(per),
                   "source": f"profiles/{a.tag}.raw_selected.csv (ncu --set full, mean of {len(per)} launches; the gradient "
                             f"stores that are still in L2 when the kernel ends are written back later and not counted)"},
                  open(a.traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main()
