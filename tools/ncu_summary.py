#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture of the three encode kernels into profiles/dram_traffic.json.

usage: python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep <details-file-name-under-profiles>

bench.py reads the file for `roofline.traffic` (dram bytes per launch of the dominant kernel).
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed.sum": "thread_inst",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc",
}


def main():
    rep, details = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    per = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]  # wp_split_kernel<3072> -> wp_split_kernel
        k = {}
        for m, short in METRICS.items():
            if m in d and d[m] != "":
                v = float(d[m].replace(",", ""))
                if short.endswith("_MB") and u[m] == "Gbyte":
                    v *= 1000.0
                if short.endswith("_MB") and u[m] == "Kbyte":
                    v /= 1000.0
                if short == "duration_us" and u[m] == "ms":
                    v *= 1000.0
                if short == "duration_us" and u[m] == "ns":
                    v /= 1000.0
                k[short] = v
        if "thread_inst" in k and "warp_inst" in k:
            k["lanes_per_inst"] = round(k.pop("thread_inst") / k["warp_inst"], 2)
        k["dram_bytes"] = int(round((k["dram_read_MB"] + k["dram_write_MB"]) * 1e6))
        per[name] = k
    res = {
        "per_kernel": per,
        "source": f"ncu --set full --clock-control none, profiles/{details}: dram__bytes_read.sum + dram__bytes_write.sum "
                  "of one launch of each kernel over one full 64 MiB range of the en workload "
                  "(139 MB algorithmic bytes per range); cold caches, so the vocabulary table and the word memo "
                  "are read from DRAM once per kernel here",
    }
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dram_traffic.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
