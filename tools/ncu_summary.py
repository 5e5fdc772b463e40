#!/usr/bin/env python
"""Summarises .ncu-rep captures (run here, no GPU needed) into profiles/<name>.md + .json.
Usage: python tools/ncu_summary.py <tag> <rep> [<rep> ...]"""
import csv
import json
import os
import re
import subprocess
import sys

PICK = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio")
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    res = []
    for vals in rows[hi + 2:]:
        if len(vals) == len(hdr):
            res.append({h: (u, v) for h, u, v in zip(hdr, units, vals)})
    return res


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    os.makedirs("profiles", exist_ok=True)
    for rep in reps:
        for k, m in enumerate(load(rep)):
            name = m["Kernel Name"][1]
            short = re.sub(r"[^a-z0-9_]+", "", name.split("(")[0].split("::")[-1].lower())[:40]
            d = {"kernel": name, "source_report": os.path.basename(rep), "metrics": {}, "stalls_per_issue": {}}
            for key in PICK:
                if key in m:
                    u, v = m[key]
                    try:
                        v = float(v.replace(",", ""))
                    except ValueError:
                        pass
                    d["metrics"][key] = {"value": v, "unit": u}
            for key, (u, v) in m.items():
                s = STALL.match(key)
                if s:
                    d["stalls_per_issue"][s.group(1)] = float(v)
            br, bw = d["metrics"].get("dram__bytes_read.sum"), d["metrics"].get("dram__bytes_write.sum")
            if br and bw:
                d["dram_bytes_per_launch"] = br["value"] * UNIT.get(br["unit"], 1) + bw["value"] * UNIT.get(bw["unit"], 1)
            base = f"profiles/{tag}_{short}" + (f"_{k}" if k else "")
            json.dump(d, open(base + ".json", "w"), indent=1)
            with open(base + ".md", "w") as f:
                f.write(f"# ncu --set full --clock-control none: `{name[:90]}`\n\nreport: `{os.path.basename(rep)}` (gpurun_out/, scratch) -- launch {k}\n\n")
                f.write("| metric | value | unit |\n|---|---|---|\n")
                for key, x in d["metrics"].items():
                    f.write(f"| {key} | {x['value']} | {x['unit']} |\n")
                if "dram_bytes_per_launch" in d:
                    f.write(f"| **dram bytes per launch (read+write)** | {d['dram_bytes_per_launch']:.0f} | byte |\n")
                f.write("\nWarp stall reasons (warps stalled per issue-active cycle):\n\n| reason | ratio |\n|---|---|\n")
                for r, v in sorted(d["stalls_per_issue"].items(), key=lambda kv: -kv[1]):
                    if v > 0.01:
                        f.write(f"| {r} | {v:.3f} |\n")
            print("wrote", base + ".md")


if __name__ == "__main__":
    main()
