#!/usr/bin/env python3
"""Summarise `ncu --page raw --csv` exports (profiles/*_raw.csv) into one line per kernel launch:
duration, DRAM bytes, pipe utilisation, issue rate, registers, occupancy.  Usage:
    python tools/ncu_summary.py profiles/r2_*_raw.csv [--json]
The raw CSV has one header row, one units row and one row per profiled launch."""
import csv
import json
import re
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_B",
    "dram__bytes_write.sum": "dram_write_B",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed.sum": "thread_inst",
    "sm__inst_executed_pipe_alu.sum": "alu_warp_inst",
    "smsp__inst_executed_pipe_alu.sum": "alu_warp_inst2",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def summarise(path):
    rows = list(csv.reader(open(path, newline="")))
    while rows and (not rows[0] or rows[0][0] != "ID"):   # ncu banner lines ("==PROF== ...") before the header
        rows.pop(0)
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(hdr):
        for k in KEYS:
            if h == k or h.endswith("." + k):
                col.setdefault(KEYS[k], i)
    name_i = hdr.index("Kernel Name")
    out = []
    for r in data:
        full = r[name_i]
        m = re.match(r"^(.*?>)\(", full) if "<" in full else None     # template kernels: keep the <...> argument list
        d = {"kernel": (m.group(1) if m else full.split("(")[0]).replace("void ", "").replace("orbfe::", "").replace("(int)", "").replace(" ", ""),
             "file": path}
        for key, i in col.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT_SCALE.get(units[i], 1.0)
            d[key] = v
        out.append(d)
    return out


if __name__ == "__main__":
    as_json = "--json" in sys.argv
    res = []
    for p in [a for a in sys.argv[1:] if not a.startswith("--")]:
        res += summarise(p)
    if as_json:
        print(json.dumps(res, indent=1))
    else:
        for d in res:
            print(" ".join("%s=%s" % (k, ("%.4g" % v) if isinstance(v, float) else v) for k, v in d.items()))
