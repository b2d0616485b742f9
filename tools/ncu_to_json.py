#!/usr/bin/env python3
"""profiles/fast_nms_ncu.json from a `ncu --page raw --csv` export of one fast_nms_tma_kernel launch: the DRAM traffic and pipe
utilisation bench.py quotes in its `roofline` entry (traffic, ALU-pipe and issue-slot utilisation scaled to the live time).
Usage: python tools/ncu_to_json.py profiles/r2_fast_nms_final_raw.csv FRAMES_IN_LAUNCH [WHAT] > profiles/fast_nms_ncu.json"""
import json
import sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_summary import summarise

path, frames = sys.argv[1], int(sys.argv[2])
what = sys.argv[3] if len(sys.argv) > 3 else "device-resident launch"
rows = [d for d in summarise(path) if "fast_nms" in d["kernel"]]
d = rows[-1]
print(json.dumps({
    "kernel": d["kernel"].replace("void ", ""),
    "source": "%s (ncu --set full --clock-control none --import-source on, %d-frame launch: %s)" % (path, frames, what),
    "frames_in_launch": frames,
    "dram_bytes_read": d["dram_read_B"], "dram_bytes_write": d["dram_write_B"],
    "duration_us": d["duration_us"],
    "alu_pipe_pct_of_peak": d["alu_pct"], "fma_pipe_pct_of_peak": d["fma_pct"], "lsu_pipe_pct_of_peak": d["lsu_pct"],
    "issue_active_pct": d["issue_pct"],
    "alu_peak_hw_thread_inst_per_clk_sm": 64, "alu_peak_measured_thread_inst_per_clk_sm": 58.7,
    "registers": int(d["regs"]), "grid": int(d["grid"]), "warp_inst": d["warp_inst"],
}, indent=1))
