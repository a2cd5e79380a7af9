#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep (ncu --set full capture) as JSON lines: duration, instructions,
issue / pipe utilisation, DRAM and L2 traffic, hit rates, occupancy.  usage: ncu_summary.py file.ncu-rep [units ...]
(units: optional divisor per kernel in capture order, e.g. mover-steps or queries of that launch)."""
import csv, json, subprocess, sys
WANT = {
    "gpu__time_duration.sum": "time", "smsp__inst_executed.sum": "warp_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sectors.sum": "l2_sectors", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_sectors.sum.pct_of_peak_sustained_elapsed": "l2_sector_throughput_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_throughput_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
kn = h.index("Kernel Name")
div = [float(x) for x in sys.argv[2:]]
for k, r in enumerate(rows[2:]):
    out = {"kernel": r[kn].split("(")[0]}
    for i, c in enumerate(h):
        if c in WANT and r[i] != "":
            v = float(r[i].replace(",", ""))
            v *= UNIT.get(units[i], 1.0)
            out[WANT[c]] = v
    if "l2_sectors" in out:
        out["l2_bytes"] = 32.0 * out["l2_sectors"]
    if "l1_global_load_sectors" in out:
        out["l1_global_load_bytes"] = 32.0 * out["l1_global_load_sectors"]
    if k < len(div):
        u = div[k]
        out["units"] = u
        out["warp_inst_per_unit"] = out.get("warp_inst", 0) / u
        out["dram_bytes_per_unit"] = (out.get("dram_read", 0) + out.get("dram_write", 0)) / u
        out["l2_bytes_per_unit"] = out.get("l2_bytes", 0) / u
    print(json.dumps(out))
