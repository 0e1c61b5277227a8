"""Summarise an .ncu-rep (read here, no GPU) into a small JSON + text file for profiles/."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum"]
res = []
for r in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, r):
        if h in KEYS:
            d[h] = v if h == "Kernel Name" else f"{v} {u}".strip()
    stalls = [(h.split("issue_stalled_")[1].split("_per_warp_active")[0], float(v)) for h, v in zip(hdr, r)
              if "average_warps_issue_stalled" in h and h.endswith("per_warp_active.pct") and "not_issued" not in h]
    stalls.sort(key=lambda t: -t[1])
    d["top_stalls_pct_of_warp_active"] = {k: round(v, 2) for k, v in stalls[:8]}
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
