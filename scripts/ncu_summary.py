"""Summarise one kernel of an .ncu-rep into the JSON kept under profiles/ (run in the build container).
usage: python scripts/ncu_summary.py REPORT.ncu-rep OUT.json "note"  [extra-metric-substring ...]"""
import csv, io, json, subprocess, sys

rep, out, note = sys.argv[1:4]
extra = sys.argv[4:]
KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_tensor.sum"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
res = {"note": note}
for i, h in enumerate(hdr):
    if h == "Kernel Name" or h in KEYS or any(e in h for e in extra):
        res[h] = {"value": vals[i], "unit": units[i]}
stalls = {h: float(vals[i].replace(",", "")) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals[i]}
if stalls:
    res["stalls_per_issue"] = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 3)
                               for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:3000])
