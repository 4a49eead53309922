"""Dev tool: turn an ncu report (gpurun_out/*.ncu-rep) into the committed summary under profiles/.
usage: python tools/summarize_profile.py gpurun_out/prof.ncu-rep profiles/r01_ncu_fused_cfg2.json [images_in_launch]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
images = int(sys.argv[3]) if len(sys.argv) > 3 else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
summary = {k: {"value": d[k][0], "unit": d[k][1]} for k in keys if k in d}
stalls = {}
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        v = float(d[h][0] or 0)
        if v > 0.02:
            stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 3)
summary["stall_cycles_per_issued_instruction"] = stalls
rd = float(d["dram__bytes_read.sum"][0]); wr = float(d["dram__bytes_write.sum"][0])
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
summary["dram_bytes_per_launch"] = rd * scale[d["dram__bytes_read.sum"][1]] + wr * scale[d["dram__bytes_write.sum"][1]]
summary["images_in_launch"] = images
summary["source_report"] = rep
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps(summary, indent=1)[:1500])
