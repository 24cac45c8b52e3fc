"""Summarise an .ncu-rep (run here on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active.ratio" in h and "not_issued" not in h]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if filt and filt not in name:
        continue
    print("=====", name[:70], r[idx["Grid Size"]], r[idx["Block Size"]])
    for k in want:
        if k in idx:
            print(f"  {k:88s} {r[idx[k]]} {units[idx[k]]}")
    st = sorted([(float(r[idx[k]] or 0), k) for k in stall], reverse=True)[:7]
    print("  stalls/issue: " + ", ".join(f'{k.split("issue_stalled_")[1].split("_per_issue")[0]}={v:.2f}' for v, k in st))
