"""Selected metrics of the first kernel of an .ncu-rep as JSON.
usage: python scripts/ncu_summary.py X.ncu-rep "capture note" > profiles/r02_x_ncu_full.json"""
import csv, json, subprocess, sys
rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ("Kernel Name", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "sm__warps_active.avg.pct", "sm__throughput.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__pipe_tensor", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency",
        "smsp__thread_inst_executed_per_inst_executed", "smsp__inst_executed_op_local")
d = {"_capture": note}
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in KEEP) and ".per_second" not in h and "pct_of_peak_sustained_elapsed" not in h or h in (
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        if v not in ("",):
            d[h] = f"{v} {u}".strip()
print(json.dumps(d, indent=1))
