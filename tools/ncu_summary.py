"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py file.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct"]
out = []
for w in want:
    if w in idx:
        out.append("%-70s %-8s %s" % (w, units[idx[w]], " | ".join(r[idx[w]][:40] for r in data)))
st = [h for h in hdr if h.startswith("smsp__average_warp") and h.endswith("per_issue_active.ratio")]
out.append("-- warp stall reasons (cycles per issued instruction), > 0.1 in any launch")
for h in st:
    vals = [float(r[idx[h]] or 0) for r in data]
    if max(vals) > 0.1:
        name = h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
        out.append("%-70s %s" % (name.replace("_per_issue_active.ratio", ""), " | ".join("%.2f" % v for v in vals)))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
