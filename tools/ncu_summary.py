"""Summarise an .ncu-rep (raw + source pages) into a short text for profiles/."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
for row in rows[2:]:
    print("kernel:", row[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, row):
        if h in KEYS:
            print(f"  {h:88s} {v:>18s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ia, isrc, iall, iex = (hdr.index(k) for k in ("Address", "Source", "Warp Stall Sampling (All Samples)", "Instructions Executed"))
data = [r for r in rows[2:] if len(r) > iall]
tot = sum(int(r[iall] or 0) for r in data)
print(f"stall samples: {tot}; top instructions:")
for r in sorted(data, key=lambda r: -int(r[iall] or 0))[:14]:
    print(f"  {int(r[iall] or 0) / max(tot, 1) * 100:5.1f}%  x{r[iex]:>10s}  {r[isrc][:80]}")
ops = {}
for r in data:
    op = r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]
    ops[op] = ops.get(op, 0) + int(r[iex] or 0)
tot_i = sum(ops.values())
print("executed warp-instructions by opcode:")
for op, c in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
    print(f"  {c / tot_i * 100:5.1f}%  {op}")
