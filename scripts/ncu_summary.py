#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (ncu --set full) for profiles/: python scripts/ncu_summary.py rep [out.md]"""
import csv, io, subprocess, sys

KEYS = [
  "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
  "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
  "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
  "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
  "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
  "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
  "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
  "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
  "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum",
  "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
  "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main():
  rep = sys.argv[1]
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units = rows[0], rows[1]
  lines = []
  for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append(f"## {d.get('Kernel Name', '?')}  (id {d.get('ID')}, grid {d.get('Grid Size')}, block {d.get('Block Size')})\n")
    lines.append("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
      if k in d and d[k] != "":
        lines.append(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
    lines.append("")
  text = "\n".join(lines)
  if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
      f.write(f"# ncu --set full summary of {rep.split('/')[-1]}\n\n" + text + "\n")
  else:
    print(text)


if __name__ == "__main__":
  main()
