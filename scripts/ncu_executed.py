#!/usr/bin/env python
"""Executed work per unit of the dominant kernel from an `ncu --set full` capture, for bench.py's `roofline_executed`
and `roofline.traffic`:
    python scripts/ncu_executed.py rep.ncu-rep <units per launch> <config> <fp_mode> [out.json]
Writes {kernel, warp_inst_per_unit, xu_warp_inst_per_unit, dram_bytes_per_launch, duration_ms, issue_active_pct, ...}."""
import csv, io, json, subprocess, sys

rep, units, config, fp_mode = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4]
out_path = sys.argv[5] if len(sys.argv) > 5 else "profiles/executed_r02.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
recs = [dict(zip(hdr, r)) for r in rows[2:]]
r = max(recs, key=lambda d: float(d["gpu__time_duration.sum"].replace(",", "")))     # the dominant launch of the capture


def f(k):
  return float(r[k].replace(",", "")) if r.get(k, "") != "" else None


unit_of = dict(zip(hdr, rows[1]))
dur = f("gpu__time_duration.sum")
dur_ms = dur / 1e6 if unit_of["gpu__time_duration.sum"] in ("ns", "nsecond") else (dur / 1e3 if unit_of["gpu__time_duration.sum"].startswith("us") else dur)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = sum(f(k) * scale.get(unit_of[k], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
# MUFU (XU pipe) warp instructions: the explicit counter when the capture has it, else from the pipe utilisation
# (peak 16 lanes/clk/SM = 0.5 warp instructions per cycle per SM) x active cycles x SMs
xu = f("smsp__inst_executed_pipe_xu.sum")
if xu is None:
  xu = f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active") / 100.0 * 0.5 * f("sm__cycles_active.avg") * 148
out = {
  "kernel": r["Kernel Name"], "config": config, "fp_mode": fp_mode, "units_per_launch": units, "report": rep.split("/")[-1],
  "duration_ms_under_ncu": dur_ms,
  "warp_inst_per_unit": f("smsp__inst_executed.sum") / units,
  "xu_warp_inst_per_unit": xu / units,
  "dram_bytes_per_launch": dram,
  "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
  "xu_pipe_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
  "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
  "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
  "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
  "l2_hit_pct": f("lts__t_sector_hit_rate.pct"), "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"),
  "registers": f("launch__registers_per_thread"), "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
}
with open(out_path, "w") as fh:
  json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
