#!/usr/bin/env python
"""Benchmark of the hierarchical-likelihood hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--fp-mode fp32|fp64]

Workload (config.workload): BASELINE.json configs[2], the north-star target -- the O5-like mock:
1000 events x 5000 posterior samples, 10^6 detected injections, 256 hyper-points on a 16x16
H0 x Om0 grid, galaxy-catalogue likelihood (pixelated, ~15 pixels/event, kind 'approximate'),
1-D Gaussian KDE without binning, z_int_res = 300.  One step = one evaluation of the whole
hyper-point batch: population reweighting -> KDE -> z-integral for every (event, hyper-point)
plus the injection-reweighted selection function for every hyper-point.
metric = hyper-point x event log-likelihood evaluations per second (whole job, all GPUs).

Multi-GPU (torchrun, one rank per GPU): WEAK scaling -- every rank holds its own 1000 events and
10^6 injections (seeded per rank), all ranks evaluate all 256 hyper-points, one NCCL all-reduce of
the (256, 3) partials per step.

`--impl reference` times the CPU oracle restatement of the reference (JAX is not installable
offline) with all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hyper-point x event log-likelihood evaluations per second"
UNIT = "evals/s"
WORKLOAD = ("C3 O5-like mock: 1000 events x 5000 samples, 1e6 injections, 256 hyper-points (16x16 H0 x Om0), "
            "pixelated galaxy catalogue (~15 px/event, 'approximate'), Gaussian KDE unbinned, z_int_res=300")


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--fp-mode", default="fp32", choices=["fp32", "fp64"])
  ap.add_argument("--nev", type=int, default=1000)
  ap.add_argument("--ns", type=int, default=5000)
  ap.add_argument("--ninj", type=int, default=1_000_000)
  ap.add_argument("--nz", type=int, default=300)
  ap.add_argument("--hyper-side", type=int, default=16)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--options", default="", help="per-handle tuning switches for A/B runs, e.g. 'fused=0,kde_win=0' (chb_set_option)")
  return ap.parse_args()


# ------------------------------------------------------------------------------------------ workload
def build_workload(args, rank):
  from chimera_b200 import synth
  ev = synth.make_events(args.nev, args.ns, seed=1234 + 1000 * rank, sky=True)
  zg = synth.make_z_grids(ev["dL"], z_int_res=args.nz, H0_prior=(20., 200.))
  ev = synth.pixelize(ev, nside_list=(512, 256, 128, 64, 32, 16, 8), mean_npixels_event=15, sky_conf=0.9)
  p_cat, P_compl = synth.smooth_p_cat(ev, zg, seed=9012 + rank)
  inj, N_inj = synth.make_injections(args.ninj, seed=5678 + 1000 * rank)
  side = args.hyper_side
  H0, Om0 = np.meshgrid(np.linspace(55., 85., side), np.linspace(0.15, 0.45, side), indexing="ij")
  hyper = dict(H0=H0.ravel(), Om0=Om0.ravel())
  return dict(ev=ev, zg=zg, p_cat=p_cat, P_compl=P_compl, inj=inj, N_inj=N_inj, hyper=hyper,
              z_range=np.array([0.073, 1.3]))


def parse_options(text):
  return {k: float(v) for k, v in (kv.split("=") for kv in text.split(",") if kv)}


def build_likelihood(w, fp_mode, distributed, kernel="gauss", binning=False, options=None):
  import chimera_b200 as cb
  ev = w["ev"]
  th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides",
                                             "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf",
                                             "pixels_pe_opt_nside")})
  gcat = cb.pixelated_catalog(cb.dVdz_completeness(w["z_range"]), p_cat=w["p_cat"], P_compl=w["P_compl"])
  pop = cb.population(cb.cosmo.flrw(H0=70., Om0=0.25, z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=gcat)
  sel = cb.selection_function(cb.theta_inj_det(**w["inj"]), w["N_inj"], N_eff=5.)
  return cb.hyperlikelihood(th, w["zg"], pop, sel, kind_p_gw3d="approximate", kernel=kernel, binning=binning, num_bins=200,
                            cut_grid=2.0, pe_neff=2.0, fp_mode=fp_mode, distributed=distributed, presharded=True,
                            options=options)


# ------------------------------------------------------------------------------------------ CPU oracle
def _oracle_units(job):
  """Worker: log-likelihoods of a chunk of events for the given hyper-points (NumPy oracle)."""
  from oracle import chimera_oracle as orc
  ev, zg, cat, npx, hypers = job
  pop0 = orc.make_pop(orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                      orc.make_rate("madau_dickinson"), catalog=cat)
  opts = orc.make_opts("approximate", "gauss", None, 2.0, False, 200, 2.0)
  out = []
  with np.errstate(all="ignore"):
    for hl in hypers:
      pop = orc.pop_update(pop0, **hl)
      out.append(np.log(orc.numlike_evs(pop, ev, zg, opts, npx)))
  return np.array(out)


def _oracle_sel(job):
  from oracle import chimera_oracle as orc
  inj, N_inj, hl = job
  pop0 = orc.make_pop(orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                      orc.make_rate("madau_dickinson"))
  with np.errstate(all="ignore"):
    dN = orc.pop_rate_det_inj(orc.pop_update(pop0, **hl), inj) / inj["p_draw"]
  return np.nansum(dN), np.sum(dN ** 2)


def cpu_oracle_rate(w, n_events, n_hyper, procs):
  """Whole-job evals/s of the CPU oracle, measured on a bounded sample and composed as
  Nev / (Nev * t_unit + t_sel): t_unit from `n_events` events x `n_hyper` hyper-points, t_sel from
  the full injection set for `n_hyper` hyper-points (chunked over `procs` processes)."""
  import multiprocessing as mp
  ev = w["ev"]
  nev = ev["dL"].shape[0]
  n_events = min(n_events, nev)
  keys = ("m1det", "m2det", "dL", "pe_prior", "pixels_opt_nsides", "gw_loc2d_pdf")
  hypers = [dict(H0=float(w["hyper"]["H0"][i]), Om0=float(w["hyper"]["Om0"][i]))
            for i in np.linspace(0, len(w["hyper"]["H0"]) - 1, n_hyper).astype(int)]
  chunks = np.array_split(np.arange(n_events), procs)
  jobs = []
  for c in chunks:
    if c.size == 0:
      continue
    sl = slice(c[0], c[-1] + 1)
    cat = dict(p_cat=w["p_cat"][sl], P_compl=w["P_compl"][sl], z_range=w["z_range"])
    jobs.append(({k: ev[k][sl] for k in keys}, w["zg"][sl], cat, ev["neff_pixels"][sl], hypers))
  inj = w["inj"]
  ninj = inj["dL"].size
  ichunks = np.array_split(np.arange(ninj), procs)
  sjobs = [({k: v[c[0]:c[-1] + 1] for k, v in inj.items()}, w["N_inj"], hl) for hl in hypers for c in ichunks if c.size]
  ctx = mp.get_context("fork")
  with ctx.Pool(procs) as pool:
    pool.map(_oracle_units, [jobs[0][:4] + (hypers[:1],)])   # warm-up (imports, page-in)
    t0 = time.perf_counter()
    res = pool.map(_oracle_units, jobs)
    t_units = time.perf_counter() - t0
    t0 = time.perf_counter()
    pool.map(_oracle_sel, sjobs)
    t_sel = time.perf_counter() - t0
  t_unit = t_units / (n_events * n_hyper)
  t_sel_per_hyper = t_sel / n_hyper
  rate = nev / (nev * t_unit + t_sel_per_hyper)
  lle = np.concatenate(res, axis=1)
  return dict(rate=rate, t_unit=t_unit, t_sel=t_sel_per_hyper, lle=lle, hypers=hypers, n_events=n_events,
              seconds=t_units + t_sel)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
  Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

  def __init__(self, index):
    self.rows, self.proc, self.index = [], None, index

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits", "-lms", "100"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      f = [x.strip() for x in r.split(",")]
      if len(f) < 7:
        continue
      try:
        sm.append(float(f[0])); mx.append(float(f[1]))
      except ValueError:
        continue
      for nm, val in zip(names, f[3:7]):
        if val.lower().startswith("active"):
          reasons.add(nm)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return float(d.get("hbm_gbs", 6650.)), "MEASURED_PEAKS.json (driver-measured copy bandwidth)"
  return 6650., "fallback (B200_PROFILING.md)"


def traffic_from_profile(args):
  """DRAM bytes (read + written) of ONE numerator launch at the bench size, from the committed ncu --set full
  capture of this same command (profiles/traffic.json, written by scripts/ncu_traffic.py); None when the
  capture does not match the running configuration."""
  p = os.path.join(ROOT, "profiles", "traffic.json")
  if not os.path.exists(p):
    return None
  with open(p) as f:
    d = json.load(f)
  same = all(d.get(k) == getattr(args, k) for k in ("nev", "ns", "nz", "hyper_side")) and d.get("fp_mode") == args.fp_mode
  return d.get("dram_bytes_per_launch") if same else None


# ------------------------------------------------------------------------------------------ arms
def run_reference(args, rank, world):
  """CPU arm: the oracle restatement of the reference on all host cores, bounded sample."""
  if rank != 0:
    return
  w = build_workload(args, 0)
  procs = os.cpu_count() or 1
  vals = []
  for i in range(args.warmup + args.steps):
    r = cpu_oracle_rate(w, n_events=4 * procs, n_hyper=2, procs=procs)
    if i >= args.warmup:
      vals.append(r)
  rate = float(np.median([v["rate"] for v in vals]))
  sample = (f"{vals[0]['n_events']} events x 2 hyper-points (reweight+KDE+z-integral) + the full "
            f"{args.ninj} injections x 2 hyper-points per step, {procs} processes; value = Nev/(Nev*t_unit+t_sel)")
  line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": 1e3 * float(np.median([v["seconds"] for v in vals])),
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": WORKLOAD, "note": "CPU oracle (NumPy fp64 restatement of the reference; JAX unavailable offline)"},
          "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
          "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
  import torch
  import torch.distributed as dist
  import __graft_entry__ as ge
  ge.build()
  import chimera_b200 as cb
  from chimera_b200 import _lib
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  w = build_workload(args, rank)
  # cold start: handle creation + one-time upload of every input from host buffers + the first evaluation
  torch.cuda.synchronize()
  t_cold = time.perf_counter()
  like = build_likelihood(w, args.fp_mode, distributed=world > 1, options=parse_options(args.options))
  like(**w["hyper"])
  torch.cuda.synchronize()
  cold_s = time.perf_counter() - t_cold
  ev_keys = ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf",
             "pixels_pe_opt_nside")
  cold_bytes = int(sum(np.asarray(w["ev"][k]).nbytes for k in ev_keys) + w["zg"].nbytes + w["p_cat"].nbytes
                   + w["P_compl"].nbytes + sum(v.nbytes for v in w["inj"].values()))
  rows, _ = like.population.update(**w["hyper"]).hyper_rows()
  n_hyper = rows.shape[0]
  nev_local = w["ev"]["dL"].shape[0]
  d_rows = torch.from_numpy(rows).to(dev)
  d_part = torch.zeros((n_hyper, 3), dtype=torch.float64, device=dev)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing (value) -----------------------------------------------------
  for _ in range(args.warmup):
    like.partials_device(d_rows, d_part)
  barrier()
  launches0 = like.engine.launches
  sampler = ClockSampler(local_rank)
  sampler.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ktimes = []
  barrier()
  e0.record()
  for _ in range(args.steps):
    like.partials_device(d_rows, d_part)
    ktimes.append(None)
  e1.record()
  barrier()
  ms_total = e0.elapsed_time(e1)
  launches = like.engine.launches - launches0
  # per-kernel device times of the last step (CUDA events recorded on the launching stream)
  kt = like.engine.timings()
  clocks = sampler.stop()
  t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_step = float(t.item()) / args.steps
  units_global = nev_local * world * n_hyper
  value = units_global / (ms_step * 1e-3)

  # ---- end-to-end through the public API with host buffers (e2e) ----------------------------
  for _ in range(max(1, args.warmup // 2)):
    like(**w["hyper"])
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    out = like(**w["hyper"])
  barrier()
  e2e_s = (time.perf_counter() - t0) / args.steps
  t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  e2e_value = units_global / float(t.item())

  if rank != 0:
    return
  # ---- roofline of the dominant kernel (fused numerator: KDE pair sums) ----------------------
  G = args.nz // 2
  exps_per_launch = float(G) * args.ns * nev_local * n_hyper
  num_ms = float(kt["numerator_ms"])
  achieved = exps_per_launch / (num_ms * 1e-3) / 1e9
  peak = np.zeros(1)
  _lib.check(_lib.load().chb_mufu_peak(local_rank, 0.2, _lib.dptr(peak)))
  sm_max = clocks.get("sm_max_mhz") or 1965.0
  nominal = 148 * 16 * sm_max * 1e6 / 1e9
  hbm_peak, hbm_src = measured_peaks()
  sel_bytes = 32.0 * args.ninj * n_hyper
  sel_ms = float(kt["selection_ms"])
  line = {
    "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f32" if args.fp_mode == "fp32" else "f64", "data": "synthetic",
    "config": {"workload": WORKLOAD, "events_per_gpu": nev_local, "samples_per_event": args.ns,
               "injections_per_gpu": args.ninj, "n_hyper": n_hyper, "z_int_res": args.nz,
               "fp_mode": (args.fp_mode + ": fp32 reweighting + KDE pair sums (MUFU), fp64 tables/statistics/z-integral/reductions"
                           if args.fp_mode == "fp32" else "fp64 throughout"),
               "l2": "inputs larger than L2 (160 MB samples + 36 MB p_cat + 32 MB injections per GPU), no flush",
               "p_cat": "smooth synthetic catalogue term (synth.smooth_p_cat), same layout/sentinels",
               "step": "all hyper-points x (all events + all injections) + all-reduce of (n_hyper,3) partials"},
    "roofline": {"bound": "mufu_fp32_exp", "kernel": "numerator_f32_kernel<0>" if args.fp_mode == "fp32" else "numerator_kernel", "achieved": achieved, "peak": float(peak[0]) / 1e9,
                 "unit": "Gexp/s", "frac": achieved / (float(peak[0]) / 1e9), "traffic": traffic_from_profile(args),
                 "algorithmic": f"G*Ns exps per unit = {G}*{args.ns}; x {nev_local * n_hyper} units per launch",
                 "kernel_ms": num_ms, "peak_source": "ex2.approx micro-benchmark measured in this run (chb_mufu_peak)",
                 "nominal_peak": nominal},
    "roofline_selection": {"bound": "hbm", "kernel": "selection_kernel", "achieved": sel_bytes / (sel_ms * 1e-3) / 1e9,
                           "peak": hbm_peak, "unit": "GB/s", "frac": sel_bytes / (sel_ms * 1e-3) / 1e9 / hbm_peak,
                           "kernel_ms": sel_ms, "algorithmic": "32 B per (injection, hyper-point), unbatched",
                           "peak_source": hbm_src},
    "kernel_ms": {k: float(v) for k, v in kt.items()},
    "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rows.nbytes), "d2h_bytes_per_step": int(n_hyper * 3 * 8),
            "call": "hyperlikelihood.__call__(H0=array, Om0=array) -> chb_eval (host buffers) -> chb_finalize"},
    "e2e_cold": {"seconds": cold_s, "h2d_bytes": cold_bytes,
                 "what": "hyperlikelihood(...) construction (chb_create, chb_set_events/pixels/catalog/injections from host "
                         "buffers, host-side sort + packing) + the first __call__; paid once per run, as in the reference"},
    "gpu_launches": int(launches), "clocks": clocks,
  }
  if world == 1 and not args.no_cpu_baseline:
    procs = 1
    r = cpu_oracle_rate(w, n_events=240, n_hyper=8, procs=procs)
    line["cpu_baseline"] = {"value": r["rate"], "unit": UNIT, "cores": procs, "kind": "port",
                            "sample": f"{r['n_events']} events x 8 hyper-points + full {args.ninj} injections x 8 hyper-points "
                                      f"({r['seconds']:.1f} s of CPU work); value = Nev/(Nev*t_unit+t_sel)",
                            "t_unit_ms": 1e3 * r["t_unit"], "t_sel_s": r["t_sel"]}
    # cross-check of the timed configuration against the oracle on the sampled units
    idx = np.linspace(0, n_hyper - 1, 8).astype(int)
    lle = like.compute_all(**{k: v[idx] for k, v in w["hyper"].items()})[0][:, :r["n_events"]]
    ref = np.nan_to_num(r["lle"], nan=-np.inf)
    fin = np.isfinite(ref) & (np.abs(ref) < 1e300)
    line["parity_check"] = {"max_err_vs_oracle": float(np.max(np.abs(lle[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0))),
                            "metric": "|d log L| / max(|log L|, 1) per (event, hyper-point)", "units": int(fin.sum())}
    del like
    # the same workload with the reference's DEFAULT KDE options (Epanechnikov kernel, 200 bins),
    # for comparison only: 3 warm-up + 3 timed device-resident steps
    try:
      like_d = build_likelihood(w, args.fp_mode, False, kernel="epan", binning=True, options=parse_options(args.options))
      for _ in range(3):
        like_d.partials_device(d_rows, d_part)
      torch.cuda.synchronize()
      e0.record()
      for _ in range(3):
        like_d.partials_device(d_rows, d_part)
      e1.record()
      torch.cuda.synchronize()
      ms_d = e0.elapsed_time(e1) / 3
      line["reference_default_kde"] = {"value": units_global / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d,
                                       "config": "same workload, kernel='epan', binning=True, num_bins=200 (reference defaults)"}
      del like_d
    except Exception as exc:       # a side measurement must never cost the headline line
      line["reference_default_kde"] = {"error": repr(exc)}
  print(json.dumps(line), flush=True)


def main():
  args = parse()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if args.impl == "reference":
    run_reference(args, rank, world)
    return
  if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  try:
    run_ours(args, rank, world, local_rank)
  finally:
    if world > 1:
      import torch.distributed as dist
      dist.destroy_process_group()


if __name__ == "__main__":
  main()
