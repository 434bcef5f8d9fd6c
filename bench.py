#!/usr/bin/env python
"""Benchmark of the hierarchical-likelihood hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3] [--fp-mode fp32|fp64]
                    [--scaling strong|weak] [--sub C1,C2,C4,C3_fp64,C3_refdefault|none] [--options fused=0,...]

Headline workload (config.workload): BASELINE.json configs[2], the north-star target -- the O5-like mock:
1000 events x 5000 posterior samples, 10^6 detected injections, 256 hyper-points on a 16x16 H0 x Om0 grid,
galaxy-catalogue likelihood (pixelated, ~15 pixels/event, kind 'approximate'), 1-D Gaussian KDE without binning,
z_int_res = 300.  One step = one evaluation of the whole hyper-point batch: population reweighting -> KDE ->
z-integral for every (event, hyper-point) plus the injection-reweighted selection function for every hyper-point.
metric = hyper-point x event log-likelihood evaluations per second (whole job, all GPUs).

The same line carries `configs`: driver-run sub-records for the other BASELINE.json configurations (C1 spectral siren,
C2 galaxy catalogue 'marginalized', C4 full 3-D KDE at nside 64 with 10^7 galaxies, the fp64 mode of C3 and C3 with
the reference's default KDE options), each with ms_per_step, its roofline and a parity check against the CPU oracle.

Multi-GPU (torchrun, one rank per GPU): STRONG scaling by default -- ONE global C3 data set, events and injections
split contiguously over the ranks (the reference's rule, CHIMERA/parallel.py:68-73,94-99), every rank evaluates all
256 hyper-points on its shard, one NCCL all-reduce of the (256, 3) partials per step; the per-step fixed costs are
itemised in `fixed_costs`.  `weak` (every rank its own 1000 events) is kept as a sub-record, and at N = 8 the C5
configuration (10^4 events x 4096 walkers, mg_flrw + mass + rate hyper-parameters) runs with the 2-D hyper x event split.

`--impl reference` times the CPU restatement of the reference (the oracle; JAX is not installable offline, a real
install under baseline/_ref is preferred when importable) with all host cores on a bounded sample of the same workload.
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
Z_RANGE = np.array([0.073, 1.3])

# BASELINE.json configs.  ns/ninj are synthetic stand-ins where the notebooks' data is unavailable.
CONFIGS = {
  "C1": dict(workload="C1 spectral-siren 1-D H0 likelihood (examples/test1dspectral.ipynb cell 11): 300 events x 5000 samples, "
                      "epan kernel + 200 bins (reference defaults), z_int_res=500, 50 H0 points, 2e5 injections",
             nev=300, ns=5000, nz=500, ninj=200_000, kind=None, kernel="epan", binning=True, cosmo="flrw",
             hyper=lambda: dict(H0=np.linspace(50., 90., 50))),
  "C2": dict(workload="C2 galaxy-catalogue 1-D H0 likelihood (examples/test1dgalaxies.ipynb cell 13): 300 events x 5000 samples, "
                      "kind 'marginalized' (always Epanechnikov) + 200 bins, ~15 px/event, 1.6e6 galaxies, z_int_res=500, 100 H0 points",
             nev=300, ns=5000, nz=500, ninj=200_000, kind="marginalized", kernel="epan", binning=True, cosmo="flrw",
             ngal=1_600_000, nside_list=(512, 256, 128, 64, 32, 16, 8), npix=15,
             hyper=lambda: dict(H0=np.linspace(20., 200., 100))),
  "C3": dict(workload="C3 O5-like mock: 1000 events x 5000 samples, 1e6 injections, 256 hyper-points (16x16 H0 x Om0), "
                      "pixelated galaxy catalogue (~15 px/event, 1.6e6 galaxies, 'approximate'), Gaussian KDE unbinned, z_int_res=300",
             nev=1000, ns=5000, nz=300, ninj=1_000_000, kind="approximate", kernel="gauss", binning=False, cosmo="flrw",
             ngal=1_600_000, nside_list=(512, 256, 128, 64, 32, 16, 8), npix=15,
             hyper=lambda: (lambda H0, Om0: dict(H0=H0.ravel(), Om0=Om0.ravel()))(
               *np.meshgrid(np.linspace(55., 85., 16), np.linspace(0.15, 0.45, 16), indexing="ij"))),
  "C4": dict(workload="C4 full 3-D KDE galaxy-catalogue run: 500 events x 5000 samples, HEALPix nside=64 pixels, 1e7 galaxies, "
                      "dVdz incompleteness correction, Gaussian 3-D KDE, z_int_res=300, 64 H0 points",
             nev=500, ns=5000, nz=300, ninj=200_000, kind="full", kernel="gauss", binning=False, cosmo="flrw",
             ngal=10_000_000, nside_list=(64,), npix=15,
             hyper=lambda: dict(H0=np.linspace(40., 120., 64))),
  "C5": dict(workload="C5 modified propagation (Xi0, n) + mass + rate hyper-parameters: 1e4 events x 1000 samples, 4096 walkers "
                      "(14 parameters, N(fiducial, 5%)), 1e6 injections, Gaussian KDE unbinned, z_int_res=300, no catalogue",
             nev=10_000, ns=1000, nz=300, ninj=1_000_000, kind=None, kernel="gauss", binning=False, cosmo="mg_flrw",
             hyper=None),
}


def c5_walkers(n=4096, seed=3456):
  rng = np.random.default_rng(seed)
  fid = dict(H0=70., Xi0=1.0, n=1.9, alpha=3.4, beta=1.1, delta_m=4.8, m_low=5.1, m_high=87., mu_g=34., sigma_g=3.6,
             lambda_peak=0.039, gamma=2.7, kappa=3.0, zp=2.0)
  return {k: v * (1.0 + 0.05 * rng.standard_normal(n)) for k, v in fid.items()}


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--config", default="C3", choices=list(CONFIGS))
  ap.add_argument("--fp-mode", default="fp32", choices=["fp32", "fp64"])
  ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
  ap.add_argument("--sub", default="auto", help="comma list of sub-records (C1,C2,C4,C3_fp64,C3_refdefault,weak,C5), 'none', or 'auto'")
  ap.add_argument("--nev", type=int, default=0, help="override the number of events (profiling runs)")
  ap.add_argument("--ninj", type=int, default=0, help="override the number of injections (A/B runs of the numerator kernels)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--options", default="", help="per-handle tuning switches for A/B runs, e.g. 'fused=0,kde_win=0' (chb_set_option)")
  ap.add_argument("--hyper-groups", type=int, default=1)
  ap.add_argument("--kde", default="", help="profiling runs: 'epan-binned' evaluates the configuration with the reference's default "
                  "KDE options (kernel='epan', binning=True); needs --no-cpu-baseline")
  return ap.parse_args()


def parse_options(text):
  return {k: float(v) for k, v in (kv.split("=") for kv in text.split(",") if kv)}


def line_config(args, world):
  """The `config` object of the JSON line -- the SAME for both arms (`--impl ours` / `--impl reference`): it names the
  workload, its sizes and how the timed region treats the caches; arm-specific facts go elsewhere in the line."""
  from chimera_b200 import parallel
  c = CONFIGS[args.config]
  weak = world > 1 and args.scaling == "weak"
  nev = args.nev or c["nev"]
  ninj = args.ninj or c["ninj"]
  n_hyper = len(next(iter((c5_walkers() if c["hyper"] is None else c["hyper"]()).values())))
  E = world // max(1, args.hyper_groups)
  lo, hi = (0, nev) if weak else parallel.shard_bounds(nev, 0, max(1, E))
  return {"workload": c["workload"], "events_total": nev * (world if weak else 1), "events_per_gpu": hi - lo,
          "samples_per_event": c["ns"], "injections_total": int(ninj * (world if weak else 1)), "n_hyper": n_hyper,
          "z_int_res": c["nz"],
          "fp_mode": (args.fp_mode + ": fp32 reweighting + KDE pair sums (MUFU), fp64 tables/statistics/z-integral/reductions"
                      if args.fp_mode == "fp32" else "fp64 throughout") + " (GPU arm; the CPU arm is fp64 like the reference)",
          "l2": "inputs larger than L2 (120 MB packed samples + 32 MB injections + catalogue rows per GPU at N=1), no flush",
          "p_cat": "GPU arm: precompute_p_cat on the GPU from 1.6e6 synthetic galaxies (z_err 0.001(1+z)), spiky rows; CPU arm: "
                   "smooth synthetic rows of the same layout/sentinels (the CPU arm never touches the GPU)",
          "sharding": ("one global data set, contiguous event/injection shards (CHIMERA/parallel.py:68-73,94-99)" if not weak
                       else "every rank its own data set") + (f", hyper_groups={args.hyper_groups}" if args.hyper_groups > 1 else ""),
          "step": "all hyper-points x (all events + all injections) + all-reduce of (n_hyper,3) partials",
          "options": parse_options(args.options)}


# ------------------------------------------------------------------------------------------ workloads
def build_workload(name, seed_rank=0, nev=None, **overrides):
  """Synthetic inputs of configuration `name` (SURVEY.md section 8d recipes, chimera_b200/synth.py).  The
  pixelisation and the catalogue term come from the library's own setup kernels (pixelize_gw_catalog,
  pixelated_catalog.precompute_p_cat -- SURVEY 8f rows f1/f2) and their times are recorded."""
  from chimera_b200 import synth
  c = dict(CONFIGS[name], **overrides)        # overrides: smaller ninj / ns for tests and profiling runs
  nev = nev or c["nev"]
  t0 = time.perf_counter()
  sky = c["kind"] is not None
  ev = synth.make_events(nev, c["ns"], seed=1234 + 1000 * seed_rank, sky=sky)
  zg = synth.make_z_grids(ev["dL"], z_int_res=c["nz"], H0_prior=(20., 200.))
  inj, N_inj = synth.make_injections(c["ninj"], seed=5678 + 1000 * seed_rank, z_scale=0.3)
  w = dict(name=name, cfg=c, ev=ev, zg=zg, inj=inj, N_inj=N_inj, z_range=Z_RANGE, setup={})
  w["hyper"] = c5_walkers() if c["hyper"] is None else c["hyper"]()
  w["setup"]["synth_s"] = time.perf_counter() - t0
  if sky:
    import chimera_b200 as cb
    t0 = time.perf_counter()
    th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec")})
    th = cb.pixelize_gw_catalog(th, nside_list=list(c["nside_list"]), mean_npixels_event=c["npix"], sky_conf=0.9)
    w["setup"]["pixelize_gpu_s"] = time.perf_counter() - t0
    for k in ("opt_nsides", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside"):
      ev[k] = np.asarray(getattr(th, k))
    ev["neff_pixels"] = np.sum(ev["ra_pix"] != -100., axis=1).astype(np.int32)
    gal = synth.make_galaxies(c["ngal"], seed=9012 + seed_rank)
    t0 = time.perf_counter()
    fid = cb.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
    gcat = cb.pixelated_catalog(cb.dVdz_completeness(Z_RANGE), cosmo=fid, z_grids=zg, data_gw_pixelated=th,
                                data_gal=dict(ra=gal["ra"], dec=gal["dec"], z=gal["z"]), z_err=0.001)
    w["setup"]["precompute_p_cat_gpu_s"] = time.perf_counter() - t0
    w["setup"]["n_galaxies"] = int(c["ngal"])
    w["setup"]["p_cat_shape"] = list(gcat.p_cat.shape)
    w["p_cat"], w["P_compl"] = gcat.p_cat, np.asarray(gcat.P_compl).reshape(nev, c["nz"])
    w["th"] = th
  return w


def build_likelihood(w, fp_mode, distributed=False, presharded=False, options=None, hyper_groups=1, kernel=None, binning=None):
  import chimera_b200 as cb
  c, ev = w["cfg"], w["ev"]
  keys = ["m1det", "m2det", "dL", "pe_prior"]
  gcat = None
  if c["kind"] is not None:
    keys += ["ra", "dec", "opt_nsides", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside"]
    gcat = cb.pixelated_catalog(cb.dVdz_completeness(w["z_range"]), p_cat=w["p_cat"], P_compl=w["P_compl"])
  th = cb.theta_pe_det(**{k: ev[k] for k in keys})
  cosmo = getattr(cb.cosmo, c["cosmo"])(H0=70., Om0=0.25, z_max=5.)
  pop = cb.population(cosmo, cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=gcat)
  sel = cb.selection_function(cb.theta_inj_det(**w["inj"]), w["N_inj"], N_eff=5.)
  return cb.hyperlikelihood(th, w["zg"], pop, sel, kind_p_gw3d=c["kind"], kernel=kernel or c["kernel"],
                            binning=c["binning"] if binning is None else binning, num_bins=200, cut_grid=2.0, pe_neff=2.0,
                            fp_mode=fp_mode, distributed=distributed, presharded=presharded, options=options,
                            hyper_groups=hyper_groups)


# ------------------------------------------------------------------------------------------ CPU oracle
def _hyper_points(w, idx):
  return [{k: float(v[i]) for k, v in w["hyper"].items()} for i in idx]


def _oracle_pop(c, cat=None):
  from oracle import chimera_oracle as orc
  return orc.make_pop(orc.make_cosmo(c["cosmo"], H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                      orc.make_rate("madau_dickinson"), catalog=cat)


def _oracle_units(job):
  """Worker: log-likelihoods of a chunk of events for the given hyper-points (NumPy oracle)."""
  from oracle import chimera_oracle as orc
  c, ev, zg, cat, npx, hypers, kernel, binning = job
  pop0 = _oracle_pop(c, cat)
  opts = orc.make_opts(c["kind"], kernel, None, 2.0, binning, 200, 2.0)
  out = []
  with np.errstate(all="ignore"):
    for hl in hypers:
      out.append(np.log(orc.numlike_evs(orc.pop_update(pop0, **hl), ev, zg, opts, npx)))
  return np.array(out)


def _oracle_sel(job):
  from oracle import chimera_oracle as orc
  c, inj, hl = job
  with np.errstate(all="ignore"):
    dN = orc.pop_rate_det_inj(orc.pop_update(_oracle_pop(c), **hl), inj) / inj["p_draw"]
  return np.nansum(dN), np.sum(dN ** 2)


def _slim(c):
  return dict(kind=c["kind"], cosmo=c["cosmo"])      # picklable part of a CONFIGS entry (the workers need no more)


def _event_jobs(w, ev_idx, hypers, nchunks, kernel=None, binning=None):
  c, ev = w["cfg"], w["ev"]
  keys = ["m1det", "m2det", "dL", "pe_prior"]
  if c["kind"] is not None:
    keys += ["ra", "dec", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside"]
  jobs = []
  for ch in np.array_split(ev_idx, nchunks):
    if ch.size == 0:
      continue
    cat, npx = None, None
    if c["kind"] is not None:
      cat = dict(p_cat=w["p_cat"][ch], P_compl=w["P_compl"][ch][:, None, :], z_range=w["z_range"])
      npx = ev["neff_pixels"][ch]
    jobs.append((_slim(c), {k: ev[k][ch] for k in keys}, w["zg"][ch], cat, npx, hypers, kernel or c["kernel"],
                 c["binning"] if binning is None else binning))
  return jobs


def cpu_oracle_rate(w, n_events, n_hyper, procs):
  """Whole-job evals/s of the CPU oracle, measured on a bounded sample and composed as
  Nev / (Nev * t_unit + t_sel): t_unit from `n_events` events x `n_hyper` hyper-points, t_sel from the full injection
  set for `n_hyper` hyper-points (chunked over `procs` processes)."""
  import multiprocessing as mp
  nev_have = w["ev"]["dL"].shape[0]
  nev = w.get("nev_full", nev_have)            # the rate is composed for the FULL event count of the configuration
  n_events = min(n_events, nev_have)
  nh_all = len(next(iter(w["hyper"].values())))
  hypers = _hyper_points(w, np.linspace(0, nh_all - 1, n_hyper).astype(int))
  jobs = _event_jobs(w, np.arange(n_events), hypers, procs)
  inj = w["inj"]
  ichunks = np.array_split(np.arange(inj["dL"].size), procs)
  sjobs = [(_slim(w["cfg"]), {k: v[ch[0]:ch[-1] + 1] for k, v in inj.items()}, hl) for hl in hypers for ch in ichunks if ch.size]
  ctx = mp.get_context("fork")
  with ctx.Pool(procs) as pool:
    pool.map(_oracle_units, [jobs[0][:5] + (hypers[:1],) + jobs[0][6:]])   # warm-up (imports, page-in)
    t0 = time.perf_counter()
    res = pool.map(_oracle_units, jobs)
    t_units = time.perf_counter() - t0
    t0 = time.perf_counter()
    pool.map(_oracle_sel, sjobs)
    t_sel = time.perf_counter() - t0
  t_unit = t_units / (n_events * n_hyper)
  t_sel_per_hyper = t_sel / n_hyper
  rate = nev / (nev * t_unit + t_sel_per_hyper)
  return dict(rate=rate, t_unit=t_unit, t_sel=t_sel_per_hyper, lle=np.concatenate(res, axis=1), hypers=hypers,
              n_events=n_events, seconds=t_units + t_sel)


def parity_check(w, like, n_events, n_hyper, kernel=None, binning=None, procs=1):
  """Per-event log-likelihoods of the timed likelihood object against the CPU oracle on sampled units (events spread
  over the whole set, hyper-points spread over the batch); runs OUTSIDE every timed region."""
  import multiprocessing as mp
  nev = w["ev"]["dL"].shape[0]
  nh_all = len(next(iter(w["hyper"].values())))
  ev_idx = np.unique(np.linspace(0, nev - 1, min(n_events, nev)).astype(int))
  h_idx = np.unique(np.linspace(0, nh_all - 1, min(n_hyper, nh_all)).astype(int))
  hypers = _hyper_points(w, h_idx)
  jobs = _event_jobs(w, ev_idx, hypers, max(1, procs), kernel, binning)
  t0 = time.perf_counter()
  if procs > 1:
    with mp.get_context("fork").Pool(procs) as pool:
      res = pool.map(_oracle_units, jobs)
  else:
    res = [_oracle_units(j) for j in jobs]
  ref = np.nan_to_num(np.concatenate(res, axis=1), nan=-np.inf)
  lle = like.compute_all(**{k: v[h_idx] for k, v in w["hyper"].items()})[0][:, ev_idx]
  fin = np.isfinite(ref) & (np.abs(ref) < 1e300)
  same = bool(np.array_equal(fin, np.isfinite(lle) & (np.abs(lle) < 1e300)))
  err = float(np.max(np.abs(lle[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0))) if fin.any() else 0.0
  return {"max_err_vs_oracle": err, "metric": "|d log L| / max(|log L|, 1) per (event, hyper-point)", "units": int(fin.sum()),
          "non_finite_classes_match": same, "oracle_seconds": time.perf_counter() - t0}


# ------------------------------------------------------------------------------------------ clocks / peaks
class ClockSampler:
  Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

  def __init__(self, index):
    self.rows, self.proc, self.index = [], None, index
    self.t_begin = self.t_end = None

  def start(self):
    """Launched BEFORE the warm-up steps (nvidia-smi needs ~0.5 s to print its first row); every row is stamped on
    arrival and stop() keeps the rows that fall inside the timed region [mark_begin, mark_end]."""
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits", "-lms", "50"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.perf_counter(), line.strip()))

  def wait_first(self, timeout=3.0):
    t0 = time.perf_counter()
    while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
      time.sleep(0.02)

  def mark_begin(self):
    self.t_begin = time.perf_counter()

  def mark_end(self):
    self.t_end = time.perf_counter()

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def parse(rows):
      sm, mx, reasons = [], [], set()
      for _, r in rows:
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
      return sm, mx, reasons

    tb = self.t_begin if self.t_begin is not None else -1e300
    te = self.t_end if self.t_end is not None else 1e300
    inside = [r for r in self.rows if tb <= r[0] <= te + 0.05]
    window = "timed region"
    sm, mx, reasons = parse(inside)
    if not sm:
      # a timed region shorter than the sampling period: the rows of the warm-up steps right before it (same load)
      window = "warm-up + timed region (timed region shorter than one sampling period)"
      sm, mx, reasons = parse([r for r in self.rows if r[0] <= te + 0.05])
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return float(d.get("hbm_gbs", 6650.)), "MEASURED_PEAKS.json (driver-measured copy bandwidth)"
  return 6650., "fallback (B200_PROFILING.md)"


def executed_profile():
  """Executed work per unit of the dominant kernel from the committed `ncu --set full` capture of this same command
  (profiles/executed_r02.json, written by scripts/ncu_executed.py): warp instructions and MUFU (XU pipe) instructions
  per (event, hyper-point) unit, DRAM bytes per launch."""
  p = os.path.join(ROOT, "profiles", "executed_r02.json")
  if not os.path.exists(p):
    return None
  with open(p) as f:
    return json.load(f)


# ------------------------------------------------------------------------------------------ timing
def time_steps(like, w, steps, warmup, world, local_rank, sample_clocks=False):
  """Device-resident timing of `steps` evaluations of the whole hyper-point batch (inputs already in HBM): barrier +
  synchronize on both sides, CUDA events on the launching stream, max over ranks."""
  import torch
  import torch.distributed as dist
  dev = torch.device("cuda", local_rank)
  rows, _ = like.population.update(**w["hyper"]).hyper_rows()
  d_rows = torch.from_numpy(rows).to(dev)
  d_part = torch.zeros((rows.shape[0], 3), dtype=torch.float64, device=dev)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  sampler = ClockSampler(local_rank) if sample_clocks else None
  if sampler:
    sampler.start()
    like.partials_device(d_rows, d_part)      # (untimed: keeps the GPU under load while nvidia-smi starts)
    sampler.wait_first()
  for _ in range(warmup):
    like.partials_device(d_rows, d_part)
  barrier()
  launches0 = like.engine.launches
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  if sampler:
    sampler.mark_begin()
  e0.record()
  for _ in range(steps):
    like.partials_device(d_rows, d_part)
  e1.record()
  barrier()
  if sampler:
    sampler.mark_end()
  ms_total = e0.elapsed_time(e1)
  out = dict(launches=like.engine.launches - launches0, kernel_ms={k: float(v) for k, v in like.engine.timings().items()},
             clocks=sampler.stop() if sampler else None, rows=rows, n_hyper=rows.shape[0])
  t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  out["ms_per_step"] = float(t.item()) / steps
  return out


def time_e2e(like, w, steps, warmup, world, local_rank):
  """The same metric through the public API with HOST buffers: hyperlikelihood.__call__(H0=array, ...) -> H2D of the
  hyper-parameter rows, kernels, [all-reduce,] D2H of the partials, host epilogue -- wall clock, max over ranks."""
  import torch
  import torch.distributed as dist
  dev = torch.device("cuda", local_rank)
  for _ in range(max(1, warmup)):
    like(**w["hyper"])
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(steps):
    like(**w["hyper"])
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def kde_roofline(w, tm, fp_mode, mufu_peak, clocks, binning=None, nev_local=None):
  """Roofline of the numerator (KDE) kernel(s).  `achieved` = ALGORITHMIC exps / summed kernel time."""
  c = w["cfg"]
  nev = nev_local if nev_local is not None else w["ev"]["dL"].shape[0]
  G = c["nz"] // 2
  binned = c["binning"] if binning is None else binning
  n_data = 200 if binned else c["ns"]
  if c["kind"] == "marginalized":
    per_unit = float(np.mean(w["ev"]["neff_pixels"])) * G * (200 if binned else c["ns"])    # x pixels; bins per pixel when binned
    what = f"neff_pixels*G*{'B' if binned else 'Ns'} kernel evaluations per unit (Epanechnikov: no exp; counted as pair evaluations)"
  elif c["kind"] == "full":
    # n_z_eff = grid points inside [min z - c std, max z + c std] (likelihood.py:225), evaluated at the fiducial cosmology
    from chimera_b200 import synth
    z = synth._FlatLCDM().z_of_dL(w["ev"]["dL"])
    lo, hi = z.min(axis=1) - 2.0 * z.std(axis=1), z.max(axis=1) + 2.0 * z.std(axis=1)
    nzeff = np.sum((w["zg"] >= lo[:, None]) & (w["zg"] <= hi[:, None]), axis=1)
    per_unit = float(np.mean(w["ev"]["neff_pixels"] * nzeff)) * c["ns"]
    what = (f"neff_pixels*n_z_eff*Ns exps per unit (mean neff_pixels {float(np.mean(w['ev']['neff_pixels'])):.1f}, mean n_z_eff "
            f"{float(np.mean(nzeff)):.1f} of {c['nz']} at the fiducial cosmology)")
  else:
    per_unit = float(G) * n_data
    what = f"G*{'B' if binned else 'Ns'} = {G}*{n_data} {'exps' if c['kernel'] == 'gauss' else 'pair evaluations'} per unit"
  units = nev * tm["n_hyper"]
  ms = tm["kernel_ms"]["numerator_kernels_ms"] or tm["kernel_ms"]["numerator_ms"]
  achieved = per_unit * units / (ms * 1e-3) / 1e9
  r = {"bound": "mufu_fp32_exp" if c["kernel"] == "gauss" and c["kind"] != "marginalized" else "fp32_fma",
       "kernel": like_kernel_name(c, fp_mode), "achieved": achieved, "peak": mufu_peak / 1e9, "unit": "Gexp/s",
       "frac": achieved / (mufu_peak / 1e9), "traffic": None, "algorithmic": f"{what}; x {units} units per launch",
       "kernel_ms": ms, "peak_source": "ex2.approx micro-benchmark measured in this run (chb_mufu_peak)"}
  return r


def like_kernel_name(c, fp_mode):
  if fp_mode == "fp64":
    return "numerator_kernel (fp64)"
  if c["kind"] in (None, "approximate"):
    return "numerator_fused_kernel"
  return "numerator_f32_kernel<%d,1|2>" % (1 if c["kind"] == "marginalized" else 2)


# ------------------------------------------------------------------------------------------ arms
def reference_impl_available():
  """A real reference install (baseline/_ref with jax importable) is preferred when it ever appears."""
  ref = os.path.join(ROOT, "baseline", "_ref")
  if not os.path.isdir(ref):
    return False, "baseline/_ref absent (pip install of /root/reference fails offline: poetry-core and jax are not in the wheelhouse)"
  sys.path.insert(0, ref)
  try:
    import jax  # noqa: F401
    import CHIMERA  # noqa: F401
    return True, "baseline/_ref"
  except Exception as exc:
    return False, f"baseline/_ref present but not importable: {exc!r}"


def run_reference(args, rank, world):
  """CPU arm: the reference's path on all host cores, bounded sample.  The oracle restatement unless a real install is importable."""
  if rank != 0:
    return
  have_ref, why = reference_impl_available()
  w = _cpu_workload(args.config)
  procs = os.cpu_count() or 1
  vals = []
  for i in range(args.warmup + args.steps):
    r = cpu_oracle_rate(w, n_events=4 * procs, n_hyper=2, procs=procs)
    if i >= args.warmup:
      vals.append(r)
  rate = float(np.median([v["rate"] for v in vals]))
  c = CONFIGS[args.config]
  sample = (f"{vals[0]['n_events']} events x 2 hyper-points (reweight+KDE+z-integral) + the full "
            f"{c['ninj']} injections x 2 hyper-points per step, {procs} processes; value = Nev/(Nev*t_unit+t_sel)")
  line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": 1e3 * float(np.median([v["seconds"] for v in vals])),
          "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": line_config(args, world),
          "note": "CPU oracle (NumPy fp64 restatement of the reference's path); real reference: " + why,
          "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
          "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  print(json.dumps(line), flush=True)


def _cpu_workload(name):
  """The C3 workload for the CPU arm without touching the GPU: host pixelisation (synth.pixelize) of the sampled
  events only and the smooth catalogue term of round 1 (same shapes and sentinels)."""
  from chimera_b200 import synth
  c = CONFIGS[name]
  procs = os.cpu_count() or 1
  nev = min(c["nev"], max(8 * procs, 64))
  sky = c["kind"] is not None
  ev = synth.make_events(nev, c["ns"], seed=1234, sky=sky)
  zg = synth.make_z_grids(ev["dL"], z_int_res=c["nz"], H0_prior=(20., 200.))
  inj, N_inj = synth.make_injections(c["ninj"], seed=5678, z_scale=0.3)
  w = dict(name=name, cfg=c, ev=ev, zg=zg, inj=inj, N_inj=N_inj, z_range=Z_RANGE, setup={})
  w["hyper"] = c5_walkers() if c["hyper"] is None else c["hyper"]()
  if sky:
    ev = synth.pixelize(ev, nside_list=c["nside_list"], mean_npixels_event=c["npix"], sky_conf=0.9)
    w["ev"] = ev
    p_cat, P_compl = synth.smooth_p_cat(ev, zg, seed=9012)
    w["p_cat"], w["P_compl"] = p_cat, np.asarray(P_compl).reshape(nev, c["nz"])
  # the rate formula composes with the FULL event count of the configuration
  w["nev_full"] = c["nev"]
  return w


def sub_record(name, args, fp_mode, local_rank, mufu_peak, kernel=None, binning=None, base=None, steps=3, warmup=3,
               parity=(8, 2), options=None):
  """One extra configuration on ONE GPU: build (or reuse `base`), time, roofline, parity."""
  import torch
  w = base or build_workload(name)
  t0 = time.perf_counter()
  like = build_likelihood(w, fp_mode, options=options, kernel=kernel, binning=binning)
  like(**w["hyper"])
  torch.cuda.synchronize()
  cold = time.perf_counter() - t0
  tm = time_steps(like, w, steps, warmup, 1, local_rank)
  c = w["cfg"]
  units = w["ev"]["dL"].shape[0] * tm["n_hyper"]
  rec = {"workload": c["workload"], "fp_mode": fp_mode, "kernel": kernel or c["kernel"],
         "binning": c["binning"] if binning is None else binning, "n_hyper": tm["n_hyper"], "units_per_step": units,
         "steps": steps, "warmup": warmup, "ms_per_step": tm["ms_per_step"], "value": units / (tm["ms_per_step"] * 1e-3),
         "unit": UNIT, "kernel_ms": tm["kernel_ms"], "gpu_launches": tm["launches"],
         "roofline": kde_roofline(w, tm, fp_mode, mufu_peak, None, binning=binning), "setup": w["setup"], "cold_start_s": cold}
  if parity:
    rec["parity_check"] = parity_check(w, like, parity[0], parity[1], kernel=kernel, binning=binning,
                                       procs=min(os.cpu_count() or 1, 8))
  del like
  torch.cuda.empty_cache()
  return rec, w


def run_ours(args, rank, world, local_rank):
  import torch
  import torch.distributed as dist
  import __graft_entry__ as ge
  ge.build()
  from chimera_b200 import _lib
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  opts = parse_options(args.options)
  name = args.config
  c = CONFIGS[name]
  weak = world > 1 and args.scaling == "weak"
  # strong scaling: ONE global data set (every rank generates the same seeded inputs and keeps its contiguous shard);
  # weak scaling: every rank its own data set of the full size
  w = build_workload(name, seed_rank=rank if weak else 0, nev=args.nev or None, **({"ninj": args.ninj} if args.ninj else {}))
  torch.cuda.synchronize()
  t_cold = time.perf_counter()
  kov = {"epan-binned": dict(kernel="epan", binning=True), "epan": dict(kernel="epan", binning=False)}.get(args.kde, {})
  like = build_likelihood(w, args.fp_mode, distributed=world > 1, presharded=weak, options=opts, hyper_groups=args.hyper_groups, **kov)
  like(**w["hyper"])
  torch.cuda.synchronize()
  cold_s = time.perf_counter() - t_cold
  nev_glob = w["ev"]["dL"].shape[0] * (world if weak else 1)
  nev_local = like.engine.Nev
  tm = time_steps(like, w, args.steps, args.warmup, world, local_rank, sample_clocks=True)
  n_hyper = tm["n_hyper"]
  units_global = nev_glob * n_hyper
  value = units_global / (tm["ms_per_step"] * 1e-3)
  e2e_s = time_e2e(like, w, args.steps, max(1, args.warmup // 2), world, local_rank)
  e2e_value = units_global / e2e_s

  # fixed per-step costs that do not shrink with the shard (strong scaling): measured on every rank, reported by rank 0
  fixed = None
  if world > 1:
    d_part = torch.zeros((n_hyper, 3), dtype=torch.float64, device=dev)
    for _ in range(5):
      dist.all_reduce(d_part)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      dist.all_reduce(d_part)
    e1.record()
    torch.cuda.synchronize()
    rows = tm["rows"]
    t0 = time.perf_counter()
    for _ in range(20):
      like.engine.finalize(rows, np.zeros((n_hyper, 3)), nev_glob)
    fin_ms = (time.perf_counter() - t0) / 20 * 1e3
    km = tm["kernel_ms"]
    fixed = {"build_tables_ms": km["tables_ms"], "reduce_ms": km["reduce_ms"], "allreduce_ms": e0.elapsed_time(e1) / 20,
             "host_finalize_ms": fin_ms, "launches_per_step": tm["launches"] / args.steps,
             "shard_scaled": {"zgrid_terms_ms": km["zgrid_terms_ms"], "numerator_kernels_ms": km["numerator_kernels_ms"],
                              "selection_ms": km["selection_ms"]},
             "events_local": nev_local, "note": "tables / reduce / all-reduce / finalize / launch overhead are paid per step on "
             "every rank whatever the shard size; the shard_scaled kernels shrink with 1/N"}

  if rank != 0:
    # the sub-records of multi-GPU runs that need every rank
    del like
    torch.cuda.empty_cache()
    if world > 1:
      _multi_gpu_subs(args, rank, world, local_rank, None, None)
    return
  # ---- roofline of the dominant kernel ------------------------------------------------------------------
  peak = np.zeros(1)
  _lib.check(_lib.load().chb_mufu_peak(local_rank, 0.2, _lib.dptr(peak)))
  clocks = tm["clocks"]
  sm_max = clocks.get("sm_max_mhz") or 1965.0
  sm_mhz = clocks.get("sm_mhz") or sm_max
  hbm_peak, hbm_src = measured_peaks()
  km = tm["kernel_ms"]
  roof = kde_roofline(w, tm, args.fp_mode, float(peak[0]), clocks, nev_local=nev_local, binning=kov.get("binning"))
  roof["nominal_peak"] = 148 * 16 * sm_max * 1e6 / 1e9
  prof = executed_profile()
  roof_exec = None
  same_cmd = prof and prof.get("config") == name and prof.get("fp_mode") == args.fp_mode and not args.options and not args.kde
  if same_cmd:
    # executed work: warp instructions of the capture per unit x units of this run / measured kernel time, against the
    # issue rate 148 SMs x 4 schedulers x f_SM (the clock sampled during THIS run); the MUFU fraction likewise
    units_local = nev_local * n_hyper
    ms = km["numerator_kernels_ms"]
    issue_peak = 148 * 4 * sm_mhz * 1e6
    inst_rate = prof["warp_inst_per_unit"] * units_local / (ms * 1e-3)
    xu_rate = prof["xu_warp_inst_per_unit"] * units_local * 32 / (ms * 1e-3)        # MUFU ops (thread level)
    roof_exec = {"kernel": prof["kernel"], "bound": "issue_slots", "achieved": inst_rate / 1e9, "peak": issue_peak / 1e9,
                 "unit": "Gwarp-inst/s", "frac": inst_rate / issue_peak, "xu_frac": xu_rate / float(peak[0]),
                 "warp_inst_per_unit": prof["warp_inst_per_unit"], "xu_warp_inst_per_unit": prof["xu_warp_inst_per_unit"],
                 "kernel_ms": ms, "sm_mhz": sm_mhz,
                 "source": "profiles/executed_r02.json (ncu --set full of this command: smsp__inst_executed.sum, "
                           "smsp__inst_executed_pipe_xu.sum per launch / units per launch)"}
    roof["traffic"] = prof.get("dram_bytes_per_launch")
  sel_bytes = 32.0 * (w["inj"]["dL"].size / (world if not weak else 1)) * n_hyper
  sel_ms = km["selection_ms"]
  ev_keys = ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf",
             "pixels_pe_opt_nside")
  cold_bytes = int(sum(np.asarray(w["ev"][k]).nbytes for k in ev_keys if k in w["ev"]) + w["zg"].nbytes
                   + (w["p_cat"].nbytes + w["P_compl"].nbytes if "p_cat" in w else 0) + sum(v.nbytes for v in w["inj"].values()))
  line = {
    "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": tm["ms_per_step"], "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
    "dtype": "f32" if args.fp_mode == "fp32" else "f64", "data": "synthetic",
    "config": line_config(args, world),
    "roofline": roof,
    "roofline_selection": {"bound": "hbm", "kernel": "selection_f32_kernel" if args.fp_mode == "fp32" else "selection_kernel",
                           "achieved": sel_bytes / (sel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": sel_bytes / (sel_ms * 1e-3) / 1e9 / hbm_peak, "kernel_ms": sel_ms,
                           "algorithmic": "32 B per (injection, hyper-point), unbatched", "peak_source": hbm_src},
    "kernel_ms": km,
    "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tm["rows"].nbytes), "d2h_bytes_per_step": int(n_hyper * 3 * 8),
            "call": "hyperlikelihood.__call__(H0=array, Om0=array) -> chb_eval (host buffers) -> chb_finalize"},
    "e2e_cold": {"seconds": cold_s, "h2d_bytes": cold_bytes, "setup": w["setup"],
                 "what": "hyperlikelihood(...) construction (chb_create, chb_set_events/pixels/catalog/injections from host "
                         "buffers, host-side sort + packing) + the first __call__; `setup` = the GPU pixelisation and the "
                         "catalogue sum (precompute_p_cat) that produce its inputs; paid once per run, as in the reference"},
    "gpu_launches": int(tm["launches"]), "clocks": clocks,
  }
  if roof_exec:
    line["roofline_executed"] = roof_exec
  if fixed:
    line["fixed_costs"] = fixed
  subs = args.sub
  if subs == "auto":
    subs = "C1,C2,C4,C3_fp64,C3_refdefault" if (world == 1 and name == "C3" and not args.nev and not args.ninj) else ("weak,C5" if world > 1 and name == "C3" else "none")
  subs = [s for s in subs.split(",") if s and s != "none"]
  if world == 1 and not args.no_cpu_baseline:
    procs = 1
    r = cpu_oracle_rate(w, n_events=240, n_hyper=8, procs=procs)
    line["cpu_baseline"] = {"value": r["rate"], "unit": UNIT, "cores": procs, "kind": "port",
                            "sample": f"{r['n_events']} events x 8 hyper-points + full {c['ninj']} injections x 8 hyper-points "
                                      f"({r['seconds']:.1f} s of CPU work); value = Nev/(Nev*t_unit+t_sel)",
                            "t_unit_ms": 1e3 * r["t_unit"], "t_sel_s": r["t_sel"]}
    # cross-check of the timed configuration against the oracle on the sampled units
    idx = np.linspace(0, n_hyper - 1, 8).astype(int)
    lle = like.compute_all(**{k: v[idx] for k, v in w["hyper"].items()})[0][:, :r["n_events"]]
    ref = np.nan_to_num(r["lle"], nan=-np.inf)
    fin = np.isfinite(ref) & (np.abs(ref) < 1e300)
    line["parity_check"] = {"max_err_vs_oracle": float(np.max(np.abs(lle[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0))),
                            "metric": "|d log L| / max(|log L|, 1) per (event, hyper-point)", "units": int(fin.sum())}
  elif world == 1:
    line["parity_check"] = parity_check(w, like, 16, 4, procs=min(os.cpu_count() or 1, 8), **kov)
  lle32_all = None
  if world == 1 and "C3_fp64" in subs and args.fp_mode == "fp32":
    lle32_all = like.compute_all(**w["hyper"])[0]
  del like
  torch.cuda.empty_cache()
  configs = {}
  for s in subs:
    try:
      if s in ("weak", "C5"):
        continue
      if s == "C3_fp64":
        configs[s], _ = sub_record("C3", args, "fp64", local_rank, float(peak[0]), base=w, parity=(8, 2))
        if lle32_all is not None:
          # every unit of the headline configuration: the timed fp32 mode against the fp64 mode (itself held to the oracle)
          l64 = build_likelihood(w, "fp64").compute_all(**w["hyper"])[0]
          fin = np.isfinite(l64) & (np.abs(l64) < 1e300)
          configs[s]["fp32_vs_fp64_all_units"] = {
            "units": int(l64.size), "finite_units": int(fin.sum()),
            "non_finite_classes_match": bool(np.array_equal(fin, np.isfinite(lle32_all) & (np.abs(lle32_all) < 1e300))),
            "max_err": float(np.max(np.abs(lle32_all[fin] - l64[fin]) / np.maximum(np.abs(l64[fin]), 1.0))),
            "metric": "|d log L| / max(|log L|, 1) per (event, hyper-point), fp32 mode vs fp64 mode"}
      elif s == "C3_refdefault":
        configs[s], _ = sub_record("C3", args, args.fp_mode, local_rank, float(peak[0]), kernel="epan", binning=True, base=w,
                                   parity=(8, 2), options=opts)
      else:
        configs[s], _ = sub_record(s, args, args.fp_mode, local_rank, float(peak[0]), parity=(8, 2), options=opts)
    except Exception as exc:       # a side measurement must never cost the headline line
      configs[s] = {"error": repr(exc)}
  if world > 1:
    _multi_gpu_subs(args, rank, world, local_rank, configs, float(peak[0]), subs)
  if configs:
    line["configs"] = configs
  print(json.dumps(line), flush=True)


def _multi_gpu_subs(args, rank, world, local_rank, configs, mufu_peak, subs=None):
  """Sub-records that need every rank: weak scaling of C3 and (N = 8) the C5 walker batch with the 2-D split.  Every rank
  enters with the same argument list (derived from args), rank 0 fills `configs`."""
  import torch
  subs = args.sub
  if subs == "auto":
    subs = "weak,C5" if args.config == "C3" else "none"
  subs = [s for s in subs.split(",") if s and s != "none"]
  for s in subs:
    try:
      if s == "weak":
        w = build_workload("C3", seed_rank=rank)
        like = build_likelihood(w, args.fp_mode, distributed=True, presharded=True, options=parse_options(args.options))
        tm = time_steps(like, w, 3, 3, world, local_rank)
        units = w["ev"]["dL"].shape[0] * world * tm["n_hyper"]
        rec = {"scaling": "weak", "events_per_gpu": w["ev"]["dL"].shape[0], "ms_per_step": tm["ms_per_step"],
               "value": units / (tm["ms_per_step"] * 1e-3), "unit": UNIT, "kernel_ms": tm["kernel_ms"]}
      elif s == "C5":
        if world < 8 and args.sub == "auto":
          continue
        w = build_workload("C5", seed_rank=0)
        rec = {"workload": CONFIGS["C5"]["workload"], "unit": UNIT}
        for k in (1, 2):
          like = build_likelihood(w, args.fp_mode, distributed=True, presharded=False, options=parse_options(args.options),
                                  hyper_groups=k)
          tm = time_steps(like, w, 2, 3, world, local_rank)
          units = w["ev"]["dL"].shape[0] * tm["n_hyper"]
          rec[f"hyper_groups_{k}"] = {"ms_per_step": tm["ms_per_step"], "value": units / (tm["ms_per_step"] * 1e-3),
                                      "kernel_ms": tm["kernel_ms"], "events_local": like.engine.Nev,
                                      "hyper_points_local": tm["n_hyper"] // k}
          del like
          torch.cuda.empty_cache()
      else:
        continue
      if configs is not None:
        configs[s] = rec
    except Exception as exc:
      if configs is not None:
        configs[s] = {"error": repr(exc)}


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
