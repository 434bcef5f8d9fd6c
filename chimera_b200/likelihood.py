"""`hyperlikelihood` with the constructor, attributes and methods of CHIMERA/likelihood.py:13-338.

Differences from the reference are confined to how work is fed to the device:
  * keyword hyper-parameters may be scalars (one hyper-point, as in the reference) or equal-length
    1-D arrays (a batch, e.g. the dict `emcee_utils.generate_dict` builds from a walker matrix);
    batched calls return arrays with a leading `n_hyper` axis;
  * `fp_mode` ('fp64' | 'fp32') selects the arithmetic of the KDE pair sums;
  * with `distributed=True` under torch.distributed the events and injections are sharded
    contiguously over the ranks and the per-rank partials are summed with one all-reduce.
The CUDA library is mandatory: there is no CPU path."""
from numbers import Number
import numpy as np
from . import _lib, parallel
from .engine import Engine
from .population._base import model_config
from .catalog.catalog import empty_catalog


def _bw_fields(bw_method):
  if bw_method is None or bw_method == "scott":
    return 0, 0.0
  if bw_method == "silverman":
    return 1, 0.0
  if isinstance(bw_method, Number) and not isinstance(bw_method, bool):
    return 2, float(bw_method)
  raise ValueError("bw_method should be 'scott', 'silverman', or a scalar")


class hyperlikelihood(object):
  def __init__(self, theta_gw_det, z_grids, population, selection_function=None, kind_p_gw3d=None,
               kernel='epan', bw_method=None, cut_grid=2.0, binning=True, num_bins=200, pe_neff=2.0,
               fp_mode='fp64', device=None, distributed=False, process_group=None, presharded=False,
               options=None, hyper_groups=1):
    self.theta_gw_det = theta_gw_det
    self.population = population
    self.z_grids = np.asarray(z_grids, dtype=np.float64)
    self.selection_function = selection_function
    self.kind_p_gw3d = kind_p_gw3d
    self.kernel = kernel
    self.bw_method = bw_method
    self.cut_grid = cut_grid
    self.binning = binning
    self.num_bins = num_bins
    self.pe_neff = pe_neff
    self.fp_mode = fp_mode

    self.pixelated = theta_gw_det.pixels_opt_nsides is not None
    self.nevents = len(theta_gw_det.dL)
    self.z_int_res = self.z_grids.shape[1]
    gal_cat = population.gal_cat
    if self.pixelated:
      assert self.kind_p_gw3d in ['approximate', 'marginalized', 'full'], \
        "`kind_p_gw3d` must be one of `approximate`, `marginalized`, or `full`"
      self.max_npixels = gal_cat.max_npixels
      self.neff_pixels = gal_cat.neff_pixels
      self.p_gw3d = {'approximate': self.p_gw3dapprox, 'marginalized': self.p_gw3dmarg,
                     'full': self.p_gw3dfull}[self.kind_p_gw3d]
      self.compute_numlike_evs = self._compute_numlike_evs_pixelated
    else:
      self.compute_numlike_evs = self._compute_numlike_evs_no_pixels
    if kernel not in _lib.KERNEL_IDS:
      raise ValueError("kernel must be 'epan' or 'gauss'")
    if fp_mode not in _lib.FP_IDS:
      raise ValueError("fp_mode must be 'fp64' or 'fp32'")
    bw_id, bw_val = _bw_fields(bw_method)

    # ---- sharding ---------------------------------------------------
    # `presharded=True`: the arrays passed in are already this rank's shard (weak-scaling runs);
    # otherwise every rank holds the global arrays and keeps its contiguous chunk.
    self.rank, self.world = parallel.dist_info(process_group) if distributed else (0, 1)
    self.group = process_group
    # 2-D layout (hyper_groups > 1): ranks in one hyper group split the events/injections, the groups split the
    # hyper-points of every batch (the reference's 'both' scheme, CHIMERA/parallel.py:132-229) -- for walker batches
    # so large that rebuilding every table and every z-grid term on every rank would dominate (C5: 4096 points).
    self.hyper_groups = int(hyper_groups)
    self._eshard, self._nshards, self._hgroup = parallel.grid_coords(self.rank, self.world, self.hyper_groups)
    self.presharded = bool(presharded) and self.world > 1
    if self.presharded:
      if self.hyper_groups != 1:
        raise ValueError("presharded inputs and hyper_groups > 1 cannot be combined")
      self._ev_counts = parallel.allgather_counts(self.nevents, process_group)
      lo, hi = 0, self.nevents
      self.nevents = int(sum(self._ev_counts))
    else:
      self._ev_counts = [parallel.shard_bounds(self.nevents, r, self._nshards) for r in range(self._nshards)]
      self._ev_counts = [b - a for a, b in self._ev_counts]
      lo, hi = parallel.shard_bounds(self.nevents, self._eshard, self._nshards)
    self._ev_slice = slice(lo, hi)
    if device is None:
      device = 0
      if self.world > 1:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0

    sel = selection_function
    kind = self.kind_p_gw3d if self.pixelated else None
    has_cat = self.pixelated and not isinstance(gal_cat, empty_catalog)
    extra = dict(fp_mode=_lib.FP_IDS[fp_mode], kind_p_gw=_lib.KIND_IDS[kind], kernel=_lib.KERNEL_IDS[kernel],
                 bw_method=bw_id, bw_value=bw_val, use_cut_grid=0 if cut_grid is None else 1,
                 cut_grid=0.0 if cut_grid is None else float(cut_grid), binning=int(bool(binning)),
                 num_bins=int(num_bins), pe_neff=float(pe_neff), scale_free=int(bool(population.scale_free)),
                 Tobs=float(population.Tobs), catalog_kind=1 if has_cat else 0)
    if has_cat:
      zr = np.asarray(gal_cat.completeness.z_range, dtype=np.float64)
      extra.update(compl_z_lo=float(zr[0]), compl_z_hi=float(zr[1]))
    if sel is not None:
      extra.update(sel._config_fields())
    self.cfg = model_config(population.cosmo, population.mass, population.rate, device=device, **extra)
    self.engine = Engine(self.cfg)
    for name, value in (options or {}).items():     # per-handle tuning switches (chb_set_option); results do not depend on them
      self.engine.set_option(name, value)

    t = theta_gw_det
    s = self._ev_slice
    take = lambda a: None if a is None else np.asarray(a)[s]
    if hi > lo:
      self.engine.set_events(take(t.m1det), take(t.m2det), take(t.dL), take(t.pe_prior), self.z_grids[s],
                             take(t.ra), take(t.dec))
      if self.pixelated:
        self.engine.set_pixels(take(t.pixels_opt_nsides), take(t.pixels_pe_opt_nside), take(t.ra_pix),
                               take(t.dec_pix), take(t.gw_loc2d_pdf))
        if has_cat:
          self.engine.set_catalog(np.asarray(gal_cat.p_cat)[s], np.asarray(gal_cat.P_compl)[s])
    if sel is not None:
      ti = sel.theta_inj_det
      arrs = [np.ravel(np.asarray(x, dtype=np.float64)) for x in (ti.m1det, ti.m2det, ti.dL, ti.p_draw)]
      ilo, ihi = (0, arrs[0].size) if self.presharded else parallel.shard_bounds(arrs[0].size, self._eshard, self._nshards)
      if ihi > ilo:
        self.engine.set_injections(*[a[ilo:ihi] for a in arrs])
    self._has_work = (hi > lo) or (sel is not None and ihi > ilo)      # an empty shard contributes zero partials

  # ---- evaluation core ---------------------------------------------------------------------
  def _evaluate(self, pop_lambdas, want_events=False, want_pgw=False):
    rows, batched = pop_lambdas.hyper_rows()
    n = rows.shape[0]
    if self.world == 1:
      lle, part, pgw = self.engine.eval(rows, want_events=want_events, want_pgw=want_pgw)
    elif not want_events and parallel.backend(self.group) == "nccl":
      # NCCL: the partials stay on the device from the kernels through the all-reduce; one D2H copy of (n, 3)
      import torch
      dev = torch.device("cuda", self.cfg.device)
      d_rows = torch.from_numpy(np.ascontiguousarray(rows)).to(dev)
      d_part = torch.zeros((n, 3), dtype=torch.float64, device=dev)
      with torch.cuda.device(dev):
        self.partials_device(d_rows, d_part)
        part = d_part.cpu().numpy()
      lle, pgw = None, None
    else:
      # this rank's hyper-points (all of them unless hyper_groups > 1) on this rank's events and injections
      hlo, hhi = parallel.shard_bounds(n, self._hgroup, self.hyper_groups)
      part = np.zeros((n, 3))
      lle, pgw = None, None
      nev_loc = self.engine.Nev
      if want_events:
        lle = np.zeros((n, nev_loc))
      if self._has_work and hhi > hlo:
        l, p, _ = self.engine.eval(rows[hlo:hhi], want_events=want_events, want_pgw=False)
        part[hlo:hhi] = p
        if want_events and l is not None:
          lle[hlo:hhi] = l
      part = parallel.allreduce_partials(part, self.group)
      if want_events:
        # per-event values: gather the event shards; with hyper groups the zero-filled blocks of the other groups add up
        lle = parallel.allgather_events(lle, self._ev_counts * self.hyper_groups if self.hyper_groups > 1 else self._ev_counts,
                                        self.group)
        if self.hyper_groups > 1:
          E = sum(self._ev_counts)
          lle = sum(lle[:, g * E:(g + 1) * E] for g in range(self.hyper_groups))
    fin = self.engine.finalize(rows, part, self.nevents)
    return rows, batched, lle, part, pgw, fin

  def partials_device(self, d_rows, d_partials, stream=None):
    """Device-resident entry (torch f64 CUDA tensors): rows (n, CHB_NPAR) -> partials (n, 3), summed over ranks with
    one NCCL all-reduce when distributed (no host round trip).  Asynchronous on torch's current stream."""
    import torch
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    st = st or 1      # 0 is the legacy default stream: name it explicitly (cudaStreamLegacy), NULL means 'handle stream'
    if self.world == 1:
      self.engine.eval_device(d_rows, d_partials, None, st)
      return d_partials
    n = d_rows.shape[0]
    hlo, hhi = parallel.shard_bounds(n, self._hgroup, self.hyper_groups)
    if hlo > 0 or hhi < n or not self._has_work:
      d_partials.zero_()
    if self._has_work and hhi > hlo:
      self.engine.eval_device(d_rows[hlo:hhi], d_partials[hlo:hhi], None, st)
    parallel.allreduce_partials(d_partials, self.group)
    return d_partials

  @staticmethod
  def _shape(x, batched):
    return x if batched else x[0]

  # ---- p_gw entry points (likelihood.py:105-260) -------------------------------------------
  def _p_gw(self, pop_lambdas):
    if self.world > 1:
      raise RuntimeError("p_gw inspection is a single-process debugging entry point")
    rows, batched, _, _, pgw, _ = self._evaluate(pop_lambdas, want_pgw=True)
    return self._shape(pgw, batched)

  def p_gw1d(self, pop_lambdas):
    if self.pixelated and self.kind_p_gw3d != 'approximate':
      raise ValueError("p_gw1d is available for non-pixelated data or kind_p_gw3d='approximate'")
    out = self._p_gw(pop_lambdas)
    if self.pixelated:   # recover p_gw1d from p_gw1d * gw_loc2d_pdf of the first pixel
      pdf0 = np.asarray(self.theta_gw_det.gw_loc2d_pdf)[:, 0]
      return out[..., 0, :] / pdf0[:, None]
    return out

  def p_gw3dapprox(self, pop_lambdas):
    return self._p_gw(pop_lambdas)

  def p_gw3dmarg(self, pop_lambdas):
    return self._p_gw(pop_lambdas)

  def p_gw3dfull(self, pop_lambdas):
    return self._p_gw(pop_lambdas)

  # ---- numerator (likelihood.py:266-301) -----------------------------------------------------
  def _numlike(self, pop_lambdas):
    if self.world > 1:
      raise RuntimeError("compute_numlike_evs is a single-process debugging entry point")
    rows, batched, _, _, _, _ = self._evaluate(pop_lambdas)
    return self._shape(self.engine.numlike_evs(rows.shape[0]), batched)

  def _compute_numlike_evs_pixelated(self, pop_lambdas):
    return self._numlike(pop_lambdas)

  def _compute_numlike_evs_no_pixels(self, pop_lambdas):
    return self._numlike(pop_lambdas)

  def compute_log_likenum(self, pop_lambdas):
    _, batched, _, _, _, fin = self._evaluate(pop_lambdas)
    return self._shape(fin["log_like_num"], batched)

  # ---- hyper-likelihood (likelihood.py:307-338) ------------------------------------------------
  def compute_log_hyperlike(self, **hyper_lambdas):
    pop_lambdas = self.population.update(**hyper_lambdas)
    if self.selection_function is None:
      raise AttributeError("'NoneType' object has no attribute 'N_exp'")
    _, batched, _, _, _, fin = self._evaluate(pop_lambdas)
    return self._shape(fin["log_hyper"], batched)

  def __call__(self, **hyper_lambdas):
    return self.compute_log_hyperlike(**hyper_lambdas)

  def compute_all(self, **hyper_lambdas):
    """(log_like_evs, log_like_num, log N_exp, log_hyper) -- the debugging entry of the reference."""
    pop_lambdas = self.population.update(**hyper_lambdas)
    _, batched, lle, _, _, fin = self._evaluate(pop_lambdas, want_events=True)
    return (self._shape(lle, batched), self._shape(fin["log_like_num"], batched),
            self._shape(fin["log_Nexp"], batched), self._shape(fin["log_hyper"], batched))
