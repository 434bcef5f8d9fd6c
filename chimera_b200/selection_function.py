"""`selection_function` with the constructor and `N_exp` of CHIMERA/selection_function.py:10-53.
The per-injection rate, importance weights and both reductions run in csrc/selection.cu; the
N_eff gate is the host epilogue `chb_finalize`."""
import numpy as np
from . import _lib
from .engine import Engine
from .population._base import model_config


class selection_function(object):
  def __init__(self, theta_inj_det, N_inj, N_eff=5., device=None, fp_mode='fp64'):
    if fp_mode not in _lib.FP_IDS:
      raise ValueError("fp_mode must be 'fp64' or 'fp32'")
    self.fp_mode = fp_mode                 # arithmetic of the stand-alone N_exp (the likelihood's own handle has its fp_mode)
    self.theta_inj_det = theta_inj_det
    self.N_inj = N_inj
    self.N_eff = N_eff
    self.device = device
    self._engines = {}

  def _config_fields(self):
    return dict(N_inj=float(self.N_inj), check_neff=0 if self.N_eff is None else 1,
                N_eff=0.0 if self.N_eff is None else float(self.N_eff))

  def _engine(self, pop):
    key = (pop.cosmo.name, pop.mass.name, pop.rate.name, int(pop.cosmo.z_grid_res), int(pop.mass.grid_res),
           float(pop.Tobs), bool(pop.scale_free))
    if key not in self._engines:
      cfg = model_config(pop.cosmo, pop.mass, pop.rate, device=self.device or 0, Tobs=float(pop.Tobs),
                         scale_free=int(bool(pop.scale_free)), fp_mode=_lib.FP_IDS[self.fp_mode], **self._config_fields())
      eng = Engine(cfg)
      t = self.theta_inj_det
      eng.set_injections(t.m1det, t.m2det, t.dL, t.p_draw)
      self._engines[key] = eng
    return self._engines[key]

  def N_exp(self, pop_lambdas):
    """Expected number of detections Tobs * xi, 0 when the injection N_eff is too small
    (selection_function.py:34-48).  Scalar or (n_hyper,) for batched hyper-parameters."""
    eng = self._engine(pop_lambdas)
    rows, batched = pop_lambdas.hyper_rows()
    _, part, _ = eng.eval(rows, want_events=False)
    out = eng.finalize(rows, part, 0)["N_exp"]
    return out if batched else np.float64(out[0])

  def __call__(self, pop_lambdas):
    return self.N_exp(pop_lambdas)
