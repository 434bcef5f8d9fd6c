"""Population wrapper with the surface of CHIMERA/population/pop_wrapper.py: `population`
(:14-64) and the glue functions `theta_det2src` (:67-75), `get_theta_src_and_weights` (:77-80),
`p_cbc` (:82-90), `pop_rate_det` (:92-121), `compute_z_grids` (:133-208).

The likelihood itself never calls these per-array helpers: the same arithmetic is fused into the
CUDA kernels.  They exist so that reference users find the functions they know; each evaluates
its model terms on the GPU through `chb_model_eval`."""
from numbers import Number
import numpy as np
from ._base import batch_size, base_rows, model_eval
from .cosmo import dVcdz_at_z, ddLdz_at_z, z_from_dGW
from .mass import p_m1m2
from .rate import merger_rate
from .. import _lib
from ..catalog.catalog import empty_catalog
from ..data import theta_src, theta_pe_det, theta_inj_det


class population:
  """Bundle of cosmology, mass and rate structs + R0, catalogue, Tobs, scale_free."""

  def __init__(self, cosmo, mass, rate, R0=1., gal_cat=None, Tobs=1, scale_free=True):
    self.cosmo = cosmo
    self.mass = mass
    self.rate = rate
    self.R0 = R0
    self.gal_cat = empty_catalog(p_bkg='dVdz') if gal_cat is None else gal_cat
    self.Tobs = Tobs
    self.scale_free = scale_free

  def __repr__(self):
    return (f"cosmo = {self.cosmo},\nmass = {self.mass},\nrate = {self.rate},\nR0 = {self.R0},\n"
            f"galcat_obj = {self.gal_cat},\nTobs = {self.Tobs},\nscale_free = {self.scale_free}")

  def update(self, **hyper_lambdas):
    """Route keyword hyper-parameters to whichever struct lists the key (pop_wrapper.py:56-64).
    Values may be scalars or equal-length 1-D arrays (a batch of hyper-points)."""
    return self.__class__(self.cosmo.update(**hyper_lambdas), self.mass.update(**hyper_lambdas),
                          self.rate.update(**hyper_lambdas), hyper_lambdas.get('R0', self.R0),
                          self.gal_cat, self.Tobs, self.scale_free)

  # ---- bridge to the C ABI -------------------------------------------------------------
  def hyper_rows(self):
    """(n, CHB_NPAR) matrix of this population's hyper-point(s) and whether it is a batch."""
    n = batch_size(self.cosmo, self.mass, self.rate, self.R0)
    rows = base_rows(1 if n is None else n, self.cosmo, self.mass, self.rate, self.R0)
    return rows, n is not None


def theta_det2src(cosmo_lambdas, theta_det, include_original_distances=False):
  z = z_from_dGW(cosmo_lambdas, theta_det.dL)
  m1s, m2s = np.asarray(theta_det.m1det) / (1. + z), np.asarray(theta_det.m2det) / (1. + z)
  if include_original_distances:
    return theta_src(m1src=m1s, m2src=m2s, z=z, original_distances=np.asarray(theta_det.dL))
  return theta_src(m1src=m1s, m2src=m2s, z=z)


def get_theta_src_and_weights(pop_lambdas, theta_det):
  th_src = theta_det2src(pop_lambdas.cosmo, theta_det)
  with np.errstate(all="ignore"):
    weights = p_m1m2(pop_lambdas.mass, th_src) / np.asarray(theta_det.pe_prior)
  return th_src, weights


def p_cbc(pop_lambdas, z):
  """Redshift prior p_gal * psi/(1+z) with the -100 sentinel preserved (pop_wrapper.py:82-90)."""
  z = np.asarray(z, dtype=np.float64)
  p_gal = np.asarray(pop_lambdas.gal_cat.p_gal(pop_lambdas.cosmo, z))
  p_rate = merger_rate(pop_lambdas.rate, z) / (1 + z)
  if p_gal.ndim > p_rate.ndim:
    return np.where(p_gal != -100, p_gal * p_rate[:, None, :], -100)
  return p_gal * p_rate


def pop_rate_det(pop_lambdas, th):
  """Detector-frame population rate; dispatches on the struct type like the reference."""
  if isinstance(th, theta_inj_det):
    return model_eval(_lib.F_POP_RATE_DET_INJ, th.m1det, th.m2det, th.dL, cosmo=pop_lambdas.cosmo,
                      mass=pop_lambdas.mass, rate=pop_lambdas.rate, R0=pop_lambdas.R0)
  if isinstance(th, theta_pe_det):
    src = theta_det2src(pop_lambdas.cosmo, th)
    p_z = p_cbc(pop_lambdas, src.z)
    dN = pop_lambdas.R0 * p_m1m2(pop_lambdas.mass, src) * p_z
    return dN / (np.abs(ddLdz_at_z(pop_lambdas.cosmo, src)) * (1. + src.z) ** 2)
  if isinstance(th, theta_src):
    p_z = pop_lambdas.gal_cat.p_bkg(pop_lambdas.cosmo, th) * merger_rate(pop_lambdas.rate, th) / (1. + th.z)
    dN = pop_lambdas.R0 * p_m1m2(pop_lambdas.mass, th) * p_z
    return dN / (np.abs(ddLdz_at_z(pop_lambdas.cosmo, th)) * (1. + th.z) ** 2)
  raise TypeError("pop_rate_det expects theta_pe_det, theta_inj_det or theta_src")


def compute_z_grids(cosmo, theta_det, cosmo_prior=None, z_int_res=300, z_conf_range=None):
  """Per-event redshift integration grids (pop_wrapper.py:133-208): the event's dL range mapped
  to z with the prior-edge cosmologies on 10 000-point tables, then `linspace`."""
  events_dL = np.asarray(theta_det.dL, dtype=np.float64)
  if isinstance(z_conf_range, list):
    dL_min, dL_max = np.percentile(events_dL, z_conf_range, axis=1)
  elif isinstance(z_conf_range, Number):
    mu, sig = np.mean(events_dL, axis=1), np.std(events_dL, axis=1)
    dL_min, dL_max = mu - z_conf_range * sig, mu + z_conf_range * sig
  else:
    dL_max = np.max(events_dL, axis=1) * 2
    dL_min = np.min(events_dL, axis=1) * 0.5
    dL_min = np.where(dL_min < 1.e-8, 1.e-8, dL_min)
  cp = {k: [v, v] for k, v in cosmo.as_dict.items()}
  if cosmo_prior is not None:
    cp.update(cosmo_prior)
  names = ["H0", "Om0", "Ok0", "Or0", "w0", "wa"]
  lc_low = {k: cp[k][0] for k in names}
  lc_high = {k: cp[k][1] for k in names}
  if cosmo.name != 'flrw':
    lc_low.update(Xi0=cp['Xi0'][1], n=cp['n'][1])
    lc_high.update(Xi0=cp['Xi0'][0], n=cp['n'][1])
  cosmo1 = cosmo.update(**lc_low, z_grid_res=10_000)
  cosmo2 = cosmo.update(**lc_high, z_grid_res=10_000)
  z_min = z_from_dGW(cosmo1, dL_min)
  z_max = z_from_dGW(cosmo2, dL_max)
  return np.linspace(z_min, z_max, z_int_res, axis=1)
