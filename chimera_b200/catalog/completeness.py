"""Completeness model with the surface of CHIMERA/catalog/completeness.py:22-67
(`dVdz_completeness`: `P_compl(zgrids)`, `fR(cosmo)`, `p_bkg(cosmo, z | theta_src)`).
Only `kind='step'` is supported: the reference's 'step_smooth' branch broadcasts a (2,) range
against (Nev,Nz) grids (completeness.py:48) and cannot run; `homogeneous_completeness`
(:73-216) reads attributes that are never set and is not reproduced."""
import numpy as np
from ..population.cosmo import dVcdz_at_z, Vc_at_z


class dVdz_completeness(object):
  def __init__(self, z_range=(0.073, 1.3), kind="step", z_sig=None):
    self.z_range = np.asarray(z_range, dtype=np.float64)
    if self.z_range.shape != (2,):
      raise ValueError("z_range must hold two redshifts")
    if kind != "step":
      raise ValueError("kind must be step (step_smooth is not usable in the reference either)")
    self.kind = kind
    self.z_sig = z_sig

  def P_compl(self, zgrids):
    """1 inside the complete range, 0 outside (completeness.py:43-46)."""
    zgrids = np.asarray(zgrids, dtype=np.float64)
    return np.where(np.logical_and(zgrids > self.z_range[0], zgrids < self.z_range[1]), 1., 0.)

  def fR(self, cosmo_lambdas, normalized=False):
    """Comoving volume of the complete shell (completeness.py:54-58)."""
    res = Vc_at_z(cosmo_lambdas, self.z_range)
    return res[1] - res[0]

  def p_bkg(self, cosmo_lambdas, z):
    """Background galaxy density: dVc/dz (completeness.py:60-67)."""
    return dVcdz_at_z(cosmo_lambdas, z)
