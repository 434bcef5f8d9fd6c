"""Cosmology plug-in: same structs, keys and free-function signatures as
CHIMERA/population/cosmo.py (`flrw` :50-84, `mg_flrw` :86-115, functions :122-264); the
arithmetic runs in the CUDA library (csrc/models.cuh, csrc/tables.cu)."""
import numpy as np
from ._base import base_struct, model_eval, model_tables
from .. import _lib
from ..data import theta_src


class base_cosmology_struct(base_struct):
  default = {'z_max': 10., 'z_grid_res': 1000}
  name = 'base_cosmology_struct'

  def _tables(self):
    if getattr(self, "_tab", None) is None:
      zg, ii, _, _, _ = model_tables(cosmo=self)
      self._tab = (zg, ii)
    return self._tab

  @property
  def z_grid_interp(self):
    """[0] U logspace(-10, log10 z_max, res-1)  (cosmo.py:43-46), built on the device."""
    return self._tables()[0]

  @property
  def integral_invE_interp(self):
    return self._tables()[1]


class flrw(base_cosmology_struct):
  """FLRW parameters (H0, Om0, Ok0, Or0, w0, wa); cosmo.py:50-84."""
  name = 'flrw'
  default = {**base_cosmology_struct.default, 'H0': 70., 'Om0': 0.25, 'Ok0': 0., 'Or0': 0., 'w0': -1., 'wa': 0.,
             'z_max': 10., 'z_grid_res': 1500}

  @property
  def Ode0(self):
    return 1.0 - self.Om0 - self.Or0 - self.Ok0

  @property
  def dH(self):
    return 299792.458e-3 / self.H0


class mg_flrw(flrw):
  """FLRW + modified GW propagation (Xi0, n); cosmo.py:86-115."""
  name = 'mg_flrw'
  default = {**flrw.default, 'Xi0': 1., 'n': 0.}


def _zd(z, distances):
  if isinstance(z, theta_src):
    return z.z, z.original_distances
  return z, distances


def E_at_z(cosmo, z):
  """Dimensionless Hubble parameter (cosmo.py:122-130)."""
  return model_eval(_lib.F_E_AT_Z, z, cosmo=cosmo)


def dCt_at_z(cosmo, z):
  """Transverse comoving distance [Gpc] (cosmo.py:142-153)."""
  return model_eval(_lib.F_DCT_AT_Z, z, cosmo=cosmo)


def dL_at_z(cosmo, z):
  """Luminosity (GW) distance [Gpc] (cosmo.py:205-210, 237-243)."""
  return model_eval(_lib.F_DL_AT_Z, z, cosmo=cosmo)


def ddLdz_at_z(cosmo, z, distances=None):
  """d dL / dz (cosmo.py:212-221, 245-257); accepts a `theta_src` like the reference's dispatch."""
  z, distances = _zd(z, distances)
  return model_eval(_lib.F_DDLDZ_AT_Z, z, distances, cosmo=cosmo)


def dVcdz_at_z(cosmo, z, distances=None):
  """Differential comoving volume (cosmo.py:188-197)."""
  z, distances = _zd(z, distances)
  return model_eval(_lib.F_DVCDZ_AT_Z, z, distances, cosmo=cosmo)


def Vc_at_z(cosmo, z, distances=None):
  """Comoving volume (cosmo.py:166-186)."""
  z, distances = _zd(z, distances)
  return model_eval(_lib.F_VC_AT_Z, z, distances, cosmo=cosmo)


def z_from_dGW(cosmo, dGWs):
  """Redshift of a GW distance by inverse table interpolation (cosmo.py:260-264)."""
  return model_eval(_lib.F_Z_FROM_DGW, dGWs, cosmo=cosmo)
