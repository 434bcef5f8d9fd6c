"""Merger-rate plug-in: structs and `merger_rate` of CHIMERA/population/rate.py:32-129."""
from ._base import base_struct, model_eval
from .. import _lib
from ..data import theta_src


class base_rate_struct(base_struct):
  default = {}
  name = 'base_rate_struct'


class power_law(base_rate_struct):
  name = 'power_law'
  default = {'gamma': 1.7}


class madau_dickinson(base_rate_struct):
  name = 'madau_dickinson'
  default = {'gamma': 2.7, 'kappa': 3.0, 'zp': 2.}


class trunc_madau_dickinson(base_rate_struct):
  name = 'trunc_madau_dickinson'
  default = {'gamma': 2.7, 'kappa': 3.0, 'zp': 2., 'zmax': 1.3}


class trunc_power_law(base_rate_struct):
  name = 'trunc_power_law'
  default = {'gamma': 1.9, 'zmax': 1.3}


def merger_rate(rate, z):
  """psi(z) (rate.py:96-129); accepts a `theta_src`."""
  if isinstance(z, theta_src):
    z = z.z
  return model_eval(_lib.F_MERGER_RATE, z, rate=rate)
