"""Mass-function plug-in: structs and `p_m1m2` with the signatures of
CHIMERA/population/mass.py (`tpl` :56-84, `bpl` :86-115, `plp` :117-149, `p_m1m2` :334-345).
`pl2p` and `pls` are not provided: they cannot run in the reference either (mass.py:310-313
uses undefined names; `pls` has no normalisation and no `p_m1m2` dispatch)."""
from ._base import base_struct, model_eval, model_tables
from .. import _lib
from ..data import theta_src


class base_mass_paired_struct(base_struct):
  default = {'m_low': 5.1, 'm_high': 87., 'grid_res': 1000}
  name = 'base_mass_paired_struct'

  def _tables(self):
    if getattr(self, "_tab", None) is None:
      _, _, mg, cdf, norm = model_tables(mass=self)
      self._tab = (mg, cdf, norm)
    return self._tab

  @property
  def m_grid(self):
    return self._tables()[0]

  @property
  def cdf_m2_conditioned(self):
    return self._tables()[1]

  @property
  def norm_p_m1(self):
    return self._tables()[2]


class tpl(base_mass_paired_struct):
  default = {**base_mass_paired_struct.default, 'alpha': 2.5, 'beta': 1.1}
  name = 'truncated_power_law'


class bpl(base_mass_paired_struct):
  default = {**base_mass_paired_struct.default, 'alpha_1': 1.6, 'alpha_2': 5.6, 'beta': 1.1, 'delta_m': 4.8,
             'break_fraction': 0.43}
  name = 'broken_power_law'


class plp(base_mass_paired_struct):
  default = {**base_mass_paired_struct.default, 'lambda_peak': 0.039, 'alpha': 3.4, 'beta': 1.1, 'delta_m': 4.8,
             'mu_g': 34., 'sigma_g': 3.6}
  name = 'power_law_plus_peak'


def primary_mass_pdf_notnorm(mass, m):
  """mass.py:285-305."""
  return model_eval(_lib.F_P_M1_NOTNORM, m, mass=mass)


def p_m1m2(mass, m1, m2=None):
  """Joint source-frame mass pdf; `p_m1m2(mass, theta_src)` or `p_m1m2(mass, m1, m2)` (mass.py:334-349)."""
  if isinstance(m1, theta_src):
    m1, m2 = m1.m1src, m1.m2src
  return model_eval(_lib.F_P_M1M2, m1, m2, mass=mass)
