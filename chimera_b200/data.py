"""Input structs of the likelihood path: same field names as the reference's
`theta_pe_det` / `theta_inj_det` / `theta_src` (CHIMERA/data.py:15-59), as plain NumPy holders."""
import copy
import numpy as np


class theta_generic:
  _fields = ()

  def __init__(self, **kwargs):
    for k in self._fields:
      v = kwargs.pop(k, None)
      setattr(self, k, v)
    if kwargs:
      raise TypeError(f"unexpected fields {list(kwargs)}")
    self.__post_init__()

  def __post_init__(self):
    pass

  def update(self, **kwargs):
    """Functional update (CHIMERA/data.py:15-25): returns a copy with the given fields replaced."""
    new = copy.copy(self)
    for k, v in kwargs.items():
      if k not in self._fields:
        raise AttributeError(k)
      setattr(new, k, v)
    return new


class theta_pe_det(theta_generic):
  """Detector-frame PE samples (+ optional pixelisation fields), CHIMERA/data.py:27-47."""
  _fields = ("m1det", "m2det", "dL", "phi", "theta", "ra", "dec", "pe_prior", "pixels_pe_all_nsides",
             "opt_nsides", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside")

  def __post_init__(self):
    if self.pe_prior is None and self.dL is not None:
      self.pe_prior = np.ones_like(np.asarray(self.dL, dtype=np.float64))


class theta_inj_det(theta_generic):
  """Detected injections, CHIMERA/data.py:49-53."""
  _fields = ("m1det", "m2det", "dL", "p_draw")


class theta_src(theta_generic):
  """Source-frame parameters, CHIMERA/data.py:55-59."""
  _fields = ("m1src", "m2src", "z", "original_distances")


theta_pe_datasets = ['m1det', 'm2det', 'dL', 'pe_prior']
theta_pe_pixelated_datasets = ['m1det', 'm2det', 'dL', 'pe_prior', 'ra', 'dec', 'theta', 'phi',
                               'opt_nsides', 'pixels_opt_nsides', 'ra_pix', 'dec_pix', 'gw_loc2d_pdf',
                               'pixels_pe_opt_nside']
theta_pe_pixelated_groups = ['pixels_pe_all_nsides']
