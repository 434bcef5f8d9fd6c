"""Shared machinery of the model plug-ins: parameter structs with the reference's
`default` / `keys` / `name` / `as_dict` / `update(**kwargs)` contract
(CHIMERA/population/cosmo.py:13-40, mass.py:13-42, rate.py:10-30) and the bridge that evaluates
the plum-dispatched free functions on the GPU through `chb_model_eval`."""
import ctypes as C
import numpy as np
from .. import _lib


class base_struct:
  default = {}
  name = "base_struct"

  def __init__(self, **kwargs):
    self.keys = list(self.default.keys())
    for key in self.keys:
      setattr(self, key, kwargs.get(key, self.default[key]))

  @property
  def as_dict(self):
    return {k: getattr(self, k) for k in self.keys}

  def update(self, **kwargs):
    """New instance with the matching keys replaced; unknown keys are ignored and `self` is
    returned unchanged when nothing matches."""
    hit = {k: v for k, v in kwargs.items() if k in self.keys}
    if not hit:
      return self
    vals = self.as_dict
    vals.update(hit)
    return self.__class__(**vals)

  def __repr__(self):
    return f"{self.__class__.__name__}({', '.join(f'{k}={getattr(self, k)!r}' for k in self.keys)})"

  def _is_scalar(self):
    return all(_ndim(getattr(self, k)) == 0 for k in self.keys)


def _ndim(v):
  """np.ndim without its dispatch overhead for the two common cases (this runs ~40 times per likelihood call)."""
  nd = getattr(v, "ndim", None)
  if nd is not None:
    return nd
  if isinstance(v, (int, float)):
    return 0
  return np.ndim(v)


def fill_rows(rows, struct):
  """Write a struct's parameters into the hyper-point matrix `rows` (n, CHB_NPAR)."""
  for k in struct.keys:
    slot = _lib.SLOT.get(k)
    if slot is not None:
      rows[:, slot] = getattr(struct, k)


def batch_size(*structs_and_scalars):
  n = None
  for s in structs_and_scalars:
    vals = [getattr(s, k) for k in s.keys] if isinstance(s, base_struct) else [s]
    for v in vals:
      nd = _ndim(v)
      if nd == 1:
        if n is not None and n != len(v):
          raise ValueError("batched hyper-parameters must have equal lengths")
        n = len(v)
      elif nd > 1:
        raise ValueError("hyper-parameters must be scalars or 1-D arrays")
  return n


_NEUTRAL = dict(H0=70., Om0=0.25, w0=-1., Xi0=1., z_max=10., m_low=5.1, m_high=87., alpha=3.4, beta=1.1, delta_m=4.8,
                alpha_2=5.6, break_fraction=0.43, lambda_peak=0.039, mu_g=34., sigma_g=3.6, gamma=2.7, kappa=3., zp=2.,
                zmax=1.3)
_template = None


def base_rows(n, cosmo=None, mass=None, rate=None, R0=1.0):
  global _template
  if _template is None:
    # neutral values for slots a model does not own
    t = np.zeros(_lib.CHB_NPAR)
    for k, v in _NEUTRAL.items():
      t[_lib.SLOT[k]] = v
    _template = t
  rows = np.empty((n, _lib.CHB_NPAR))
  rows[:] = _template
  for s in (cosmo, mass, rate):
    if s is not None:
      fill_rows(rows, s)
  rows[:, _lib.SLOT["R0"]] = R0
  # end knots of m_grid exactly as jnp.logspace(log10 m_low, log10 m_high, res) yields them
  # (mass.py:46): whether they pass `m_low <= m <= m_high` depends on the last ulp.
  rows[:, 26] = np.power(10., np.log10(rows[:, _lib.SLOT["m_low"]]))
  rows[:, 27] = np.power(10., np.log10(rows[:, _lib.SLOT["m_high"]]))
  return rows


def model_config(cosmo=None, mass=None, rate=None, device=0, **extra):
  """chb_config carrying the model ids / table resolutions of the given structs."""
  cfg = _lib.chb_config()
  cfg.abi_version = _lib.CHB_ABI_VERSION
  cfg.device = int(device)
  cfg.cosmo_model = _lib.COSMO_IDS[cosmo.name] if cosmo is not None else 0
  cfg.mass_model = _lib.MASS_IDS[mass.name] if mass is not None else 2
  cfg.rate_model = _lib.RATE_IDS[rate.name] if rate is not None else 1
  cfg.cosmo_grid_res = int(getattr(cosmo, "z_grid_res", 1500)) if cosmo is not None else 1500
  cfg.mass_grid_res = int(getattr(mass, "grid_res", 1000)) if mass is not None else 1000
  cfg.use_cut_grid = 1
  cfg.cut_grid = 2.0
  cfg.num_bins = 200
  cfg.scale_free = 1
  cfg.Tobs = 1.0
  cfg.N_inj = 1.0
  for k, v in extra.items():
    setattr(cfg, k, v)
  return cfg


def model_eval(which, a, b=None, c=None, cosmo=None, mass=None, rate=None, R0=1.0):
  """Element-wise evaluation of one plug-in function on the GPU for a scalar parameter set."""
  for s in (cosmo, mass, rate):
    if s is not None and not s._is_scalar():
      raise ValueError("array-valued parameters are only accepted by the batched likelihood entry points")
  lib = _lib.load()
  a = np.asarray(a, dtype=np.float64)
  shape = a.shape
  arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), shape)).ravel() if x is not None else None
          for x in (a, b, c)]
  out = np.empty(arrs[0].shape, dtype=np.float64)
  cfg = model_config(cosmo, mass, rate)
  rows = base_rows(1, cosmo, mass, rate, R0)
  rc = lib.chb_model_eval(C.byref(cfg), int(which), _lib.dptr(rows), arrs[0].size, _lib.dptr(arrs[0]),
                          _lib.dptr(arrs[1]), _lib.dptr(arrs[2]), _lib.dptr(out))
  _lib.check(rc)
  return out.reshape(shape) if shape else np.float64(out[0])


def model_tables(cosmo=None, mass=None):
  lib = _lib.load()
  cfg = model_config(cosmo, mass, None)
  rows = base_rows(1, cosmo, mass, None)
  zg = np.empty(cfg.cosmo_grid_res)
  ii = np.empty(cfg.cosmo_grid_res)
  mg = np.empty(cfg.mass_grid_res)
  cdf = np.empty(cfg.mass_grid_res)
  norm = np.empty(1)
  _lib.check(lib.chb_model_tables(C.byref(cfg), _lib.dptr(rows), _lib.dptr(zg), _lib.dptr(ii), _lib.dptr(mg),
                                  _lib.dptr(cdf), _lib.dptr(norm)))
  return zg, ii, mg, cdf, float(norm[0])
