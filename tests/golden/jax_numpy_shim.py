"""NumPy-backed stand-ins for jax / equinox / plum / healpy / h5py.

PURPOSE (test infrastructure only, never imported by the product):
JAX, equinox, plum-dispatch, healpy and h5py are not installable offline, so the reference
package at /root/reference cannot be imported as is.  This module registers minimal stand-ins
in `sys.modules` so that the UNMODIFIED reference source can be imported and executed in the
build container, with NumPy (IEEE fp64, same as `jax_enable_x64`) providing `jax.numpy`.
`tests/golden/make_golden.py` uses it to produce the golden input/output fixtures that pin
`oracle/chimera_oracle.py`.

What is and is not the reference here:
  * every arithmetic statement executed is the reference's own Python source;
  * `jnp.*` resolves to the NumPy function of the same name (identical semantics for the
    calls the hot path makes: interp/trapezoid/std/linspace/cumsum/where/nan_to_num/...;
    results can differ from XLA in the last ulp because reductions associate differently);
  * `jax.jit` is the identity, `jax.vmap` is a Python loop + stack, `lax.cond`/`fori_loop`
    are Python control flow, `io_callback` calls the callback directly;
  * `equinox.Module` is a plain attribute container, `plum.dispatch` is a small
    isinstance-based multiple dispatcher;
  * `healpy` resolves to `chimera_b200.healpix` (our RING restatement) -- healpy itself stays
    unpinned; `h5py` is an empty stub (no file I/O is exercised).
"""
import sys
import types
import inspect
import typing
import numpy as np


# --------------------------------------------------------------------------- jax.numpy
class JArr(np.ndarray):
  """ndarray with the `.at[idx].add/set` functional-update helpers of jax arrays."""

  @property
  def at(self):
    return _At(self)


class _At:
  def __init__(self, arr):
    self.arr = arr

  def __getitem__(self, idx):
    return _AtIdx(self.arr, idx)


class _AtIdx:
  def __init__(self, arr, idx):
    self.arr, self.idx = arr, idx

  def add(self, vals):
    out = np.array(self.arr, copy=True)
    idx = self.idx
    if isinstance(idx, np.ndarray) and idx.dtype.kind in "iu":
      # jax scatter semantics: out-of-bounds updates are dropped
      n = out.shape[0]
      idx = np.where(idx < 0, idx + n, idx)
      ok = (idx >= 0) & (idx < n)
      vals = np.broadcast_to(np.asarray(vals), idx.shape)
      np.add.at(out, idx[ok], vals[ok])
    else:
      np.add.at(out, idx, vals)
    return out.view(JArr)

  def set(self, vals):
    out = np.array(self.arr, copy=True)
    out[self.idx] = vals
    return out.view(JArr)


def _wrap_out(x):
  if isinstance(x, np.ndarray) and not isinstance(x, JArr):
    return x.view(JArr)
  if isinstance(x, tuple):
    return tuple(_wrap_out(v) for v in x)
  if isinstance(x, list):
    return [_wrap_out(v) for v in x]
  return x


def _wrap_fn(fn):
  def wrapped(*a, **k):
    with np.errstate(all="ignore"):
      return _wrap_out(fn(*a, **k))
  wrapped.__name__ = getattr(fn, "__name__", "fn")
  return wrapped


class _JnpModule(types.ModuleType):
  ndarray = np.ndarray

  def __getattr__(self, name):
    if name == "trapz":
      raise AttributeError(name)
    attr = getattr(np, name)
    if isinstance(attr, type) or not callable(attr):
      return attr
    w = _wrap_fn(attr)
    setattr(self, name, w)
    return w


jnp = _JnpModule("jax.numpy")
jnp.array = _wrap_fn(np.array)
jnp.asarray = _wrap_fn(np.asarray)
jnp.isscalar = np.isscalar
jnp.linalg = types.SimpleNamespace(inv=_wrap_fn(np.linalg.inv), cholesky=_wrap_fn(np.linalg.cholesky),
                                   det=_wrap_fn(np.linalg.det))


def _interp(x, xp, fp, left=None, right=None, period=None):
  with np.errstate(all="ignore"):
    return _wrap_out(np.interp(x, xp, fp, left=left, right=right, period=period))


jnp.interp = _interp


# --------------------------------------------------------------------------- jax
def _jit(fn=None, **kwargs):
  if fn is None:
    return lambda f: f
  return fn


def _vmap(fn, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    n = None
    for a, ax in zip(args, axes):
      if ax is not None:
        n = np.shape(a)[ax]
        break
    outs = []
    for i in range(n):
      call = [np.take(a, i, axis=ax) if ax is not None else a for a, ax in zip(args, axes)]
      outs.append(fn(*call))
    if isinstance(outs[0], tuple):
      return tuple(_wrap_out(np.stack([o[k] for o in outs], axis=out_axes)) for k in range(len(outs[0])))
    return _wrap_out(np.stack(outs, axis=out_axes))
  return mapped


def _cond(pred, true_fn, false_fn, *operands, operand=None):
  if operands:
    return true_fn(*operands) if bool(pred) else false_fn(*operands)
  return true_fn(operand) if bool(pred) else false_fn(operand)


def _fori_loop(lo, hi, body, init):
  val = init
  for i in range(int(lo), int(hi)):
    val = body(i, val)
  return val


def _io_callback(cb, result_shape, *args, **kwargs):
  return _wrap_out(cb(*args))


class _Config:
  def update(self, *a, **k):
    pass


def install():
  """Register the stand-in modules. Idempotent."""
  if "jax" in sys.modules and getattr(sys.modules["jax"], "_chb_shim", False):
    return
  import scipy.special
  import scipy.integrate

  jax = types.ModuleType("jax")
  jax._chb_shim = True
  jax.__version__ = "0.5.0"
  jax.config = _Config()
  jax.numpy = jnp
  jax.jit = _jit
  jax.vmap = _vmap
  jax.lax = types.SimpleNamespace(cond=_cond, fori_loop=_fori_loop)
  jax.ShapeDtypeStruct = lambda shape, dtype: (shape, dtype)
  jax.random = types.SimpleNamespace(PRNGKey=lambda s: s)
  jscipy = types.ModuleType("jax.scipy")
  jscipy.special = types.SimpleNamespace(erf=_wrap_fn(scipy.special.erf),
                                         logsumexp=_wrap_fn(scipy.special.logsumexp))
  jscipy.integrate = types.SimpleNamespace(trapezoid=_wrap_fn(np.trapezoid))
  jax.scipy = jscipy
  jexp = types.ModuleType("jax.experimental")
  jexp.io_callback = _io_callback
  jax.experimental = jexp
  sys.modules["jax"] = jax
  sys.modules["jax.numpy"] = jnp
  sys.modules["jax.scipy"] = jscipy
  sys.modules["jax.experimental"] = jexp

  # ------------------------------------------------------------------------- equinox
  eqx = types.ModuleType("equinox")

  class Module:
    """Attribute container: dataclass-like __init__ from annotations unless overridden."""

    def __init__(self, *args, **kwargs):
      fields = []
      for klass in reversed(type(self).__mro__):
        for name in getattr(klass, "__annotations__", {}):
          if name not in fields:
            fields.append(name)
      for name, val in zip(fields, args):
        object.__setattr__(self, name, val)
      for name in fields[len(args):]:
        if name in kwargs:
          object.__setattr__(self, name, kwargs.pop(name))
        else:
          object.__setattr__(self, name, getattr(type(self), name, None))
      if kwargs:
        raise TypeError(f"unexpected fields {list(kwargs)}")
      post = getattr(self, "__post_init__", None)
      if post is not None:
        post()

  def field(static=False, default=None, **kw):
    return default

  def tree_at(where, pytree, replace, is_leaf=None):
    import copy

    class _Probe:
      def __getattr__(self, name):
        return ("__field__", name)

    tag = where(_Probe())
    if not (isinstance(tag, tuple) and tag and tag[0] == "__field__"):
      raise ValueError("tree_at: only single-attribute selectors are supported")
    new = copy.copy(pytree)
    object.__setattr__(new, tag[1], replace)
    return new

  eqx.Module = Module
  eqx.field = field
  eqx.tree_at = tree_at
  sys.modules["equinox"] = eqx

  # ------------------------------------------------------------------------- plum
  plum = types.ModuleType("plum")
  _registry = {}

  def _match(ann, val):
    """(matches, specificity) of value `val` against annotation `ann`."""
    if ann is inspect.Parameter.empty or ann is typing.Any or ann is object:
      return True, 1000
    origin = typing.get_origin(ann)
    if origin is typing.Union:
      best = None
      for sub in typing.get_args(ann):
        ok, sc = _match(sub, val)
        if ok and (best is None or sc < best):
          best = sc
      return (best is not None), (best if best is not None else 0)
    if origin is not None:
      ann = origin
    if ann is type(None):
      return val is None, 0
    if isinstance(ann, type):
      if isinstance(val, ann):
        mro = type(val).__mro__
        return True, (mro.index(ann) if ann in mro else 500)
      return False, 0
    return True, 1000

  def dispatch(fn):
    key = (fn.__module__, fn.__qualname__)
    _registry.setdefault(key, []).append((inspect.signature(fn), fn))

    def dispatcher(*args, **kwargs):
      best, best_score = None, None
      for sig, cand in _registry[key]:
        try:
          bound = sig.bind(*args, **kwargs)
        except TypeError:
          continue
        score, ok = 0, True
        for name, val in bound.arguments.items():
          m, sc = _match(sig.parameters[name].annotation, val)
          if not m:
            ok = False
            break
          score += sc
        if ok and (best is None or score < best_score):
          best, best_score = cand, score
      if best is None:
        raise TypeError(f"no dispatch for {key} with {[type(a).__name__ for a in args]}")
      return best(*args, **kwargs)
    dispatcher.__name__ = fn.__name__
    dispatcher.__qualname__ = fn.__qualname__
    return dispatcher

  plum.dispatch = dispatch
  sys.modules["plum"] = plum

  # ------------------------------------------------------------------------- healpy / h5py
  from chimera_b200 import healpix as _hpx
  hp = types.ModuleType("healpy")
  hp.ang2pix = _hpx.ang2pix
  hp.pix2ang = _hpx.pix2ang
  hp.nside2npix = _hpx.nside2npix
  sys.modules["healpy"] = hp
  sys.modules["h5py"] = types.ModuleType("h5py")


def import_reference(path="/root/reference"):
  """Import the unmodified reference package under the stand-ins and return it."""
  install()
  import logging
  if path not in sys.path:
    sys.path.insert(0, path)
  np.seterr(all="ignore")
  import CHIMERA  # noqa: E402
  logging.getLogger("CHIMERA").setLevel(logging.WARNING)
  return CHIMERA
