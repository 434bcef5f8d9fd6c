"""Batched-sampler front end: the walker matrix of a vectorised ensemble sampler
-> ONE batched evaluation of the hyper-likelihood on the GPU.

`generate_dict` is the reference's `CHIMERA/utils/emcee_utils.py:54-64` (NumPy arrays instead of jnp).
`log_prob_fn` builds the callable an `emcee.EnsembleSampler(..., vectorize=True)` expects: it receives the
(nwalkers, ndim) position matrix, evaluates the prior on the host, and sends only the walkers with a finite
prior to `hyperlikelihood.__call__` as equal-length arrays (one `chb_eval` for the whole ensemble)."""
import numpy as np


def generate_dict(params, params_keys, to_calc=None):
  params = np.asarray(params)
  if params.ndim > 1:
    if to_calc is None:
      return {k: np.array(params[:, i]) for i, k in enumerate(params_keys)}
    return {k: np.array(params[to_calc, i]) for i, k in enumerate(params_keys)}
  return {k: params[i] for i, k in enumerate(params_keys)}


def uniform_log_prior(priors):
  """Flat prior inside the box `priors` (ndim, 2): 0 inside, -inf outside; works on (ndim,) and (n, ndim)."""
  priors = np.asarray(priors, dtype=np.float64)

  def log_prior(params):
    p = np.asarray(params, dtype=np.float64)
    inside = np.all((p >= priors[:, 0]) & (p <= priors[:, 1]), axis=-1)
    return np.where(inside, 0.0, -np.inf)
  return log_prior


def log_prob_fn(likelihood, params_keys, log_prior):
  """log-posterior callable for a vectorised ensemble sampler.  NaN likelihoods map to -inf."""
  keys = list(params_keys)

  def log_prob(params):
    p = np.asarray(params, dtype=np.float64)
    if p.ndim == 1:
      lp = float(log_prior(p))
      if not np.isfinite(lp):
        return -np.inf
      ll = float(likelihood(**generate_dict(p, keys)))
      return lp + ll if np.isfinite(ll) or ll == -np.inf else -np.inf
    lp = np.asarray(log_prior(p), dtype=np.float64)
    out = np.full(p.shape[0], -np.inf)
    to_calc = np.flatnonzero(np.isfinite(lp))
    if to_calc.size:
      ll = np.atleast_1d(likelihood(**generate_dict(p, keys, to_calc)))
      ll = np.where(np.isnan(ll), -np.inf, ll)
      out[to_calc] = lp[to_calc] + ll
    return out
  return log_prob
