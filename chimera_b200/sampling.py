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


def get_initial_state(nwalkers, ndim, log_prior, distribution="gaussian", priors=None, gaussian_bests=None,
                      gaussian_sigmas=None, rng=None):
  """Initial walker positions (`emcee_utils.py:66-160`, without the chain-restart branch that needs h5py):
  'gaussian' / 'uniform' redraw a walker until its prior is finite, 'truncgauss' replaces the coordinates that
  fall outside the box by uniform draws."""
  rng = np.random.default_rng() if rng is None else rng
  priors = np.tile([-np.inf, np.inf], (ndim, 1)) if priors is None else np.asarray(priors, dtype=np.float64)
  best = np.ones(ndim) if gaussian_bests is None else np.asarray(gaussian_bests, dtype=np.float64)
  sig = np.full(ndim, 0.2) if gaussian_sigmas is None else np.asarray(gaussian_sigmas, dtype=np.float64)
  if distribution == "truncgauss":
    start = rng.normal(best, sig, size=(nwalkers, ndim))
    out = (start < priors[:, 0]) | (start > priors[:, 1])
    for i in range(ndim):
      start[out[:, i], i] = rng.uniform(priors[i, 0], priors[i, 1], size=int(out[:, i].sum()))
    return start
  if distribution not in ("gaussian", "uniform"):
    raise ValueError("Only admitted distributions are 'gaussian', 'uniform', and 'truncgauss'.")
  draw = (lambda: rng.normal(best, sig)) if distribution == "gaussian" else (lambda: rng.uniform(priors[:, 0], priors[:, 1]))
  start = np.empty((nwalkers, ndim))
  for i in range(nwalkers):
    p = draw()
    while not np.isfinite(log_prior(p)):
      p = draw()
    start[i] = p
  return start


class EnsembleSampler:
  """Affine-invariant ensemble sampler (Goodman & Weare stretch move, red-blue split) that always evaluates a whole
  half-ensemble in ONE call of a vectorised `log_prob_fn` -- the access pattern of `emcee.EnsembleSampler(...,
  vectorize=True)` which the reference drives (`emcee_utils.py:226-334`); emcee itself is not available offline.
  Like the reference's `CustomEnsembleSampler`, non-finite coordinates are not rejected before the call."""

  def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, rng=None):
    if nwalkers < 2 * ndim or nwalkers % 2:
      raise ValueError("need an even number of walkers, at least twice the number of dimensions")
    self.nwalkers, self.ndim, self.log_prob_fn, self.a = nwalkers, ndim, log_prob_fn, float(a)
    self.rng = np.random.default_rng() if rng is None else rng
    self.chain = self.log_prob = None
    self.acceptance_fraction = np.zeros(nwalkers)

  def run_mcmc(self, p0, nsteps):
    x = np.array(p0, dtype=np.float64)
    if x.shape != (self.nwalkers, self.ndim):
      raise ValueError("p0 must have shape (nwalkers, ndim)")
    lp = np.asarray(self.log_prob_fn(x), dtype=np.float64)
    if np.any(np.isnan(lp)):
      raise ValueError("Probability function returned NaN")
    chain = np.empty((nsteps, self.nwalkers, self.ndim))
    lps = np.empty((nsteps, self.nwalkers))
    accepted = np.zeros(self.nwalkers)
    half = self.nwalkers // 2
    for t in range(nsteps):
      order = self.rng.permutation(self.nwalkers)
      for s, c in ((order[:half], order[half:]), (order[half:], order[:half])):
        zz = ((self.a - 1.0) * self.rng.random(half) + 1.0) ** 2 / self.a          # g(z) ~ 1/sqrt(z) on [1/a, a]
        partner = x[c[self.rng.integers(0, half, half)]]
        q = partner + zz[:, None] * (x[s] - partner)
        lq = np.asarray(self.log_prob_fn(q), dtype=np.float64)
        if np.any(np.isnan(lq)):
          raise ValueError("Probability function returned NaN")
        with np.errstate(invalid="ignore"):
          lnp = (self.ndim - 1.0) * np.log(zz) + lq - lp[s]
        acc = lnp > np.log(self.rng.random(half))
        x[s[acc]] = q[acc]
        lp[s[acc]] = lq[acc]
        accepted[s[acc]] += 1
      chain[t], lps[t] = x, lp
    self.chain, self.log_prob = chain, lps
    self.acceptance_fraction = accepted / max(nsteps, 1)
    return x, lp
