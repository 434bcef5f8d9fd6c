"""CPU oracle for the CHIMERA hierarchical-likelihood hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, not the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing under
`chimera_b200/` imports it, and the product fails loudly when the CUDA library is missing.

It is a NumPy fp64 restatement of the reference algorithm (CHIMERA v2.0.0), written from the
behaviour of the reference files cited at each function (paths relative to /root/reference).

PINNING STATUS.  The reference ships no tests, golden vectors or fixtures,
and its run-time dependencies (jax, equinox, plum, healpy, h5py) cannot be installed offline,
so "the reference itself run here" is only possible in this form: the UNMODIFIED reference
source executed with NumPy standing in for `jax.numpy` (`tests/golden/jax_numpy_shim.py`).
`tests/golden/make_golden.py` generated the committed fixtures `tests/golden/*.npz` that way,
and `tests/test_oracle_golden.py` checks every function below against them.  The oracle is
therefore pinned to the reference's source semantics; it is NOT pinned to XLA's floating-point
evaluation order (expected last-ulp differences) nor to healpy (third-party, unpinned: the
RING indexing is restated in `chimera_b200/healpix.py` and property-tested).

Conventions: model parameters are plain dicts; `model` keys name the reference struct
(`flrw`, `mg_flrw`; `tpl`, `bpl`, `plp`; `power_law`, `madau_dickinson`,
`trunc_madau_dickinson`, `trunc_power_law`).
"""
import math
import numpy as np
from scipy.special import erf

np_trapz = np.trapezoid if hasattr(np, "trapezoid") else np.trapz

C_KM_S_E3 = 299792.458e-3   # dH = c/H0 in Gpc (cosmo.py:83-84)

COSMO_DEFAULT = dict(model="flrw", H0=70., Om0=0.25, Ok0=0., Or0=0., w0=-1., wa=0.,
                     Xi0=1., n=0., z_max=10., z_grid_res=1500)           # cosmo.py:77,115
MASS_DEFAULT = {
  "tpl": dict(m_low=5.1, m_high=87., grid_res=1000, alpha=2.5, beta=1.1),              # mass.py:78
  "bpl": dict(m_low=5.1, m_high=87., grid_res=1000, alpha_1=1.6, alpha_2=5.6, beta=1.1,
              delta_m=4.8, break_fraction=0.43),                                       # mass.py:114
  "plp": dict(m_low=5.1, m_high=87., grid_res=1000, lambda_peak=0.039, alpha=3.4, beta=1.1,
              delta_m=4.8, mu_g=34., sigma_g=3.6),                                     # mass.py:148
}
RATE_DEFAULT = {
  "power_law": dict(gamma=1.7),                                    # rate.py:48
  "madau_dickinson": dict(gamma=2.7, kappa=3.0, zp=2.),            # rate.py:71
  "trunc_madau_dickinson": dict(gamma=2.7, kappa=3.0, zp=2., zmax=1.3),   # rate.py:80
  "trunc_power_law": dict(gamma=1.9, zmax=1.3),                    # rate.py:87
}


def cumtrapz(y, x):
  """Cumulative trapezoid with a leading 0 (utils/math.py:22-26)."""
  dx = np.diff(x)
  return np.concatenate([[0.0], np.cumsum(0.5 * (y[:-1] + y[1:]) * dx)])


# =========================================================================== cosmology
def make_cosmo(model="flrw", **kw):
  c = dict(COSMO_DEFAULT)
  c["model"] = model
  for k, v in kw.items():
    if k in c:
      c[k] = v
  return cosmo_setup(c)


def cosmo_setup(c):
  """Interpolation tables rebuilt for every hyper-point (cosmo.py:43-46)."""
  res = int(c["z_grid_res"])
  zg = np.concatenate([[0.0], np.logspace(-10, np.log10(c["z_max"]), res - 1)])
  c = dict(c)
  c["z_grid_interp"] = zg
  c["integral_invE_interp"] = cumtrapz(1.0 / E_at_z(c, zg), zg)
  return c


def cosmo_update(c, **kw):
  """`update` ignores unknown keys and rebuilds tables when something matched (cosmo.py:33-40)."""
  keys = [k for k in kw if k in COSMO_DEFAULT and k != "model"]
  if c["model"] == "flrw":
    keys = [k for k in keys if k not in ("Xi0", "n")]
  if not keys:
    return c
  new = {k: c[k] for k in COSMO_DEFAULT}
  for k in keys:
    new[k] = kw[k]
  return cosmo_setup(new)


def E_at_z(c, z):
  """cosmo.py:122-130."""
  z = np.asarray(z, dtype=np.float64)
  Ode0 = 1.0 - c["Om0"] - c["Or0"] - c["Ok0"]
  w_z = c["w0"] + c["wa"] * z / (1 + z)
  return np.sqrt(c["Om0"] * (1. + z) ** 3 + c["Or0"] * (1. + z) ** 4 + c["Ok0"] * (1. + z) ** 2
                 + Ode0 * (1. + z) ** (3. * (1. + w_z)))


def dH(c):
  return C_KM_S_E3 / c["H0"]


def dCt_at_z(c, z):
  """Transverse comoving distance from the table (cosmo.py:132-153)."""
  dCr = dH(c) * np.interp(z, c["z_grid_interp"], c["integral_invE_interp"])
  Ok0 = c["Ok0"]
  if Ok0 == 0.0:
    return dCr
  s = np.sqrt(np.abs(Ok0 + 1.e-10))
  if Ok0 > 0.0:
    return (dH(c) / s) * np.sinh(s * dCr / dH(c))
  return (dH(c) / s) * np.sin(s * dCr / dH(c))


def Xi_at_z(c, z):
  """cosmo.py:225-228."""
  return c["Xi0"] + (1. - c["Xi0"]) / ((1. + z) ** c["n"])


def _dL2dCt(c, distances, z):
  """cosmo.py:201-203 (flrw), 230-235 (mg_flrw)."""
  if c["model"] == "mg_flrw":
    return (distances / Xi_at_z(c, z)) / (1. + z)
  return distances / (1. + z)


def dL_at_z(c, z):
  """cosmo.py:205-210 / 237-243."""
  z = np.asarray(z, dtype=np.float64)
  dL = dCt_at_z(c, z) * (1. + z)
  if c["model"] == "mg_flrw":
    dL = dL * Xi_at_z(c, z)
  return dL


def ddLdz_at_z(c, z, distances=None):
  """cosmo.py:212-221 / 245-257."""
  z = np.asarray(z, dtype=np.float64)
  dCt = _dL2dCt(c, distances, z) if distances is not None else dCt_at_z(c, z)
  ddL = dCt + (dH(c) / E_at_z(c, z)) * (1. + z)
  if c["model"] == "mg_flrw":
    dLflrw = dCt * (1. + z)
    dXiz = c["n"] * (c["Xi0"] - 1.) / ((1. + z) ** (c["n"] + 1))
    return ddL * Xi_at_z(c, z) + dLflrw * dXiz
  return ddL


def dVcdz_at_z(c, z, distances=None):
  """cosmo.py:188-197."""
  z = np.asarray(z, dtype=np.float64)
  dCt = _dL2dCt(c, distances, z) if distances is not None else dCt_at_z(c, z)
  return 4 * np.pi * dH(c) * dCt ** 2 / E_at_z(c, z)


def Vc_at_z(c, z, distances=None):
  """cosmo.py:166-186."""
  z = np.asarray(z, dtype=np.float64)
  dCt = _dL2dCt(c, distances, z) if distances is not None else dCt_at_z(c, z)
  Ok0 = c["Ok0"]
  if Ok0 == 0.0:
    return 4. * np.pi * dCt ** 3 / 3.
  reg = Ok0 + 1e-10
  s = np.sqrt(np.abs(reg))
  d = dH(c)
  pref = 4. * np.pi * d ** 3 / (2. * reg)
  if Ok0 > 0.0:
    return pref * ((dCt / d) * np.sqrt(1 + reg * dCt ** 2 / d ** 2) - np.arcsinh(s * dCt / d) / s)
  return pref * ((dCt / d) * np.sqrt(1 + reg * dCt ** 2 / d ** 2) - np.arcsin(s * dCt / d) / s)


def z_from_dGW(c, dGW):
  """Inverse of dL(z) by linear interpolation, clamped (cosmo.py:260-264)."""
  return np.interp(dGW, dL_at_z(c, c["z_grid_interp"]), c["z_grid_interp"])


# =========================================================================== mass
def make_mass(model="plp", **kw):
  m = dict(MASS_DEFAULT[model])
  m["model"] = model
  for k, v in kw.items():
    if k in m:
      m[k] = v
  return mass_setup(m)


def mass_update(m, **kw):
  """mass.py:35-42."""
  keys = [k for k in kw if k in MASS_DEFAULT[m["model"]]]
  if not keys:
    return m
  new = {k: m[k] for k in MASS_DEFAULT[m["model"]]}
  new["model"] = m["model"]
  for k in keys:
    new[k] = kw[k]
  return mass_setup(new)


def mass_setup(m):
  """Normalisation tables (mass.py:45-52)."""
  m = dict(m)
  g = np.logspace(np.log10(m["m_low"]), np.log10(m["m_high"]), int(m["grid_res"]))
  m["m_grid"] = g
  m["cdf_m2_conditioned"] = cumtrapz(secondary_notnorm(m, g, m["m_high"]), g)
  m["norm_p_m1"] = np_trapz(primary_notnorm(m, g), x=g)
  return m


def tpl_notnorm(x, alpha, lo, hi):
  """mass.py:240-245."""
  x = np.asarray(x, dtype=np.float64)
  with np.errstate(all="ignore"):
    return np.where((lo <= x) & (x <= hi), x ** alpha, 0.)


def tpl_cdf(alpha, lo, x):
  """mass.py:247-252 (the alpha==-1 branch reproduces the reference's sign as written)."""
  if alpha == -1:
    return np.log(lo) - np.log(x)
  return (x ** (1 + alpha) - lo ** (1 + alpha)) / (1 + alpha)


def smoothing(x, delta_m, lo):
  """mass.py:255-264."""
  x = np.asarray(x, dtype=np.float64)
  eps = 1.e-99
  with np.errstate(all="ignore"):
    mid = -np.logaddexp(0.0, (delta_m / (x - lo + eps) + delta_m / (x - lo - delta_m + eps)))
    ls = np.where(x < lo, -np.inf, np.where(x > (lo + delta_m), 0.0, mid))
    return np.exp(ls)


def gaussian(x, mu, sigma):
  """mass.py:267-269."""
  return np.exp(-0.5 * np.log(2 * np.pi) - np.log(sigma) - (x - mu) ** 2 / (2. * sigma ** 2))


def truncated_gaussian(x, mu, sigma, lo, hi):
  """mass.py:271-279."""
  x = np.asarray(x, dtype=np.float64)
  a = (hi - mu) / (sigma * np.sqrt(2.))
  b = (lo - mu) / (sigma * np.sqrt(2.))
  norm = 0.5 * erf(a) - 0.5 * erf(b)
  return np.where((lo <= x) & (x <= hi), gaussian(x, mu, sigma) / norm, 0.)


def primary_notnorm(m, x):
  """mass.py:285-305."""
  x = np.asarray(x, dtype=np.float64)
  lo, hi = m["m_low"], m["m_high"]
  if m["model"] == "tpl":
    return tpl_notnorm(x, -m["alpha"], lo, hi)
  if m["model"] == "bpl":
    mb = lo + m["break_fraction"] * (hi - lo)
    p1b = tpl_notnorm(mb, -m["alpha_1"], lo, mb)
    p2b = tpl_notnorm(mb, -m["alpha_2"], mb, hi)
    pdf = tpl_notnorm(x, -m["alpha_1"], lo, mb)
    pdf = pdf + tpl_notnorm(x, -m["alpha_2"], mb, hi) * p1b / p2b
    return pdf * smoothing(x, m["delta_m"], lo)
  if m["model"] == "plp":
    P = tpl_notnorm(x, -m["alpha"], lo, hi) / tpl_cdf(-m["alpha"], lo, hi)
    G = truncated_gaussian(x, m["mu_g"], m["sigma_g"], lo, m["mu_g"] + 5 * m["sigma_g"])
    pdf = (1 - m["lambda_peak"]) * P + m["lambda_peak"] * G
    return pdf * smoothing(x, m["delta_m"], lo)
  raise ValueError(m["model"])


def secondary_notnorm(m, m2, m1):
  """mass.py:320-328."""
  pdf = tpl_notnorm(m2, m["beta"], m["m_low"], m1)
  if m["model"] != "tpl":
    pdf = pdf * smoothing(m2, m["delta_m"], m["m_low"])
  return pdf


def p_m1m2(m, m1, m2):
  """Joint source-frame mass pdf (mass.py:334-345)."""
  m1 = np.asarray(m1, dtype=np.float64)
  m2 = np.asarray(m2, dtype=np.float64)
  p1 = primary_notnorm(m, m1) / m["norm_p_m1"]
  with np.errstate(all="ignore"):
    p2 = secondary_notnorm(m, m2, m1) / np.interp(m1, m["m_grid"], m["cdf_m2_conditioned"])
  p2 = np.where(np.isnan(p2), 0., p2)
  return p1 * p2


# =========================================================================== rate
def make_rate(model="madau_dickinson", **kw):
  r = dict(RATE_DEFAULT[model])
  r["model"] = model
  for k, v in kw.items():
    if k in r:
      r[k] = v
  return r


def rate_update(r, **kw):
  """rate.py:23-30."""
  new = dict(r)
  for k in kw:
    if k in RATE_DEFAULT[r["model"]]:
      new[k] = kw[k]
  return new


def merger_rate(r, z):
  """rate.py:96-129."""
  z = np.asarray(z, dtype=np.float64)
  mdl = r["model"]
  if mdl == "power_law":
    return (1. + z) ** r["gamma"]
  if mdl == "trunc_power_law":
    norm = ((1 + r["zmax"]) ** (r["gamma"] + 1) - 1) / (r["gamma"] + 1)
    return np.where(z < r["zmax"], (1. + z) ** r["gamma"] / norm, 0.)
  g, k, zp = r["gamma"], r["kappa"], r["zp"]
  md = (1. + z) ** g / (1. + ((1. + z) / (1. + zp)) ** (g + k))
  val = (1. + (1. + zp) ** (-g - k)) * md
  if mdl == "trunc_madau_dickinson":
    return np.where(z < r["zmax"], val, 0.)
  return val


# =========================================================================== population
def make_pop(cosmo, mass, rate, R0=1., catalog=None, Tobs=1., scale_free=True):
  """`population` (pop_wrapper.py:14-43). `catalog`: None (empty, dVc/dz background) or a dict
  with `p_cat (Nev,P,Nz)`, `P_compl (Nev,1,Nz)`, `z_range (2,)` (catalog.py:78-141)."""
  return dict(cosmo=cosmo, mass=mass, rate=rate, R0=R0, catalog=catalog, Tobs=Tobs,
              scale_free=scale_free)


def pop_update(pop, **hl):
  """pop_wrapper.py:56-64."""
  new = dict(pop)
  new["cosmo"] = cosmo_update(pop["cosmo"], **hl)
  new["mass"] = mass_update(pop["mass"], **hl)
  new["rate"] = rate_update(pop["rate"], **hl)
  new["R0"] = hl.get("R0", pop["R0"])
  return new


def theta_det2src(c, m1det, m2det, dL):
  """pop_wrapper.py:67-75."""
  z = z_from_dGW(c, dL)
  return m1det / (1. + z), m2det / (1. + z), z


def src_and_weights(pop, ev):
  """pop_wrapper.py:77-80.  `ev`: dict with m1det, m2det, dL, pe_prior of shape (Nev, Ns)."""
  m1s, m2s, z = theta_det2src(pop["cosmo"], ev["m1det"], ev["m2det"], ev["dL"])
  return z, p_m1m2(pop["mass"], m1s, m2s) / ev["pe_prior"]


def completeness_P(z_range, zgrids):
  """dVdz_completeness.P_compl, kind='step' (completeness.py:43-46)."""
  return np.where((zgrids > z_range[0]) & (zgrids < z_range[1]), 1., 0.)


def completeness_fR(c, z_range):
  """dVdz_completeness.fR (completeness.py:54-58)."""
  v = Vc_at_z(c, np.asarray(z_range, dtype=np.float64))
  return v[1] - v[0]


def p_gal(pop, z):
  """empty_catalog.p_gal (catalog.py:40-43) / pixelated_catalog.p_gal (catalog.py:197-203)."""
  cat = pop["catalog"]
  if cat is None:
    return dVcdz_at_z(pop["cosmo"], z)
  fR = completeness_fR(pop["cosmo"], cat["z_range"])
  p_bkg = dVcdz_at_z(pop["cosmo"], z)[:, None, :]
  pg = fR * cat["p_cat"] + (1. - cat["P_compl"]) * p_bkg
  return np.where(cat["p_cat"] != -100., pg, -100.)


def p_cbc(pop, z):
  """pop_wrapper.py:82-90."""
  pg = p_gal(pop, z)
  p_rate = merger_rate(pop["rate"], z) / (1 + z)
  if pg.ndim > p_rate.ndim:
    return np.where(pg != -100, pg * p_rate[:, None, :], -100)
  return pg * p_rate


def pop_rate_det_inj(pop, inj):
  """Detector-frame rate of the injections (pop_wrapper.py:102-111). `inj`: m1det,m2det,dL."""
  c = pop["cosmo"]
  m1s, m2s, z = theta_det2src(c, inj["m1det"], inj["m2det"], inj["dL"])
  p_z = dVcdz_at_z(c, z, inj["dL"]) * (merger_rate(pop["rate"], z) / (1. + z))
  dN = pop["R0"] * p_m1m2(pop["mass"], m1s, m2s) * p_z
  jac = np.abs(ddLdz_at_z(c, z, inj["dL"])) * (1. + z) ** 2
  return dN / jac


def N_exp(pop, inj, N_inj, N_eff=5.):
  """selection_function.N_exp (selection_function.py:34-48). Returns (Nexp, xi, neff)."""
  with np.errstate(all="ignore"):
    dN = pop_rate_det_inj(pop, inj) / inj["p_draw"]
    xi = np.nansum(dN, axis=-1) / N_inj
    Nexp = pop["Tobs"] * xi
    neff = np.nan
    if N_eff is not None:
      var = np.sum(dN ** 2, axis=-1) / N_inj ** 2 - xi ** 2 / N_inj
      neff = xi ** 2 / var
      if neff < N_eff:
        Nexp = 0.0
  return Nexp, xi, neff


# =========================================================================== KDE numerics
def binning1d(x, w, num_bins=200):
  """utils/math.py:32-46: returns (bin centres, bin sums)."""
  lo, hi = np.min(x), np.max(x)
  edges = np.linspace(lo, hi, num_bins + 1)
  centers = (edges[:-1] + edges[1:]) / 2
  with np.errstate(all="ignore"):
    f = np.clip(np.floor((x - lo) / (hi - lo) * num_bins), 0, num_bins - 1)
  sums = np.zeros(num_bins)
  ok = np.isfinite(f)
  np.add.at(sums, f[ok].astype(np.int64), w[ok])   # NaN indices: out-of-bounds scatter is dropped
  return centers, sums


def kde1d(x, grid, w, kernel="epan", bw_method=None):
  """utils/math.py:52-89."""
  with np.errstate(all="ignore"):
    w = w / np.sum(w)
    neff = 1.0 / np.sum(w ** 2)
    if bw_method == "scott" or bw_method is None:
      bw = neff ** (-1. / 5) * np.std(x)
    elif bw_method == "silverman":
      bw = (neff * 3 / 4.0) ** (-1. / 5) * np.std(x)
    elif np.isscalar(bw_method) and not isinstance(bw_method, str):
      bw = bw_method * np.std(x)
    else:
      raise ValueError("bw_method should be 'scott', 'silverman', or a scalar")
    u = (grid[:, None] - x) / bw
    if kernel == "epan":
      kv = np.where(np.abs(u) <= 1, 3 / 4 * (1 - u ** 2), 0)
    else:
      kv = np.exp(-0.5 * u ** 2) / np.sqrt(2 * np.pi)
    return np.sum(w * kv, axis=-1) / bw


def gkde_nd(dataset, points, weights=None, bw_method=None):
  """n-D weighted Gaussian KDE (utils/math.py:154-229; same algorithm as jax_gkde_nd :95-148
  and scipy.stats.gaussian_kde).  dataset (d,N), points (d,M) -> (M,)."""
  dataset = np.atleast_2d(dataset)
  d, n = dataset.shape
  points = np.atleast_2d(points)
  w = np.full(n, 1.0 / n) if weights is None else weights / np.sum(weights)
  neff = 1.0 / np.sum(w ** 2)
  if bw_method == "scott" or bw_method is None:
    factor = neff ** (-1. / (d + 4))
  elif bw_method == "silverman":
    factor = (neff * (d + 2) / 4.0) ** (-1. / (d + 4))
  elif np.isscalar(bw_method) and not isinstance(bw_method, str):
    factor = bw_method
  else:
    raise ValueError("`bw_method` should be 'scott', 'silverman', a scalar")
  mean = np.sum(w * dataset, axis=1)
  res = dataset - mean[:, None]
  cov = np.atleast_2d(np.dot(res * w, res.T)) / (1 - np.sum(w ** 2))
  inv_cov = np.linalg.inv(cov) / factor ** 2
  L = np.linalg.cholesky(inv_cov)
  pw = points.T @ L
  dw = dataset.T @ L
  log_norm = np.sum(np.log(np.diag(L))) - 0.5 * d * np.log(2 * np.pi)
  out = np.empty(pw.shape[0])
  chunk = max(1, int(4e6 // max(n, 1)))
  for s in range(0, pw.shape[0], chunk):
    d2 = np.sum((pw[s:s + chunk, None, :] - dw[None, :, :]) ** 2, axis=-1)
    out[s:s + chunk] = np.sum(w * np.exp(log_norm - 0.5 * d2), axis=1)
  return out


# =========================================================================== likelihood
def make_opts(kind_p_gw3d=None, kernel="epan", bw_method=None, cut_grid=2.0, binning=True,
              num_bins=200, pe_neff=2.0):
  """hyperlikelihood constructor options (likelihood.py:48-62)."""
  return dict(kind_p_gw3d=kind_p_gw3d, kernel=kernel, bw_method=bw_method, cut_grid=cut_grid,
              binning=binning, num_bins=num_bins, pe_neff=pe_neff)


def p_gw1d(pop, ev, z_grids, opts):
  """likelihood.py:105-144 -> (Nev, Nz)."""
  nev, nz = z_grids.shape
  with np.errstate(all="ignore"):
    z, w = src_and_weights(pop, ev)
    norms = np.mean(w, axis=-1)
    n_effs = np.sum(w, axis=-1) ** 2 / np.sum(w ** 2, axis=-1)
    out = np.zeros((nev, nz))
    for e in range(nev):
      if not (n_effs[e] >= opts["pe_neff"]):
        continue
      ze, we = z[e], w[e]
      if opts["cut_grid"] is not None:
        zmin, zmax, sig = np.min(ze), np.max(ze), np.std(ze)
        lb = zmin - opts["cut_grid"] * sig if (zmin - opts["cut_grid"] * sig > 0.) else 1.e-8
        ub = zmax + opts["cut_grid"] * sig
        eff = np.linspace(lb, ub, nz // 2)
      else:
        eff = z_grids[e]
      if opts["binning"]:
        ze, we = binning1d(ze, we, opts["num_bins"])
      dens = kde1d(ze, eff, we, opts["kernel"], opts["bw_method"]) * norms[e]
      out[e] = np.interp(z_grids[e], eff, dens, left=0., right=0.)
  return out


def p_gw3dapprox(pop, ev, z_grids, opts):
  """likelihood.py:150-154 -> (Nev, P, Nz)."""
  return p_gw1d(pop, ev, z_grids, opts)[:, None, :] * ev["gw_loc2d_pdf"][:, :, None]


def p_gw3dmarg(pop, ev, z_grids, opts):
  """likelihood.py:160-205 -> (Nev, P, Nz).  kde1d is called without `kernel=` there, so the
  kernel is always Epanechnikov; eff-grid uses the unmasked z statistics."""
  nev, nz = z_grids.shape
  P = ev["pixels_opt_nsides"].shape[1]
  out = np.zeros((nev, P, nz))
  with np.errstate(all="ignore"):
    z, w = src_and_weights(pop, ev)
    norms = np.mean(w, axis=-1)
    n_effs = np.sum(w, axis=-1) ** 2 / np.sum(w ** 2, axis=-1)
    for e in range(nev):
      if not (n_effs[e] >= opts["pe_neff"]):
        continue
      ze, we = z[e], w[e]
      for i in range(P):
        mask = ev["pixels_pe_opt_nside"][e] == ev["pixels_opt_nsides"][e, i]
        zm = np.where(mask, ze, np.min(ze))
        wm = np.where(mask, we, 0.0)
        zp, wp = binning1d(zm, wm, opts["num_bins"]) if opts["binning"] else (zm, wm)
        if opts["cut_grid"] is not None:
          lo = max(np.min(ze) - opts["cut_grid"] * np.std(ze), 1e-8)
          hi = np.max(ze) + opts["cut_grid"] * np.std(ze)
          eff = np.linspace(lo, hi, nz // 2)
        else:
          eff = z_grids[e]
        dens = kde1d(zp, eff, wp, "epan", opts["bw_method"])
        out[e, i] = np.interp(z_grids[e], eff, dens, left=0., right=0.) * norms[e] * ev["gw_loc2d_pdf"][e, i]
  return out


def p_gw3dfull(pop, ev, z_grids, opts, neff_pixels):
  """likelihood.py:211-260 -> (Nev, P, Nz)."""
  nev, nz = z_grids.shape
  P = ev["pixels_opt_nsides"].shape[1]
  out = np.zeros((nev, P, nz))
  with np.errstate(all="ignore"):
    z, w = src_and_weights(pop, ev)
    norms = np.mean(w, axis=-1)
    n_effs = np.sum(w, axis=-1) ** 2 / np.sum(w ** 2, axis=-1)
    for e in range(nev):
      if n_effs[e] < opts["pe_neff"]:
        continue
      zs, zmax, zmin = np.std(z[e]), np.max(z[e]), np.min(z[e])
      zmask = (z_grids[e] <= zmax + opts["cut_grid"] * zs) & (z_grids[e] >= zmin - opts["cut_grid"] * zs)
      zeff = z_grids[e][zmask]
      npix = int(neff_pixels[e])
      if zeff.size == 0 or npix == 0:
        continue
      pts = np.array([np.tile(zeff, npix),
                      np.repeat(ev["ra_pix"][e, :npix], zeff.size),
                      np.repeat(ev["dec_pix"][e, :npix], zeff.size)])
      data = np.array([z[e], ev["ra"][e], ev["dec"][e]])
      vals = gkde_nd(data, pts, weights=w[e], bw_method=opts["bw_method"])
      blk = np.zeros((npix, nz))
      blk[:, zmask] = vals.reshape(npix, zeff.size)
      out[e, :npix, :] = blk * norms[e]
  return out


def numlike_evs(pop, ev, z_grids, opts, neff_pixels=None):
  """likelihood.py:266-292 -> (Nev,)."""
  with np.errstate(all="ignore"):
    jac = ddLdz_at_z(pop["cosmo"], z_grids) * (1. + z_grids) ** 2
    pz = p_cbc(pop, z_grids)
    kind = opts["kind_p_gw3d"]
    if kind is None:
      pgw = p_gw1d(pop, ev, z_grids, opts)
      return np_trapz(pgw * pz / jac, z_grids, axis=-1)
    if kind == "approximate":
      pgw = p_gw3dapprox(pop, ev, z_grids, opts)
    elif kind == "marginalized":
      pgw = p_gw3dmarg(pop, ev, z_grids, opts)
    elif kind == "full":
      pgw = p_gw3dfull(pop, ev, z_grids, opts, neff_pixels)
    else:
      raise AssertionError("`kind_p_gw3d` must be one of `approximate`, `marginalized`, or `full`")
    integrand = np.where(pz != -100, pgw * pz / jac[:, None, :], 0.0)
    return np.sum(np_trapz(integrand, z_grids[:, None, :], axis=-1), axis=-1)


def compute_all(pop0, ev, z_grids, opts, inj, N_inj, N_eff=5., neff_pixels=None, **hyper):
  """hyperlikelihood.compute_all (likelihood.py:326-338) for ONE hyper-point.
  Returns (log_like_evs (Nev,), log_like_num, log N_exp, log_hyper)."""
  pop = pop_update(pop0, **hyper)
  nev = z_grids.shape[0]
  with np.errstate(all="ignore"):
    lle = np.nan_to_num(np.log(numlike_evs(pop, ev, z_grids, opts, neff_pixels)), nan=-np.inf)
    lnum = np.sum(lle, axis=-1)
    nexp, _, _ = N_exp(pop, inj, N_inj, N_eff)
    if not pop["scale_free"]:
      lnum = lnum + nev * np.log(pop["R0"] * pop["Tobs"])
      lh = lnum - nexp
    else:
      lh = lnum - nev * np.log(nexp)
    return lle, lnum, np.log(nexp), lh


# =========================================================================== setup helpers
def compute_z_grids(cosmo, dL, cosmo_prior=None, z_int_res=300, z_conf_range=None):
  """pop_wrapper.py:133-208."""
  if isinstance(z_conf_range, list):
    dL_min, dL_max = np.percentile(dL, z_conf_range, axis=1)
  elif isinstance(z_conf_range, (int, float)):
    mu, sig = np.mean(dL, axis=1), np.std(dL, axis=1)
    dL_min, dL_max = mu - z_conf_range * sig, mu + z_conf_range * sig
  else:
    dL_max = np.max(dL, axis=1) * 2
    dL_min = np.min(dL, axis=1) * 0.5
    dL_min = np.where(dL_min < 1.e-8, 1.e-8, dL_min)
  names = ["H0", "Om0", "Ok0", "Or0", "w0", "wa"] + (["Xi0", "n"] if cosmo["model"] == "mg_flrw" else [])
  cp = {k: [cosmo[k], cosmo[k]] for k in names}
  if cosmo_prior is not None:
    cp.update(cosmo_prior)
  lo = {k: cp[k][0] for k in names[:6]}
  hi = {k: cp[k][1] for k in names[:6]}
  if cosmo["model"] == "mg_flrw":
    lo.update(Xi0=cp["Xi0"][1], n=cp["n"][1])
    hi.update(Xi0=cp["Xi0"][0], n=cp["n"][1])
  c1 = cosmo_update(cosmo, z_grid_res=10_000, **lo)
  c2 = cosmo_update(cosmo, z_grid_res=10_000, **hi)
  z_min = z_from_dGW(c1, dL_min)
  z_max = z_from_dGW(c2, dL_max)
  return np.linspace(z_min, z_max, z_int_res, axis=1)


def sum_gaussians_ucv(z_grid, mu, sigma, cosmo, weights=None):
  """catalog.py:209-221."""
  if len(mu) == 0:
    return np.zeros_like(z_grid)
  if weights is None:
    weights = np.ones(len(mu))
  zg = z_grid[:, None]
  with np.errstate(all="ignore"):
    g = np.power(2 * np.pi * (sigma ** 2), -0.5) * np.exp(-0.5 * np.power((zg - mu) / sigma, 2.))
    g = g * dVcdz_at_z(cosmo, zg)
    norm = np_trapz(g, zg, axis=0)
    return np.sum(weights * g / norm, axis=1) / np.sum(weights)


def precompute_p_cat(cosmo, gal, opt_nsides, pixels_opt_nsides, z_grids, gal_pix_by_nside):
  """pixelated_catalog.precompute_p_cat (catalog.py:143-195).
  gal: dict z, z_err, w; gal_pix_by_nside: {nside: int64 pixel of every galaxy}.
  Returns (p_cat (Nev,P,Nz) padded with -100, N_gal (Nev,))."""
  nev, P = pixels_opt_nsides.shape
  nz = z_grids.shape[1]
  p_cat = np.full((nev, P, nz), -100.)
  ngal = np.zeros(nev)
  for e in range(nev):
    pixs = pixels_opt_nsides[e]
    good = pixs[pixs != -100]
    gpix = gal_pix_by_nside[int(opt_nsides[e])]
    sel = np.isin(gpix, good)
    zsel, esel, wsel, psel = gal["z"][sel], gal["z_err"][sel], gal["w"][sel], gpix[sel]
    mz = (zsel > z_grids[e][0]) & (zsel < z_grids[e][-1])
    zsel, esel, wsel, psel = zsel[mz], esel[mz], wsel[mz], psel[mz]
    for i, p in enumerate(good):
      m = psel == p
      row = sum_gaussians_ucv(z_grids[e], zsel[m], esel[m], cosmo, wsel[m])
      row[~np.isfinite(row)] = 0.
      p_cat[e, i] = row
      ngal[e] += np.count_nonzero(m)
  return p_cat, ngal
