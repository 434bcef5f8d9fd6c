"""Seeded synthetic workloads of the shapes BASELINE.json names.

Input generation only -- nothing here is on the likelihood path, and nothing here imports the
oracle.  It carries its own small flat-LCDM helper (dense-grid quadrature) so that the inputs do
not depend on either implementation under test.  The arrays follow the reference's input
contract: `theta_pe_det` fields (`CHIMERA/data.py:27-47`), `theta_inj_det` (`:49-53`), the
pixelised-catalogue layout padded with -100 (`CHIMERA/data.py:343-351`), `p_cat (Nev,P,Nz)`
(`CHIMERA/catalog/catalog.py:152-195`) and `z_grids (Nev,Nz)` (`pop_wrapper.py:133-208`).
"""
import numpy as np
from . import healpix

_trapz = np.trapezoid if hasattr(np, "trapezoid") else np.trapz
_C = 299792.458e-3  # Gpc * (km/s/Mpc)


class _FlatLCDM:
  """Dense-grid flat LCDM distances (Gpc); independent of the library/oracle tables."""

  def __init__(self, H0=70., Om0=0.25, zmax=12., n=200_001):
    self.H0, self.Om0 = H0, Om0
    self.dH = _C / H0
    self.z = np.linspace(0., zmax, n)
    E = np.sqrt(Om0 * (1 + self.z) ** 3 + (1 - Om0))
    dz = self.z[1] - self.z[0]
    self.dC = self.dH * np.concatenate([[0.], np.cumsum(0.5 * (1 / E[1:] + 1 / E[:-1]) * dz)])
    self.dLtab = self.dC * (1 + self.z)

  def E(self, z):
    return np.sqrt(self.Om0 * (1 + z) ** 3 + (1 - self.Om0))

  def dL(self, z):
    return np.interp(z, self.z, self.dLtab)

  def z_of_dL(self, dL):
    return np.interp(dL, self.dLtab, self.z)

  def dVcdz(self, z):
    dC = np.interp(z, self.z, self.dC)
    return 4 * np.pi * self.dH * dC ** 2 / self.E(z)

  def ddLdz(self, z):
    dC = np.interp(z, self.z, self.dC)
    return dC + self.dH * (1 + z) / self.E(z)


def _md_rate(z, gamma=2.7, kappa=3.0, zp=2.0):
  return (1 + z) ** gamma / (1 + ((1 + z) / (1 + zp)) ** (gamma + kappa))


def _sample_from_grid(rng, x, pdf, size):
  cdf = np.concatenate([[0.], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(x))])
  cdf /= cdf[-1]
  return np.interp(rng.random(size), cdf, x)


def make_events(nev, ns, seed=1234, sky=False, zmax_true=1.5):
  """PE samples of `nev` events x `ns` samples. Returns dict of (nev, ns) f64 arrays."""
  rng = np.random.default_rng(seed)
  cos = _FlatLCDM()
  zg = np.linspace(0.02, zmax_true, 4000)
  z_true = _sample_from_grid(rng, zg, cos.dVcdz(zg) * _md_rate(zg) / (1 + zg), nev)
  # primary: power law (-3.4) on [6.5, 80] with a 4% Gaussian peak; secondary: m2^1.1 on [5.5, m1]
  a = -3.4
  lo, hi = 6.5, 80.
  u = rng.random(nev)
  m1 = (u * (hi ** (a + 1) - lo ** (a + 1)) + lo ** (a + 1)) ** (1 / (a + 1))
  peak = rng.random(nev) < 0.04
  m1 = np.where(peak, np.clip(rng.normal(34., 3.6, nev), lo, hi), m1)
  b = 1.1
  lo2 = 5.5
  u = rng.random(nev)
  m2 = (u * (m1 ** (b + 1) - lo2 ** (b + 1)) + lo2 ** (b + 1)) ** (1 / (b + 1))
  dL_true = cos.dL(z_true)
  sig_d = rng.uniform(0.05, 0.25, nev)
  dL_obs = dL_true * np.exp(sig_d * rng.standard_normal(nev))
  dL = dL_obs[:, None] * np.exp(sig_d[:, None] * rng.standard_normal((nev, ns)))
  m1d_c = m1 * (1 + z_true)
  m2d_c = m2 * (1 + z_true)
  m1d = m1d_c[:, None] * (1 + 0.05 * rng.standard_normal((nev, ns)))
  m2d = m2d_c[:, None] * (1 + 0.05 * rng.standard_normal((nev, ns)))
  m1det = np.maximum(m1d, m2d)
  m2det = np.maximum(np.minimum(m1d, m2d), 0.5)
  out = dict(m1det=m1det, m2det=m2det, dL=dL, pe_prior=dL ** 2, z_true=z_true)
  if sky:
    # centre uniform on the sphere (kept 12 deg away from the poles), Gaussian blob around it
    ra0 = rng.uniform(0, 2 * np.pi, nev)
    dec0 = np.arcsin(rng.uniform(-0.95, 0.95, nev))
    sig = np.deg2rad(rng.uniform(0.5, 3.0, nev))
    dec = dec0[:, None] + sig[:, None] * rng.standard_normal((nev, ns))
    dec = np.clip(dec, -0.5 * np.pi + 1e-6, 0.5 * np.pi - 1e-6)
    ra = ra0[:, None] + sig[:, None] * rng.standard_normal((nev, ns)) / np.cos(dec0)[:, None]
    out["ra"] = np.mod(ra, 2 * np.pi)
    out["dec"] = dec
    out["ra_true"], out["dec_true"] = ra0, dec0
  return out


def make_injections(ninj_det, seed=5678, snr_thr=None, zmax=2.5, z_scale=None):
  """Detected injections with analytic detector-frame `p_draw`.
  Returns (dict m1det, m2det, dL, p_draw of shape (ninj_det,), N_inj drawn).
  `z_scale`: when given, the redshift proposal is tapered by exp(-z / z_scale) (a campaign that does not waste draws
  where nothing is detectable: ~10x fewer draws per detection); `p_draw` is the density actually drawn from."""
  rng = np.random.default_rng(seed)
  cos = _FlatLCDM()
  zg = np.linspace(1e-4, zmax, 8000)
  pz = cos.dVcdz(zg) * (1 + zg)
  if z_scale is not None:
    pz = pz * np.exp(-zg / z_scale)
  pz_norm = _trapz(pz, zg)
  lo, hi = 2., 120.
  kept = {k: [] for k in ("m1det", "m2det", "dL", "p_draw")}
  ndrawn, nkept = 0, 0
  batch = max(4 * ninj_det, 100_000)
  thr = 9.0 if snr_thr is None else snr_thr
  while nkept < ninj_det:
    u = rng.random(batch)
    m1 = 1.0 / (1.0 / lo - u * (1.0 / lo - 1.0 / hi))          # p(m1) ~ m1^-2
    m2 = rng.uniform(lo, m1)
    z = _sample_from_grid(rng, zg, pz, batch)
    dL = cos.dL(z)
    mc = (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2 * (1 + z)
    snr = 8.0 * mc ** (5. / 6.) / dL * rng.lognormal(0., 0.25, batch) / 6.0
    det = snr > thr
    p_m1 = m1 ** -2. / (1.0 / lo - 1.0 / hi)
    p_m2 = 1.0 / (m1 - lo)
    p_z = np.interp(z, zg, pz) / pz_norm
    p_draw = p_m1 * p_m2 * p_z / ((1 + z) ** 2 * cos.ddLdz(z))
    idx = np.flatnonzero(det)
    need = ninj_det - nkept
    if idx.size >= need:
      idx = idx[:need]
      ndrawn += int(idx[-1]) + 1
    else:
      ndrawn += batch
    kept["m1det"].append(m1[idx] * (1 + z[idx]))
    kept["m2det"].append(m2[idx] * (1 + z[idx]))
    kept["dL"].append(dL[idx])
    kept["p_draw"].append(p_draw[idx])
    nkept += idx.size
  return {k: np.concatenate(v) for k, v in kept.items()}, int(ndrawn)


def make_galaxies(ngal, seed=9012, zmax=1.4, z_err=0.001):
  rng = np.random.default_rng(seed)
  cos = _FlatLCDM()
  zg = np.linspace(1e-4, zmax, 4000)
  z = _sample_from_grid(rng, zg, cos.dVcdz(zg), ngal)
  ra = rng.uniform(0, 2 * np.pi, ngal)
  dec = np.arcsin(rng.uniform(-1, 1, ngal))
  return dict(ra=ra, dec=dec, z=z, z_err=z_err * (1 + z), w=np.ones(ngal))


def make_z_grids(dL, z_int_res=300, H0_prior=(20., 200.), Om0=0.25):
  """Default-branch z grids (pop_wrapper.py:159-162,202-206): [z(0.5 min dL; H0_lo), z(2 max dL; H0_hi)]."""
  c1 = _FlatLCDM(H0_prior[0], Om0)
  c2 = _FlatLCDM(H0_prior[1], Om0)
  dmin = np.maximum(np.min(dL, axis=1) * 0.5, 1e-8)
  dmax = np.max(dL, axis=1) * 2
  zmin = c1.z_of_dL(dmin)
  zmax = np.minimum(c2.z_of_dL(dmax), 9.5)
  return np.linspace(zmin, zmax, z_int_res, axis=1)


def _threshold(p, level):
  ps = np.sort(p)[::-1]
  idx = np.searchsorted(np.cumsum(ps), level)
  return ps[min(idx, ps.size - 1)]


def pixelize(ev, nside_list=(512, 256, 128, 64, 32, 16, 8), mean_npixels_event=15, sky_conf=0.9):
  """Host-side pixelisation of the sky samples, following the procedure of
  `CHIMERA/data.py:262-392` (per-event optimal nside, 90% area pixels, nearest-valid-pixel
  assignment, 2-D KDE at pixel centres, -100 padding). Adds the pixel fields to a copy of `ev`."""
  from scipy.stats import gaussian_kde
  ra, dec = ev["ra"], ev["dec"]
  nev, ns = ra.shape
  pix_all = {ns_: healpix.find_pix_RAdec(ra, dec, ns_) for ns_ in nside_list}

  def conf_pixels(pe_pix):
    uniq, counts = np.unique(pe_pix, return_counts=True)
    p = counts / pe_pix.shape[0]
    return uniq[p >= _threshold(p, sky_conf)]

  counts = np.array([[conf_pixels(pix_all[n_][e]).size for n_ in nside_list] for e in range(nev)])
  best = np.argmin(np.abs(counts - mean_npixels_event), axis=1)
  opt_nsides = np.asarray(nside_list)[best]
  ev_pix = [conf_pixels(pix_all[int(opt_nsides[e])][e]) for e in range(nev)]
  P = max(p.size for p in ev_pix)
  pixels = np.full((nev, P), -100, dtype=np.int64)
  ra_pix = np.full((nev, P), -100.)
  dec_pix = np.full((nev, P), -100.)
  pdf = np.full((nev, P), -100.)
  pe_pix = np.zeros((nev, ns), dtype=np.int64)
  for e in range(nev):
    npx = ev_pix[e].size
    r, d = healpix.find_ra_dec(ev_pix[e], int(opt_nsides[e]))
    pixels[e, :npx], ra_pix[e, :npx], dec_pix[e, :npx] = ev_pix[e], r, d
    sp = pix_all[int(opt_nsides[e])][e]
    valid = np.isin(sp, ev_pix[e])
    cosang = (np.sin(dec[e])[:, None] * np.sin(d)[None, :]
              + np.cos(dec[e])[:, None] * np.cos(d)[None, :] * np.cos(ra[e][:, None] - r[None, :]))
    closest = np.argmin(np.arccos(np.clip(cosang, -1, 1)), axis=1)
    pe_pix[e] = np.where(valid, sp, ev_pix[e][closest])
    pdf[e, :npx] = gaussian_kde(np.array([ra[e], dec[e]]))(np.array([r, d]))
  out = dict(ev)
  out.update(opt_nsides=opt_nsides, pixels_opt_nsides=pixels, ra_pix=ra_pix, dec_pix=dec_pix,
             gw_loc2d_pdf=pdf, pixels_pe_opt_nside=pe_pix,
             neff_pixels=np.array([p.size for p in ev_pix], dtype=np.int32))
  return out


def make_p_cat(gal, pix_ev, z_grids, z_range=(0.073, 1.3)):
  """Catalogue term on the event grids: per pixel sum of galaxy Gaussians x dVc/dz, each
  normalised on the grid (the `sumgauss='dVdz'` recipe of `catalog.py:209-221`); -100 padding.
  Returns (p_cat (Nev,P,Nz), P_compl (Nev,1,Nz))."""
  cos = _FlatLCDM()
  nev, P = pix_ev["pixels_opt_nsides"].shape
  nz = z_grids.shape[1]
  p_cat = np.full((nev, P, nz), -100.)
  nsides = np.unique(pix_ev["opt_nsides"])
  gpix = {int(n_): healpix.find_pix_RAdec(gal["ra"], gal["dec"], int(n_)) for n_ in nsides}
  order = {n_: np.argsort(gpix[n_], kind="stable") for n_ in gpix}
  sorted_pix = {n_: gpix[n_][order[n_]] for n_ in gpix}
  for e in range(nev):
    n_ = int(pix_ev["opt_nsides"][e])
    zg = z_grids[e]
    dv = cos.dVcdz(zg)
    for i in range(int(pix_ev["neff_pixels"][e])):
      p = pix_ev["pixels_opt_nsides"][e, i]
      a, b = np.searchsorted(sorted_pix[n_], [p, p + 1])
      idx = order[n_][a:b]
      zgal, egal, wgal = gal["z"][idx], gal["z_err"][idx], gal["w"][idx]
      m = (zgal > zg[0]) & (zgal < zg[-1])
      zgal, egal, wgal = zgal[m], egal[m], wgal[m]
      if zgal.size == 0:
        p_cat[e, i] = 0.
        continue
      g = np.exp(-0.5 * ((zg[:, None] - zgal) / egal) ** 2) / np.sqrt(2 * np.pi * egal ** 2) * dv[:, None]
      with np.errstate(all="ignore"):
        norm = _trapz(g, zg[:, None], axis=0)
        row = np.sum(wgal * g / norm, axis=1) / np.sum(wgal)
      row[~np.isfinite(row)] = 0.
      p_cat[e, i] = row
  P_compl = np.where((z_grids > z_range[0]) & (z_grids < z_range[1]), 1., 0.)[:, None, :]
  return p_cat, P_compl


def smooth_p_cat(pix_ev, z_grids, seed=777, z_range=(0.073, 1.3)):
  """Cheap stand-in catalogue term for very large configs: a few random Gaussian 'clusters' per
  pixel times dVc/dz, normalised per pixel -- same layout/sentinels as `make_p_cat`."""
  rng = np.random.default_rng(seed)
  cos = _FlatLCDM()
  nev, P = pix_ev["pixels_opt_nsides"].shape
  nz = z_grids.shape[1]
  p_cat = np.full((nev, P, nz), -100.)
  for e in range(nev):
    zg = z_grids[e]
    npx = int(pix_ev["neff_pixels"][e])
    mu = rng.uniform(zg[0], zg[-1], (npx, 6, 1))
    sg = rng.uniform(0.01, 0.05, (npx, 6, 1)) * (zg[-1] - zg[0])
    amp = rng.random((npx, 6, 1))
    prof = np.sum(amp * np.exp(-0.5 * ((zg[None, None, :] - mu) / sg) ** 2), axis=1) * cos.dVcdz(zg)[None, :]
    prof /= np.maximum(_trapz(prof, zg[None, :], axis=1)[:, None], 1e-300)
    p_cat[e, :npx] = prof
  P_compl = np.where((z_grids > z_range[0]) & (z_grids < z_range[1]), 1., 0.)[:, None, :]
  return p_cat, P_compl
