"""Catalogue prior objects with the surface of CHIMERA/catalog/catalog.py:
`empty_catalog` (:19-43) and `pixelated_catalog` (:51-203): `.p_gal(cosmo, z)`, `.p_bkg`, `.fR`,
`.p_cat (Nev,P,Nz)`, `.P_compl (Nev,1,Nz)`, `.N_gal`, `.max_npixels`, `.neff_pixels`.

In the likelihood the assembly `fR * p_cat + (1 - P_compl) * p_bkg` is fused into the CUDA
numerator kernel (csrc/numerator.cu); `p_gal` below is the inspection entry point.
`precompute_p_cat` (setup, once per run) currently runs on the host."""
import numpy as np
from .. import healpix
from ..population.cosmo import dVcdz_at_z

_trapz = np.trapezoid if hasattr(np, "trapezoid") else np.trapz


class empty_catalog(object):
  """No catalogue: p_gal = background dVc/dz (spectral sirens), catalog.py:19-43."""

  def __init__(self, p_bkg="dVdz"):
    self.p_cat = 0.
    self.N_gal = 0.
    self.P_compl = 0.
    if p_bkg != "dVdz":
      raise ValueError("only p_bkg='dVdz' is implemented in the CUDA kernels")
    self.p_bkg = dVcdz_at_z

  def p_gal(self, cosmo_lambdas, z):
    return self.p_bkg(cosmo_lambdas, z)


class pixelated_catalog(object):
  """Pixelated galaxy catalogue (catalog.py:51-203).

  Either pass precomputed arrays (`p_cat`, `P_compl` -- e.g. loaded from a reference cache
  file) or galaxy data (`data_gal` dict with ra, dec, z [rad]) plus the pixelated GW catalogue
  and the event z-grids, in which case `precompute_p_cat` builds them."""

  def __init__(self, completeness, cosmo=None, z_grids=None, data_gw_pixelated=None, data_gal=None,
               z_err=1, weights=None, mask_gal=None, sumgauss="dVdz", p_cat=None, P_compl=None, N_gal=None,
               neff_pixels=None):
    self.completeness = completeness
    self.p_bkg = completeness.p_bkg
    self.fR = completeness.fR
    if sumgauss != "dVdz":
      raise ValueError("only sumgauss='dVdz' is implemented")
    if p_cat is not None:
      self.p_cat = np.ascontiguousarray(p_cat, dtype=np.float64)
      if self.p_cat.ndim != 3:
        raise ValueError("p_cat must have shape (Nev, max_npixels, Nz)")
      self.P_compl = np.asarray(P_compl, dtype=np.float64).reshape(self.p_cat.shape[0], 1, self.p_cat.shape[2])
      self.N_gal = N_gal
      self.max_npixels = self.p_cat.shape[1]
      if neff_pixels is None:
        neff_pixels = np.sum(self.p_cat[:, :, 0] != -100., axis=1)
      self.neff_pixels = np.asarray(neff_pixels, dtype=np.int32)
      return
    if data_gal is None or data_gw_pixelated is None or z_grids is None or cosmo is None:
      raise ValueError("need either p_cat/P_compl or cosmo + z_grids + data_gw_pixelated + data_gal")
    self.cosmo = cosmo
    self.z_grids = np.asarray(z_grids, dtype=np.float64)
    self.data_gw_pixelated = data_gw_pixelated
    self.z_err = z_err
    self.data_gal = {k: np.asarray(v) for k, v in data_gal.items() if k in ("ra", "dec", "z")}
    self.data_gal['w'] = np.asarray(weights, dtype=np.float64) if weights is not None else np.ones_like(self.data_gal['z'])
    self.data_gal['z_err'] = self.z_err * (1. + self.data_gal['z'])
    if mask_gal is not None:
      m = np.asarray(mask_gal)
      self.data_gal = {k: v[m] for k, v in self.data_gal.items()}
    ra_pix = np.asarray(data_gw_pixelated.ra_pix)
    self.nevents = ra_pix.shape[0]
    self.max_npixels = ra_pix.shape[1]
    self.neff_pixels = np.sum(ra_pix != -100., axis=1).astype(np.int32)       # catalog.py:121
    self.precompute_p_cat(self.z_grids)

  def precompute_p_cat(self, zgrids):
    """p_cat[e, i, :] = sum_g w_g N(z; z_g, s_g) dVc/dz(z) / norm_g / sum_g w_g over the galaxies in
    pixel i of event e with z strictly inside the event grid (catalog.py:143-231); -100 padding."""
    zgrids = np.asarray(zgrids, dtype=np.float64)
    nsides = np.asarray(self.data_gw_pixelated.opt_nsides)
    pixels = np.asarray(self.data_gw_pixelated.pixels_opt_nsides)
    nev, P = pixels.shape
    nz = zgrids.shape[1]
    g = self.data_gal
    order, spix = {}, {}
    for ns in np.unique(nsides):
      gp = healpix.find_pix_RAdec(g['ra'], g['dec'], int(ns))
      order[int(ns)] = np.argsort(gp, kind="stable")
      spix[int(ns)] = gp[order[int(ns)]]
    p_cat = np.full((nev, P, nz), -100.)
    N_gal = np.zeros(nev)
    for e in range(nev):
      ns = int(nsides[e])
      zg = zgrids[e]
      dv = np.asarray(dVcdz_at_z(self.cosmo, zg))
      for i in range(int(self.neff_pixels[e])):
        a, b = np.searchsorted(spix[ns], [pixels[e, i], pixels[e, i] + 1])
        idx = order[ns][a:b]
        zgal, sg, wg = g['z'][idx], g['z_err'][idx], g['w'][idx]
        m = (zgal > zg[0]) & (zgal < zg[-1])
        zgal, sg, wg = zgal[m], sg[m], wg[m]
        N_gal[e] += zgal.size
        if zgal.size == 0:
          p_cat[e, i] = 0.
          continue
        with np.errstate(all="ignore"):
          gauss = np.power(2 * np.pi * sg ** 2, -0.5) * np.exp(-0.5 * ((zg[:, None] - zgal) / sg) ** 2) * dv[:, None]
          norm = _trapz(gauss, zg[:, None], axis=0)
          row = np.sum(wg * gauss / norm, axis=1) / np.sum(wg)
        row[~np.isfinite(row)] = 0.
        p_cat[e, i] = row
    self.p_cat = p_cat
    self.N_gal = N_gal
    self.P_compl = self.completeness.P_compl(zgrids)[:, np.newaxis, :]

  def p_gal(self, cosmo_lambdas, z):
    """fR * p_cat + (1 - P_compl) * p_bkg with the -100 sentinel preserved (catalog.py:197-203)."""
    fR = self.fR(cosmo_lambdas)
    p_bkg = np.asarray(self.p_bkg(cosmo_lambdas, z))[:, np.newaxis, :]
    p_gal = fR * self.p_cat + (1. - self.P_compl) * p_bkg
    return np.where(self.p_cat != -100., p_gal, -100.)
