"""Catalogue prior objects with the surface of CHIMERA/catalog/catalog.py:
`empty_catalog` (:19-43) and `pixelated_catalog` (:51-203): `.p_gal(cosmo, z)`, `.p_bkg`, `.fR`,
`.p_cat (Nev,P,Nz)`, `.P_compl (Nev,1,Nz)`, `.N_gal`, `.max_npixels`, `.neff_pixels`.

In the likelihood the assembly `fR * p_cat + (1 - P_compl) * p_bkg` is fused into the CUDA
numerator kernel (csrc/numerator.cu); `p_gal` below is the inspection entry point.
`precompute_p_cat` (setup, once per run) runs on the GPU through `chb_precompute_p_cat`."""
import numpy as np
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
    pixel i of event e with z strictly inside the event grid (catalog.py:143-231); -100 padding.
    Runs on the GPU (`chb_precompute_p_cat`, csrc/setup.cu): HEALPix ids of the galaxies at every nside in
    use, bucketing by pixel, then one CTA per (event, pixel)."""
    import ctypes as C
    from .. import _lib
    zgrids = _lib.f64(zgrids)
    nsides = _lib.i64(self.data_gw_pixelated.opt_nsides)
    pixels = _lib.i64(self.data_gw_pixelated.pixels_opt_nsides)
    nev, P = pixels.shape
    nz = zgrids.shape[1]
    g = self.data_gal
    dv = _lib.f64(np.asarray(dVcdz_at_z(self.cosmo, zgrids)).reshape(nev, nz))     # fiducial cosmology, on the GPU
    gra, gdec, gz, ge, gw = (_lib.f64(g[k]) for k in ("ra", "dec", "z", "z_err", "w"))
    neff = np.ascontiguousarray(self.neff_pixels, dtype=np.int32)
    p_cat = np.empty((nev, P, nz), dtype=np.float64)
    N_gal = np.zeros(nev, dtype=np.float64)
    _lib.check(_lib.load().chb_precompute_p_cat(
      int(getattr(self, "device", 0)), nev, P, nz, _lib.dptr(zgrids), _lib.dptr(dv), _lib.iptr(nsides), _lib.iptr(pixels),
      neff.ctypes.data_as(C.POINTER(C.c_int32)), gz.size, _lib.dptr(gra), _lib.dptr(gdec), _lib.dptr(gz), _lib.dptr(ge),
      _lib.dptr(gw), _lib.dptr(p_cat), _lib.dptr(N_gal)))
    self.p_cat = p_cat
    self.N_gal = N_gal
    self.P_compl = self.completeness.P_compl(zgrids)[:, np.newaxis, :]

  def p_gal(self, cosmo_lambdas, z):
    """fR * p_cat + (1 - P_compl) * p_bkg with the -100 sentinel preserved (catalog.py:197-203)."""
    fR = self.fR(cosmo_lambdas)
    p_bkg = np.asarray(self.p_bkg(cosmo_lambdas, z))[:, np.newaxis, :]
    p_gal = fR * self.p_cat + (1. - self.P_compl) * p_bkg
    return np.where(self.p_cat != -100., p_gal, -100.)
