"""Sky pixelisation on the GPU: the callers on the input side of the likelihood path.

Mirrors, name for name, the functions the reference takes from healpy and `CHIMERA/utils/angles.py`
(`find_pix_RAdec` :32-45, `find_pix` :47-59, `find_theta_phi` :61-71, `find_ra_dec` :73-85) and
`pixelize_gw_catalog` (`CHIMERA/data.py:262-392`).  All pixel arithmetic runs in the CUDA kernels of
`csrc/setup.cu` through the C ABI (`chb_healpix_*`, `chb_pixelize_samples`); there is no CPU fallback
(`chimera_b200.healpix` is the NumPy restatement used to generate synthetic inputs and to check these kernels).
Only the RING scheme exists because the reference never passes nest=True on this path."""
import numpy as np

from . import _lib
from .data import theta_pe_det


def _device():
  import os
  return int(os.environ.get("LOCAL_RANK", "0")) if _lib.device_count() > 1 else 0


def nside2npix(nside):
  return 12 * int(nside) * int(nside)


def ang2pix(nside, theta, phi, nest=False):
  """healpy.ang2pix(nside, theta, phi, nest=False): RING pixel ids (int64), computed on the GPU."""
  if nest:
    raise NotImplementedError("only the RING scheme is used by the likelihood path")
  theta, phi = np.broadcast_arrays(np.asarray(theta, dtype=np.float64), np.asarray(phi, dtype=np.float64))
  shape = theta.shape
  t, p = _lib.f64(theta.ravel()), _lib.f64(phi.ravel())
  out = np.empty(t.size, dtype=np.int64)
  _lib.check(_lib.load().chb_healpix_ang2pix_ring(_device(), int(nside), t.size, _lib.dptr(t), _lib.dptr(p), _lib.iptr(out)))
  return out.reshape(shape) if shape else np.int64(out[0])


def pix2ang(nside, ipix, nest=False):
  """healpy.pix2ang(nside, ipix, nest=False): (theta, phi) of RING pixel centres, computed on the GPU."""
  if nest:
    raise NotImplementedError("only the RING scheme is used by the likelihood path")
  ipix = np.asarray(ipix, dtype=np.int64)
  shape = ipix.shape
  px = _lib.i64(ipix.ravel())
  th, ph = np.empty(px.size), np.empty(px.size)
  _lib.check(_lib.load().chb_healpix_pix2ang_ring(_device(), int(nside), px.size, _lib.iptr(px), _lib.dptr(th), _lib.dptr(ph)))
  if shape:
    return th.reshape(shape), ph.reshape(shape)
  return np.float64(th[0]), np.float64(ph[0])


def find_pix(theta, phi, nside, nest=False):
  return ang2pix(nside, theta, phi, nest=nest)


def find_pix_RAdec(ra, dec, nside, nest=False):
  """utils/angles.py:32-45."""
  return ang2pix(nside, 0.5 * np.pi - np.asarray(dec, dtype=np.float64), np.asarray(ra, dtype=np.float64), nest=nest)


def find_theta_phi(pix, nside, nest=False):
  return pix2ang(nside, pix, nest=nest)


def find_ra_dec(pix, nside, nest=False):
  """utils/angles.py:73-85."""
  theta, phi = pix2ang(nside, pix, nest=nest)
  return phi, 0.5 * np.pi - theta


def _get_threshold(norm_counts, level):
  """data.py:239-244."""
  prob_sorted = np.sort(norm_counts)[::-1]
  idx = np.searchsorted(np.cumsum(prob_sorted), level)
  return prob_sorted[idx]


def compute_sky_conf_event(healpix_pe, sky_conf, nside):
  """data.py:246-260: pixels whose sample fraction reaches the `sky_conf` credible level.  The reference fills a
  dense map of 12 nside^2 fractions and sorts it; pixels without samples hold 0 and sort last, so the threshold
  (`_get_threshold`) only ever depends on the occupied pixels: the same index is found on the sparse counts."""
  unique, counts = np.unique(healpix_pe, return_counts=True)
  p = counts / healpix_pe.shape[0]
  prob_sorted = np.sort(p)[::-1]
  idx = np.searchsorted(np.cumsum(prob_sorted), sky_conf)
  if idx >= prob_sorted.size:          # the dense map would index a pixel without samples (or run off its end)
    return _compute_sky_conf_event_dense(healpix_pe, sky_conf, nside)
  return unique[p >= prob_sorted[idx]]


def _compute_sky_conf_event_dense(healpix_pe, sky_conf, nside):
  """The reference's literal procedure (dense map), kept for the degenerate levels and as the checker."""
  unique, counts = np.unique(healpix_pe, return_counts=True)
  p = np.zeros(nside2npix(nside))
  p[unique] = counts / healpix_pe.shape[0]
  return np.argwhere(p >= _get_threshold(p, sky_conf)).flatten()


def _pad(arrs, pad_value):
  n = max(len(a) for a in arrs)
  out = np.full((len(arrs), n), pad_value, dtype=np.asarray(arrs[0]).dtype)
  for i, a in enumerate(arrs):
    out[i, :len(a)] = a
  return out


def pixelize_gw_catalog(theta_gw, nside_list, mean_npixels_event, sky_conf, nest=False):
  """`CHIMERA/data.py:262-392` with the per-sample work on the GPU: HEALPix ids of every (ra, dec) sample at
  every nside, nearest-valid-pixel assignment and the 2-D KDE `gw_loc2d_pdf` at the pixel centres.  The
  per-event bookkeeping (credible-region pixel sets, optimal nside) stays on the host, as in the reference.
  Returns a `theta_pe_det` with the pixelisation fields filled."""
  if nest:
    raise NotImplementedError("only the RING scheme is used by the likelihood path")
  ra = np.ascontiguousarray(theta_gw.ra, dtype=np.float64)
  dec = np.ascontiguousarray(theta_gw.dec, dtype=np.float64)
  nev, ns = ra.shape
  pix_all = {f"nside_{n}": find_pix_RAdec(ra, dec, n) for n in nside_list}
  counts = np.array([[len(compute_sky_conf_event(pix_all[f"nside_{n}"][e], sky_conf, n)) for n in nside_list]
                     for e in range(nev)])
  best = np.argmin(np.abs(counts - mean_npixels_event), axis=1)
  opt_nsides = np.asarray(nside_list, dtype=np.int64)[best]
  event_pixels = [compute_sky_conf_event(pix_all[f"nside_{opt_nsides[e]}"][e], sky_conf, int(opt_nsides[e]))
                  for e in range(nev)]
  pixel_ra, pixel_dec = zip(*[find_ra_dec(event_pixels[e], int(opt_nsides[e])) for e in range(nev)])
  pixels = _pad([p.astype(np.int64) for p in event_pixels], -100)
  ra_pix, dec_pix = _pad(pixel_ra, -100.), _pad(pixel_dec, -100.)
  P = pixels.shape[1]
  pe_pix = np.empty((nev, ns), dtype=np.int64)
  pdf = np.empty((nev, P), dtype=np.float64)
  _lib.check(_lib.load().chb_pixelize_samples(_device(), nev, ns, P, _lib.iptr(opt_nsides), _lib.dptr(ra), _lib.dptr(dec),
                                               _lib.iptr(pixels), _lib.dptr(ra_pix), _lib.dptr(dec_pix),
                                               _lib.iptr(pe_pix), _lib.dptr(pdf)))
  fields = {k: getattr(theta_gw, k, None) for k in ("m1det", "m2det", "dL", "pe_prior")}
  return theta_pe_det(ra=ra, dec=dec, pixels_pe_all_nsides=pix_all, opt_nsides=opt_nsides, pixels_opt_nsides=pixels,
                      ra_pix=ra_pix, dec_pix=dec_pix, gw_loc2d_pdf=pdf, pixels_pe_opt_nside=pe_pix,
                      **{k: v for k, v in fields.items() if v is not None})
