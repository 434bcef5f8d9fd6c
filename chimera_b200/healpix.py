"""HEALPix RING-scheme pixel indexing (host side, NumPy, int64-exact).

The reference calls healpy (an un-vendored third-party dependency, `pyproject.toml:27`,
`healpy>=1.14`, unpinned) at `CHIMERA/utils/angles.py:45` (`hp.ang2pix(nside, pi/2-dec, ra,
nest=False)`), `:71` (`hp.pix2ang`) and `CHIMERA/data.py:258` (`hp.nside2npix`).  healpy is
absent offline, so this module restates the published RING algorithm (Gorski et al. 2005,
ApJ 622, 759; the `loc2pix` / `pix2loc` routines of healpix_cxx that healpy wraps).  Only the
RING scheme is implemented because the reference never passes `nest=True` on the hot path.

Pixel ids feed `pixels_pe_opt_nside == pixels_opt_nsides[i]` comparisons
(`CHIMERA/likelihood.py:176`) and the galaxy bucketing (`CHIMERA/catalog/catalog.py:146`), so
they must be bit-exact integers: all index arithmetic below is int64, the floating part is
IEEE double with the same operation order as healpix_cxx.
"""
import numpy as np

_TWOTHIRD = 2.0 / 3.0
_HALFPI = 0.5 * np.pi
_INV_HALFPI = 2.0 / np.pi


def nside2npix(nside):
  nside = int(nside)
  return 12 * nside * nside


def _check_nside(nside):
  nside = int(nside)
  if nside < 1 or (nside & (nside - 1)) != 0:
    raise ValueError("nside must be a positive power of 2")
  return nside


def ang2pix(nside, theta, phi, nest=False):
  """RING pixel index of colatitude `theta` / longitude `phi` [rad] (healpy.ang2pix)."""
  if nest:
    raise NotImplementedError("only the RING scheme is used by the likelihood path")
  nside = _check_nside(nside)
  theta = np.asarray(theta, dtype=np.float64)
  phi = np.asarray(phi, dtype=np.float64)
  theta, phi = np.broadcast_arrays(theta, phi)
  if np.any((theta < 0.0) | (theta > np.pi)):
    raise ValueError("theta out of range [0, pi]")
  shape = theta.shape
  theta = theta.ravel()
  phi = phi.ravel()

  z = np.cos(theta)
  have_sth = (theta < 0.01) | (theta > 3.14159 - 0.01)
  sth = np.where(have_sth, np.sin(theta), 0.0)
  za = np.abs(z)
  # fmodulo(phi*2/pi, 4) in [0,4)
  tt = np.mod(phi * _INV_HALFPI, 4.0)
  tt = np.where(tt >= 4.0, 0.0, tt)

  nl4 = np.int64(4 * nside)
  ncap = np.int64(2 * nside * (nside - 1))
  npix = np.int64(12 * nside * nside)
  pix = np.empty(theta.shape, dtype=np.int64)

  eq = za <= _TWOTHIRD
  # equatorial belt
  if np.any(eq):
    tte = tt[eq]
    ze = z[eq]
    temp1 = nside * (0.5 + tte)
    temp2 = nside * ze * 0.75
    jp = (temp1 - temp2).astype(np.int64)   # ascending edge line (truncation, values >= 0)
    jm = (temp1 + temp2).astype(np.int64)   # descending edge line
    ir = nside + 1 + jp - jm                # ring counted from z = 2/3, in [1, 2n+1]
    kshift = 1 - (ir & 1)
    t1 = jp + jm - nside + kshift + 1 + nl4 + nl4
    ip = (t1 >> 1) & (nl4 - 1)
    pix[eq] = ncap + (ir - 1) * nl4 + ip
  # polar caps
  cap = ~eq
  if np.any(cap):
    ttc = tt[cap]
    zc = z[cap]
    zac = za[cap]
    tp = ttc - np.floor(ttc)
    use_sqrt = (zac < 0.99) | (~have_sth[cap])
    with np.errstate(invalid="ignore"):
      tmp = np.where(use_sqrt, nside * np.sqrt(3.0 * (1.0 - zac)),
                     nside * sth[cap] / np.sqrt((1.0 + zac) / 3.0))
    jp = (tp * tmp).astype(np.int64)
    jm = ((1.0 - tp) * tmp).astype(np.int64)
    ir = jp + jm + 1                        # ring counted from the closest pole
    ip = (ttc * ir).astype(np.int64)
    ip = np.minimum(ip, 4 * ir - 1)
    pix[cap] = np.where(zc > 0, 2 * ir * (ir - 1) + ip, npix - 2 * ir * (ir + 1) + ip)
  return pix.reshape(shape) if shape else np.int64(pix[0])


def _isqrt(v):
  """Exact integer square root of a non-negative int64 array."""
  v = np.asarray(v, dtype=np.int64)
  r = np.floor(np.sqrt(v.astype(np.float64) + 0.5)).astype(np.int64)
  r = np.where(r * r > v, r - 1, r)
  r = np.where((r + 1) * (r + 1) <= v, r + 1, r)
  return r


def pix2ang(nside, ipix, nest=False):
  """(theta, phi) [rad] of RING pixel centres (healpy.pix2ang)."""
  if nest:
    raise NotImplementedError("only the RING scheme is used by the likelihood path")
  nside = _check_nside(nside)
  ipix = np.asarray(ipix, dtype=np.int64)
  shape = ipix.shape
  pix = ipix.ravel()
  npix = np.int64(12 * nside * nside)
  if np.any((pix < 0) | (pix >= npix)):
    raise ValueError("pixel index out of range")
  ncap = np.int64(2 * nside * (nside - 1))
  nl4 = np.int64(4 * nside)
  fact2 = 4.0 / float(npix)
  fact1 = (2 * nside) * fact2

  z = np.empty(pix.shape, dtype=np.float64)
  phi = np.empty(pix.shape, dtype=np.float64)
  sth = np.zeros(pix.shape, dtype=np.float64)
  have_sth = np.zeros(pix.shape, dtype=bool)

  north = pix < ncap
  south = pix >= (npix - ncap)
  belt = ~(north | south)
  if np.any(north):
    p = pix[north]
    iring = (1 + _isqrt(1 + 2 * p)) >> 1
    iphi = (p + 1) - 2 * iring * (iring - 1)
    tmp = (iring * iring) * fact2
    zz = 1.0 - tmp
    hs = zz > 0.99
    z[north] = zz
    sth[north] = np.where(hs, np.sqrt(tmp * (2.0 - tmp)), 0.0)
    have_sth[north] = hs
    phi[north] = (iphi - 0.5) * _HALFPI / iring
  if np.any(belt):
    p = pix[belt] - ncap
    tmp = p // nl4
    iring = tmp + nside
    iphi = p - nl4 * tmp + 1
    fodd = np.where(((iring + nside) & 1) == 1, 1.0, 0.5)
    z[belt] = (2 * nside - iring) * fact1
    phi[belt] = (iphi - fodd) * np.pi * 0.75 * fact1
  if np.any(south):
    p = npix - pix[south]
    iring = (1 + _isqrt(2 * p - 1)) >> 1
    iphi = 4 * iring + 1 - (p - 2 * iring * (iring - 1))
    tmp = (iring * iring) * fact2
    zz = tmp - 1.0
    hs = zz < -0.99
    z[south] = zz
    sth[south] = np.where(hs, np.sqrt(tmp * (2.0 - tmp)), 0.0)
    have_sth[south] = hs
    phi[south] = (iphi - 0.5) * _HALFPI / iring
  theta = np.where(have_sth, np.arctan2(sth, z), np.arccos(np.clip(z, -1.0, 1.0)))
  if shape:
    return theta.reshape(shape), phi.reshape(shape)
  return np.float64(theta[0]), np.float64(phi[0])


def find_pix_RAdec(ra, dec, nside, nest=False):
  """Pixel of (RA, dec) [rad]; mirrors `CHIMERA/utils/angles.py:32-45`."""
  return ang2pix(nside, 0.5 * np.pi - np.asarray(dec), np.asarray(ra), nest=nest)


def find_ra_dec(pix, nside, nest=False):
  """(RA, dec) [rad] of pixel centres; mirrors `CHIMERA/utils/angles.py:73-85`."""
  theta, phi = pix2ang(nside, pix, nest=nest)
  return phi, 0.5 * np.pi - theta
