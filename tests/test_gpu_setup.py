"""GPU parity tests of the setup-side callers of the likelihood path, through
the C ABI: HEALPix RING indexing and the pixelisation of a GW catalogue (bit-exact integers), the 2-D
localisation KDE, and pixelated_catalog.precompute_p_cat -- against the fixtures written by the reference's own
code (tests/golden) and against the NumPy restatements."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
  import chimera_b200
  from chimera_b200 import _lib
  if _lib.device_count() == 0:
    pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
  return chimera_b200


def _directions(rng, n):
  th = np.arccos(rng.uniform(-1, 1, n))
  ph = rng.uniform(-2 * np.pi, 4 * np.pi, n)          # longitudes outside [0, 2pi) wrap like healpy's fmodulo
  # edge cases: poles, equator, cap boundaries z = +-2/3, phi on the base-pixel meridians
  th_e = np.array([0.0, np.pi, 0.5 * np.pi, np.arccos(2. / 3.), np.arccos(-2. / 3.), 1e-9, np.pi - 1e-9, 0.0099, 0.0101,
                   np.pi - 0.0099, 3.14159 - 0.01])
  ph_e = np.array([0.0, 0.5 * np.pi, np.pi, 1.5 * np.pi, 2 * np.pi, np.nextafter(2 * np.pi, 0), -0.0, 1e-300, 0.25 * np.pi])
  T, Pm = np.meshgrid(th_e, ph_e, indexing="ij")
  return np.concatenate([th, T.ravel()]), np.concatenate([ph, Pm.ravel()])


@pytest.mark.parametrize("nside", [1, 2, 8, 64, 512, 4096])
def test_ang2pix_bit_exact_vs_host(cb, nside):
  from chimera_b200 import healpix as hp_host
  th, ph = _directions(np.random.default_rng(nside), 1_000_000)
  got = cb.sky.ang2pix(nside, th, ph)
  ref = hp_host.ang2pix(nside, th, ph)
  assert got.dtype == np.int64
  np.testing.assert_array_equal(got, ref)
  ra, dec = ph, 0.5 * np.pi - th
  np.testing.assert_array_equal(cb.sky.find_pix_RAdec(ra, dec, nside), hp_host.find_pix_RAdec(ra, dec, nside))


@pytest.mark.parametrize("nside", [1, 4, 64, 1024])
def test_pix2ang_and_roundtrip(cb, nside):
  from chimera_b200 import healpix as hp_host
  npix = cb.sky.nside2npix(nside)
  pix = np.arange(npix) if npix <= 200_000 else np.random.default_rng(3).integers(0, npix, 200_000)
  th, ph = cb.sky.pix2ang(nside, pix)
  th0, ph0 = hp_host.pix2ang(nside, pix)
  np.testing.assert_allclose(th, th0, rtol=0, atol=1e-15)       # CUDA acos vs libm acos: <= 2 ulp
  np.testing.assert_allclose(ph, ph0, rtol=0, atol=1e-15)
  np.testing.assert_array_equal(cb.sky.ang2pix(nside, th, ph), pix)        # centres map back to their pixel
  ra, dec = cb.sky.find_ra_dec(pix, nside)
  np.testing.assert_array_equal(cb.sky.find_pix_RAdec(ra, dec, nside), pix)


def test_healpix_errors(cb):
  with pytest.raises(ValueError):
    cb.sky.ang2pix(3, 0.1, 0.1)
  with pytest.raises(ValueError):
    cb.sky.ang2pix(8, -0.1, 0.1)
  with pytest.raises(ValueError):
    cb.sky.pix2ang(8, 12 * 64)
  with pytest.raises(NotImplementedError):
    cb.sky.ang2pix(8, 0.1, 0.1, nest=True)


def test_pixelize_gw_catalog_matches_reference(cb, golden_setup):
  """The fixture holds the output of the reference's own pixelize_gw_catalog (data.py:262-392) on these samples."""
  g = golden_setup
  th = cb.theta_pe_det(dL=g["dL"], ra=g["ra"], dec=g["dec"])
  out = cb.pixelize_gw_catalog(th, nside_list=[64, 32, 16, 8], mean_npixels_event=6, sky_conf=0.9)
  np.testing.assert_array_equal(out.opt_nsides, g["pix_opt_nsides"])
  np.testing.assert_array_equal(out.pixels_opt_nsides, g["pix_pixels"])
  np.testing.assert_array_equal(out.pixels_pe_opt_nside, g["pix_pe"])
  np.testing.assert_allclose(out.ra_pix, g["pix_ra"], rtol=1e-15)
  np.testing.assert_allclose(out.dec_pix, g["pix_dec"], rtol=1e-15, atol=1e-15)
  np.testing.assert_allclose(out.gw_loc2d_pdf, g["pix_pdf"], rtol=1e-10)


def test_pixelize_large_matches_host(cb):
  """Larger seeded case against the NumPy restatement (synth.pixelize): ids bit-exact."""
  from chimera_b200 import synth
  ev = synth.make_events(40, 3000, seed=77, sky=True)
  ref = synth.pixelize(ev, nside_list=(256, 128, 64, 32, 16, 8), mean_npixels_event=12, sky_conf=0.9)
  th = cb.theta_pe_det(dL=ev["dL"], ra=ev["ra"], dec=ev["dec"])
  out = cb.pixelize_gw_catalog(th, nside_list=[256, 128, 64, 32, 16, 8], mean_npixels_event=12, sky_conf=0.9)
  np.testing.assert_array_equal(out.opt_nsides, ref["opt_nsides"])
  np.testing.assert_array_equal(out.pixels_opt_nsides, ref["pixels_opt_nsides"])
  np.testing.assert_array_equal(out.pixels_pe_opt_nside, ref["pixels_pe_opt_nside"])
  np.testing.assert_allclose(out.gw_loc2d_pdf, ref["gw_loc2d_pdf"], rtol=1e-9)


def test_precompute_p_cat_matches_reference(cb, golden_inpix):
  """p_cat / N_gal / P_compl of the reference's pixelated_catalog constructor (catalog.py:78-195) on the same
  galaxies, pixels and z grids."""
  g = golden_inpix
  th = cb.theta_pe_det(**{k: g[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides",
                                            "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf",
                                            "pixels_pe_opt_nside")})
  fid = cb.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  gcat = cb.pixelated_catalog(cb.dVdz_completeness(g["z_range"]), cosmo=fid, z_grids=g["z_grids"],
                              data_gw_pixelated=th, data_gal=dict(ra=g["gal_ra"], dec=g["gal_dec"], z=g["gal_z"]),
                              z_err=0.001)
  np.testing.assert_allclose(gcat.p_cat, g["p_cat"], rtol=1e-10, atol=1e-300)
  np.testing.assert_array_equal(gcat.N_gal, g["N_gal"])
  np.testing.assert_array_equal(gcat.P_compl, g["P_compl"])
  np.testing.assert_array_equal(gcat.neff_pixels, g["neff_pixels"])


def test_precompute_p_cat_large_matches_oracle(cb):
  """C4-like shape at reduced size: 2e5 galaxies, nside up to 64, weights, events whose grids exclude galaxies,
  pixels without galaxies, and a galaxy error so small that its Gaussian underflows on the grid (row -> 0,
  catalog.py:191)."""
  from oracle import chimera_oracle as orc
  from chimera_b200 import synth, healpix as hp_host
  ev = synth.make_events(24, 800, seed=5, sky=True)
  zg = synth.make_z_grids(ev["dL"], z_int_res=96, H0_prior=(40., 120.))
  ev = synth.pixelize(ev, nside_list=(64, 32, 16, 8), mean_npixels_event=10)
  gal = synth.make_galaxies(200_000, seed=6)
  rng = np.random.default_rng(8)
  w = rng.uniform(0.2, 3.0, gal["z"].size)
  z_err = 0.001
  gal_zerr = z_err * (1 + gal["z"])
  th = cb.theta_pe_det(**{k: ev[k] for k in ("dL", "ra", "dec", "opt_nsides", "pixels_opt_nsides", "ra_pix", "dec_pix",
                                             "gw_loc2d_pdf", "pixels_pe_opt_nside")})
  fid = cb.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  gcat = cb.pixelated_catalog(cb.dVdz_completeness([0.073, 1.3]), cosmo=fid, z_grids=zg, data_gw_pixelated=th,
                              data_gal=dict(ra=gal["ra"], dec=gal["dec"], z=gal["z"]), z_err=z_err, weights=w)
  fid0 = orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.)
  gpix = {int(n): hp_host.find_pix_RAdec(gal["ra"], gal["dec"], int(n)) for n in np.unique(ev["opt_nsides"])}
  ref, ngal = orc.precompute_p_cat(fid0, dict(z=gal["z"], z_err=gal_zerr, w=w), ev["opt_nsides"], ev["pixels_opt_nsides"],
                                   zg, gpix)
  np.testing.assert_array_equal(gcat.p_cat == -100., ref == -100.)
  np.testing.assert_allclose(gcat.p_cat, ref, rtol=1e-9, atol=1e-300)
  np.testing.assert_array_equal(gcat.N_gal, ngal)
  assert np.any(ref == 0.0)
