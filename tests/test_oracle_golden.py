"""Pins oracle/chimera_oracle.py to the golden fixtures produced by the unmodified reference
source (tests/golden/make_golden.py).  Tolerances: 1e-12 relative (same NumPy arithmetic, the
only differences are operation order) unless stated."""
import numpy as np
import pytest
from oracle import chimera_oracle as orc
from cases import LIKE_CASES2, SEL_BPL_MG_HYPERS, COSMO_CASES, MASS_CASES, RATE_CASES, LIKE_CASES, GOLDEN_NUM_BINS

RTOL = 1e-12


def close(a, b, rtol=RTOL, atol=0.0):
  np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=True)


@pytest.mark.parametrize("i", range(len(COSMO_CASES)))
def test_cosmology(golden_models, i):
  g = golden_models
  model, kw = COSMO_CASES[i]
  c = orc.make_cosmo(model, **kw)
  z = g["z"]
  close(np.stack([c["z_grid_interp"], c["integral_invE_interp"]]), g[f"cosmo{i}_tab"])
  close(orc.E_at_z(c, z), g[f"cosmo{i}_E"])
  close(orc.dL_at_z(c, z), g[f"cosmo{i}_dL"])
  close(orc.ddLdz_at_z(c, z), g[f"cosmo{i}_ddL"])
  close(orc.dVcdz_at_z(c, z), g[f"cosmo{i}_dV"])
  close(orc.Vc_at_z(c, z), g[f"cosmo{i}_Vc"], atol=1e-40)
  dq = g[f"cosmo{i}_dq"]
  zq = orc.z_from_dGW(c, dq)
  close(zq, g[f"cosmo{i}_zq"])
  close(orc.ddLdz_at_z(c, zq, dq), g[f"cosmo{i}_ddL_dist"])
  close(orc.dVcdz_at_z(c, zq, dq), g[f"cosmo{i}_dV_dist"])


@pytest.mark.parametrize("i", range(len(MASS_CASES)))
def test_mass(golden_models, i):
  g = golden_models
  model, kw = MASS_CASES[i]
  m = orc.make_mass(model, **kw)
  close(m["norm_p_m1"], g[f"mass{i}_norm"])
  close(m["cdf_m2_conditioned"], g[f"mass{i}_cdf"])
  close(orc.primary_notnorm(m, g["m1"]), g[f"mass{i}_p1"])
  close(orc.p_m1m2(m, g["m1"], g["m2"]), g[f"mass{i}_p"])


@pytest.mark.parametrize("i", range(len(RATE_CASES)))
def test_rate(golden_models, i):
  model, kw = RATE_CASES[i]
  close(orc.merger_rate(orc.make_rate(model, **kw), golden_models["zr"]), golden_models[f"rate{i}"])


def test_binning_and_kde1d(golden_math):
  g = golden_math
  c, s = orc.binning1d(g["x"], g["w"], 50)
  close(c, g["bin_centers"])
  close(s, g["bin_sums"])
  for kern in ("epan", "gauss"):
    for j, bw in enumerate((None, "silverman", 0.37)):
      close(orc.kde1d(g["x"], g["grid"], g["w"], kern, bw), g[f"kde_{kern}_{j}"], rtol=1e-11)
  close(orc.kde1d(c, g["grid"], s, "epan", None), g["kde_binned_epan"], rtol=1e-11)


def test_gkde_nd(golden_math):
  g = golden_math
  close(orc.gkde_nd(g["data3"], g["pts3"], g["w"], None), g["gkde3_numba"], rtol=1e-10)
  close(orc.gkde_nd(g["data3"], g["pts3"], g["w"], "silverman"), g["gkde3_numba_silv"], rtol=1e-10)
  close(orc.gkde_nd(g["data3"][:2], g["pts3"][:2]), g["gkde2_jax"], rtol=1e-10)


def test_gkde_nd_vs_scipy(golden_math):
  """Independent check: scipy.stats.gaussian_kde is the algorithm the reference says it copies
  (utils/math.py:96)."""
  from scipy.stats import gaussian_kde
  g = golden_math
  ref = gaussian_kde(g["data3"], weights=g["w"])(g["pts3"])
  close(orc.gkde_nd(g["data3"], g["pts3"], g["w"], None), ref, rtol=1e-9)


def _inputs(g, pix):
  ev = {k: g[k] for k in ("m1det", "m2det", "dL", "pe_prior")}
  if pix:
    for k in ("ra", "dec", "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside"):
      ev[k] = g[k]
  inj = dict(m1det=g["inj_m1det"], m2det=g["inj_m2det"], dL=g["inj_dL"], p_draw=g["inj_p_draw"])
  return ev, g["z_grids"], inj, float(g["N_inj"])


def same_class(a, b, rtol):
  """Compare log-likelihood entries by class (finite / -DBL_MAX / -inf), SURVEY section 8e."""
  a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
  fin = np.isfinite(a) & (np.abs(a) < 1e300)
  assert np.array_equal(fin, np.isfinite(b) & (np.abs(b) < 1e300))
  np.testing.assert_allclose(a[fin], b[fin], rtol=rtol)
  assert np.array_equal(a[~fin], b[~fin])


@pytest.mark.parametrize("name", list(LIKE_CASES))
def test_likelihood(golden_like, golden_in1d, golden_inpix, name):
  kind, kernel, binning, cmodel, hypers = LIKE_CASES[name]
  pix = kind is not None
  g = golden_inpix if pix else golden_in1d
  ev, zg, inj, N_inj = _inputs(g, pix)
  cat = dict(p_cat=g["p_cat"], P_compl=g["P_compl"], z_range=g["z_range"]) if pix else None
  pop0 = orc.make_pop(orc.make_cosmo(cmodel, H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                      orc.make_rate("madau_dickinson"), catalog=cat)
  opts = orc.make_opts(kind, kernel, None, 2.0, binning, GOLDEN_NUM_BINS, 2.0)
  npx = g["neff_pixels"] if pix else None
  for h, hl in enumerate(hypers):
    lle, lnum, lnexp, lh = orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., npx, **hl)
    same_class(lle, golden_like[f"{name}_h{h}_lle"], 1e-10)
    same_class([lnum, lnexp, lh], golden_like[f"{name}_h{h}_tot"], 1e-10)
    pop = orc.pop_update(pop0, **hl)
    if kind is None:
      pgw = orc.p_gw1d(pop, ev, zg, opts)
    elif kind == "approximate":
      pgw = orc.p_gw3dapprox(pop, ev, zg, opts)
    elif kind == "marginalized":
      pgw = orc.p_gw3dmarg(pop, ev, zg, opts)
    else:
      pgw = orc.p_gw3dfull(pop, ev, zg, opts, npx)
    scale = np.nanmax(np.abs(golden_like[f"{name}_h{h}_pgw"]))
    close(pgw, golden_like[f"{name}_h{h}_pgw"], rtol=1e-9, atol=1e-12 * scale)


@pytest.mark.parametrize("name", list(LIKE_CASES2))
def test_likelihood_model_matrix(golden_like2, golden_in1d, golden_inpix, name):
  """tpl / bpl masses, power-law and truncated rates, silverman / scalar bandwidths, curved and w0-wa cosmologies,
  mg_flrw with a catalogue: the oracle against the reference's own outputs (tests/golden/make_golden.py --like2)."""
  c = LIKE_CASES2[name]
  pix = c["kind"] is not None
  g = golden_inpix if pix else golden_in1d
  ev, zg, inj, N_inj = _inputs(g, pix)
  cat = dict(p_cat=g["p_cat"], P_compl=g["P_compl"], z_range=g["z_range"]) if pix else None
  pop0 = orc.make_pop(orc.make_cosmo(c["cosmo"][0], H0=70., Om0=0.25, z_max=5., **c["cosmo"][1]),
                      orc.make_mass(c["mass"][0], **c["mass"][1]), orc.make_rate(c["rate"][0], **c["rate"][1]), catalog=cat)
  opts = orc.make_opts(c["kind"], c["kernel"], c["bw"], 2.0, c["binning"], GOLDEN_NUM_BINS, 2.0)
  npx = g["neff_pixels"] if pix else None
  for h, hl in enumerate(c["hypers"]):
    lle, lnum, lnexp, lh = orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., npx, **hl)
    same_class(lle, golden_like2[f"{name}_h{h}_lle"], 1e-10)
    same_class([lnum, lnexp, lh], golden_like2[f"{name}_h{h}_tot"], 1e-10)


def test_selection_bpl_mg(golden_like2, golden_in1d):
  _, _, inj, N_inj = _inputs(golden_in1d, False)
  pop0 = orc.make_pop(orc.make_cosmo("mg_flrw", H0=70., Om0=0.25, z_max=5.), orc.make_mass("bpl"),
                      orc.make_rate("trunc_madau_dickinson", zmax=2.0))
  got = [orc.N_exp(orc.pop_update(pop0, **hl), inj, N_inj, 5.)[0] for hl in SEL_BPL_MG_HYPERS]
  close(got, golden_like2["sel_bpl_mg_nexp"], rtol=1e-11)


def test_not_scale_free_and_neff_gate(golden_like, golden_in1d):
  ev, zg, inj, N_inj = _inputs(golden_in1d, False)
  opts = orc.make_opts(None, "epan", None, 2.0, True, GOLDEN_NUM_BINS, 2.0)
  c = orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.)
  pop0 = orc.make_pop(c, orc.make_mass("plp"), orc.make_rate("madau_dickinson"), R0=17., Tobs=2.5, scale_free=False)
  out = orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., None, H0=68., R0=21.)
  close(out[1:], golden_like["notscalefree_tot"], rtol=1e-10)
  pop0 = orc.make_pop(c, orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  out = orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 1e9, None, H0=68.)
  assert np.array_equal(np.asarray(out[1:], dtype=np.float64)[1:], golden_like["neffgate_tot"][1:])  # -inf, +inf


def test_compute_z_grids(golden_setup):
  g = golden_setup
  fid = orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.)
  close(orc.compute_z_grids(fid, g["dL"], {"H0": [40., 120.]}, 40), g["zgrid_default"])
  close(orc.compute_z_grids(fid, g["dL"], {"H0": [40., 120.], "Om0": [0.2, 0.4]}, 40, 3.), g["zgrid_sigma"])
  close(orc.compute_z_grids(fid, g["dL"], None, 40, [1., 99.]), g["zgrid_pct"])
  mg = orc.make_cosmo("mg_flrw", H0=70., Om0=0.25, z_max=5.)
  close(orc.compute_z_grids(mg, g["dL"], {"H0": [50., 90.], "Xi0": [0.5, 2.], "n": [1., 3.]}, 40), g["zgrid_mg"])


def test_precompute_p_cat(golden_inpix):
  from chimera_b200 import healpix
  g = golden_inpix
  fid = orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.)
  gal = dict(z=g["gal_z"], z_err=0.001 * (1 + g["gal_z"]), w=np.ones_like(g["gal_z"]))
  gpix = {int(n): healpix.find_pix_RAdec(g["gal_ra"], g["gal_dec"], int(n)) for n in np.unique(g["opt_nsides"])}
  p_cat, ngal = orc.precompute_p_cat(fid, gal, g["opt_nsides"], g["pixels_opt_nsides"], g["z_grids"], gpix)
  close(p_cat, g["p_cat"], rtol=1e-10, atol=1e-300)
  close(ngal, g["N_gal"])
  close(orc.completeness_P(g["z_range"], g["z_grids"])[:, None, :], g["P_compl"])
