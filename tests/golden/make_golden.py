"""Generate the golden fixtures `tests/golden/*.npz` by executing the UNMODIFIED reference
source (/root/reference/CHIMERA) under `jax_numpy_shim` (NumPy standing in for jax.numpy).

Run once in the build container (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Every fixture stores the inputs next to the reference outputs, so the tests need neither the
reference nor this script at run time.  Sizes are kept small (a few hundred kB in total).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import jax_numpy_shim as shim  # noqa: E402
from chimera_b200 import synth  # noqa: E402

CH = shim.import_reference()
from CHIMERA.data import theta_pe_det, theta_inj_det  # noqa: E402
from CHIMERA.utils import math as rmath  # noqa: E402
from CHIMERA.catalog import catalog as rcatalog  # noqa: E402
from CHIMERA.catalog.completeness import dVdz_completeness  # noqa: E402
from CHIMERA import data as rdata  # noqa: E402


def J(x):
  return np.asarray(x).view(shim.JArr)


def save(name, **arrs):
  path = os.path.join(HERE, name)
  np.savez_compressed(path, **arrs)
  print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} kB")


COSMO_CASES = [
  ("flrw", dict(H0=70., Om0=0.25)),
  ("flrw", dict(H0=55., Om0=0.4, z_max=5.)),
  ("flrw", dict(H0=67.7, Om0=0.31, w0=-0.9, wa=0.2, Or0=8e-5)),
  ("flrw", dict(H0=80., Om0=0.3, Ok0=0.05)),
  ("flrw", dict(H0=80., Om0=0.3, Ok0=-0.05)),
  ("mg_flrw", dict(H0=70., Om0=0.25, Xi0=1.8, n=1.9)),
  ("mg_flrw", dict(H0=64., Om0=0.3, Xi0=0.6, n=2.5, z_max=5.)),
]
MASS_CASES = [
  ("tpl", {}),
  ("tpl", dict(alpha=1.8, beta=0.3, m_low=4., m_high=60.)),
  ("bpl", {}),
  ("bpl", dict(alpha_1=2.1, alpha_2=4.4, beta=0.5, delta_m=3., break_fraction=0.3, m_low=6., m_high=70.)),
  ("plp", {}),
  ("plp", dict(lambda_peak=0.1, alpha=2.6, beta=-0.4, delta_m=6., mu_g=30., sigma_g=5., m_low=4.2, m_high=95.)),
]
RATE_CASES = [
  ("power_law", {}), ("power_law", dict(gamma=-0.5)),
  ("madau_dickinson", {}), ("madau_dickinson", dict(gamma=1.9, kappa=4.2, zp=1.4)),
  ("trunc_madau_dickinson", dict(zmax=0.9)), ("trunc_power_law", dict(gamma=2.3, zmax=1.1)),
]


def gen_models():
  out = {}
  z = np.concatenate([[0., 1e-12, 1e-6], np.geomspace(1e-4, 9.9, 40), [4.99, 5.0, 7.5, 12.]])
  out["z"] = z
  for i, (model, kw) in enumerate(COSMO_CASES):
    c = getattr(CH.cosmo, model)(**kw)
    dL = CH.cosmo.dL_at_z(c, J(z))
    out[f"cosmo{i}_dL"] = dL
    out[f"cosmo{i}_E"] = CH.cosmo.E_at_z(c, J(z))
    out[f"cosmo{i}_ddL"] = CH.cosmo.ddLdz_at_z(c, J(z))
    out[f"cosmo{i}_dV"] = CH.cosmo.dVcdz_at_z(c, J(z))
    out[f"cosmo{i}_Vc"] = CH.cosmo.Vc_at_z(c, J(z))
    dq = np.concatenate([[0., 1e-9], np.geomspace(1e-3, 60., 50), [1e4]])
    out[f"cosmo{i}_dq"] = dq
    zq = CH.cosmo.z_from_dGW(c, J(dq))
    out[f"cosmo{i}_zq"] = zq
    out[f"cosmo{i}_ddL_dist"] = CH.cosmo.ddLdz_at_z(c, J(zq), J(dq))
    out[f"cosmo{i}_dV_dist"] = CH.cosmo.dVcdz_at_z(c, J(zq), J(dq))
    out[f"cosmo{i}_tab"] = np.stack([c.z_grid_interp, c.integral_invE_interp])
  rng = np.random.default_rng(42)
  m1 = np.concatenate([rng.uniform(2., 110., 300), [5.1, 87., 5.1 + 4.8, 4.0, 100., 34.]])
  m2 = np.concatenate([rng.uniform(2., 110., 300) * rng.random(300), [5.1, 30., 5.1, 3.0, 50., 34.]])
  out["m1"], out["m2"] = m1, m2
  for i, (model, kw) in enumerate(MASS_CASES):
    m = getattr(CH.mass, model)(**kw)
    out[f"mass{i}_p"] = CH.mass.p_m1m2(m, J(m1), J(m2))
    out[f"mass{i}_norm"] = np.float64(m.norm_p_m1)
    out[f"mass{i}_cdf"] = np.asarray(m.cdf_m2_conditioned)
    out[f"mass{i}_p1"] = CH.mass.primary_mass_pdf_notnorm(m, J(m1))
  zr = np.linspace(0., 3., 61)
  out["zr"] = zr
  for i, (model, kw) in enumerate(RATE_CASES):
    r = getattr(CH.rate, model)(**kw)
    out[f"rate{i}"] = CH.rate.merger_rate(r, J(zr))
  save("golden_models.npz", **out)


def gen_math():
  rng = np.random.default_rng(7)
  out = {}
  x = rng.normal(0.6, 0.08, 700)
  w = rng.random(700) ** 3
  grid = np.linspace(0.2, 1.0, 90)
  out.update(x=x, w=w, grid=grid)
  c, s = rmath.binning1d(J(x), J(w), 50)
  out["bin_centers"], out["bin_sums"] = c, s
  for kern in ("epan", "gauss"):
    for j, bw in enumerate((None, "silverman", 0.37)):
      out[f"kde_{kern}_{j}"] = rmath.kde1d(J(x), J(grid), J(w), kernel=kern, bw_method=bw)
  out["kde_binned_epan"] = rmath.kde1d(c, J(grid), s, kernel="epan", bw_method=None)
  data = np.stack([x, rng.normal(1.0, 0.3, 700) + 2 * (x - 0.6), rng.normal(-0.2, 0.1, 700)])
  pts = np.stack([rng.normal(0.6, 0.08, 60), rng.normal(1.0, 0.3, 60), rng.normal(-0.2, 0.1, 60)])
  out["data3"], out["pts3"] = data, pts
  out["gkde3_numba"] = rmath.numba_gkde_nd(data, pts, weights=w, bw_method=None)
  out["gkde3_numba_silv"] = rmath.numba_gkde_nd(data, pts, weights=w, bw_method="silverman")
  out["gkde2_jax"] = rmath.jax_gkde_nd(J(data[:2]), J(pts[:2]))
  save("golden_math.npz", **out)


def build_inputs(nev, ns, nz, ninj, sky, seed, npix_target=6):
  ev = synth.make_events(nev, ns, seed=seed, sky=sky)
  z_grids = synth.make_z_grids(ev["dL"], z_int_res=nz, H0_prior=(40., 120.))
  inj, N_inj = synth.make_injections(ninj, seed=seed + 1)
  if sky:
    ev = synth.pixelize(ev, nside_list=(64, 32, 16, 8), mean_npixels_event=npix_target, sky_conf=0.9)
  return ev, z_grids, inj, N_inj


def ref_theta(ev, pixelated):
  kw = dict(m1det=J(ev["m1det"]), m2det=J(ev["m2det"]), dL=J(ev["dL"]), pe_prior=J(ev["pe_prior"]))
  if pixelated:
    kw.update(ra=J(ev["ra"]), dec=J(ev["dec"]), opt_nsides=J(ev["opt_nsides"]),
              pixels_opt_nsides=J(ev["pixels_opt_nsides"]), ra_pix=J(ev["ra_pix"]),
              dec_pix=J(ev["dec_pix"]), gw_loc2d_pdf=J(ev["gw_loc2d_pdf"]),
              pixels_pe_opt_nside=J(ev["pixels_pe_opt_nside"]))
  return theta_pe_det(**kw)


LIKE_CASES = {
  # name: (pixel kind, kernel, binning, cosmo model, hyper-points)
  "1d_epan_binned": (None, "epan", True, "flrw",
                     [dict(H0=h) for h in (50., 62., 70., 81., 95.)]),
  "1d_gauss_unbinned": (None, "gauss", False, "flrw",
                        [dict(H0=60., Om0=0.2), dict(H0=70., Om0=0.25), dict(H0=78., Om0=0.4)]),
  "1d_gauss_binned_mg": (None, "gauss", True, "mg_flrw",
                         [dict(H0=70., Xi0=1.0, n=0.), dict(H0=66., Xi0=1.6, n=1.9, alpha=3.0, mu_g=32., gamma=2.2),
                          dict(H0=74., Xi0=0.7, n=2.4, beta=0.8, delta_m=5.5, m_low=4.8, m_high=90., sigma_g=4.2,
                               lambda_peak=0.06, kappa=3.5, zp=1.8)]),
  "1d_epan_unbinned": (None, "epan", False, "flrw", [dict(H0=65.), dict(H0=75., Om0=0.3)]),
  "approx_gauss_unbinned": ("approximate", "gauss", False, "flrw",
                            [dict(H0=58., Om0=0.22), dict(H0=70., Om0=0.25), dict(H0=84., Om0=0.33)]),
  "approx_epan_binned": ("approximate", "epan", True, "flrw", [dict(H0=64.), dict(H0=70.), dict(H0=77.)]),
  "marg_binned": ("marginalized", "epan", True, "flrw", [dict(H0=61.), dict(H0=70.), dict(H0=88.)]),
  "marg_unbinned": ("marginalized", "epan", False, "flrw", [dict(H0=66.), dict(H0=73., Om0=0.28)]),
  "full_gauss": ("full", "gauss", False, "flrw", [dict(H0=63.), dict(H0=70.), dict(H0=79., Om0=0.3)]),
}


def gen_like():
  ev1, zg1, inj, N_inj = build_inputs(nev=10, ns=400, nz=60, ninj=3000, sky=False, seed=11)
  evp, zgp, _, _ = build_inputs(nev=6, ns=500, nz=50, ninj=10, sky=True, seed=21)
  gal = synth.make_galaxies(60_000, seed=31)
  # pixelated catalogue through the reference's own constructor (file loader patched only)
  rcatalog.load_galaxy_catalog = lambda fname, backend="numpy": dict(ra=gal["ra"], dec=gal["dec"], z=gal["z"])
  fid = CH.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  thp = ref_theta(evp, True)
  gcat = rcatalog.pixelated_catalog(completeness=dVdz_completeness([0.073, 1.3]), cosmo=fid,
                                    z_grids=J(zgp), fname_data_gal="synthetic", data_gw_pixelated=thp,
                                    z_err=0.001)
  common = dict(inj_m1det=inj["m1det"], inj_m2det=inj["m2det"], inj_dL=inj["dL"], inj_p_draw=inj["p_draw"],
                N_inj=np.float64(N_inj), N_eff=np.float64(5.))
  save("golden_inputs_1d.npz", z_grids=zg1, **{k: ev1[k] for k in ("m1det", "m2det", "dL", "pe_prior")}, **common)
  save("golden_inputs_pix.npz", z_grids=zgp, p_cat=np.asarray(gcat.p_cat), P_compl=np.asarray(gcat.P_compl),
       N_gal=np.asarray(gcat.N_gal), z_range=np.array([0.073, 1.3]),
       gal_ra=gal["ra"], gal_dec=gal["dec"], gal_z=gal["z"],
       **{k: evp[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides", "pixels_opt_nsides",
                              "ra_pix", "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside", "neff_pixels")},
       **common)
  sel = CH.selection_function(theta_inj_det(m1det=J(inj["m1det"]), m2det=J(inj["m2det"]), dL=J(inj["dL"]),
                                            p_draw=J(inj["p_draw"])), N_inj, 5.)
  out = {}
  for name, (kind, kernel, binning, cmodel, hypers) in LIKE_CASES.items():
    pix = kind is not None
    ev, zg = (evp, zgp) if pix else (ev1, zg1)
    th = thp if pix else ref_theta(ev, False)
    cosmo = getattr(CH.cosmo, cmodel)(H0=70., Om0=0.25, z_max=5.)
    pop = CH.population(cosmo, CH.mass.plp(), CH.rate.madau_dickinson(), gal_cat=gcat if pix else None)
    like = CH.hyperlikelihood(th, J(zg), pop, sel, kind_p_gw3d=kind, kernel=kernel, binning=binning,
                              num_bins=40, pe_neff=2.0, cut_grid=2.0)
    for h, hl in enumerate(hypers):
      lle, lnum, lnexp, lh = like.compute_all(**hl)
      out[f"{name}_h{h}_lle"] = np.asarray(lle)
      out[f"{name}_h{h}_tot"] = np.array([lnum, lnexp, lh], dtype=np.float64)
      popl = pop.update(**hl)
      pgw = like.p_gw3d(popl) if pix else like.p_gw1d(popl)
      out[f"{name}_h{h}_pgw"] = np.asarray(pgw)
      if h == 0:
        out[f"{name}_h0_xi"] = np.float64(sel.N_exp(popl))
    out[f"{name}_nh"] = np.int64(len(hypers))
    print(name, [float(out[f"{name}_h{h}_tot"][2]) for h in range(len(hypers))])
  # scale_free=False and N_eff gate variants on the 1-D case
  cosmo = CH.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  pop = CH.population(cosmo, CH.mass.plp(), CH.rate.madau_dickinson(), R0=17., Tobs=2.5, scale_free=False)
  like = CH.hyperlikelihood(ref_theta(ev1, False), J(zg1), pop, sel, kernel="epan", binning=True, num_bins=40)
  out["notscalefree_tot"] = np.array([float(v) for v in like.compute_all(H0=68., R0=21.)[1:]])
  sel_strict = CH.selection_function(sel.theta_inj_det, N_inj, 1e9)
  pop = CH.population(cosmo, CH.mass.plp(), CH.rate.madau_dickinson())
  like = CH.hyperlikelihood(ref_theta(ev1, False), J(zg1), pop, sel_strict, kernel="epan", binning=True, num_bins=40)
  out["neffgate_tot"] = np.array([float(v) for v in like.compute_all(H0=68.)[1:]])
  save("golden_like.npz", **out)


def gen_like2():
  """Round 2: the model / option matrix of tests/cases.py LIKE_CASES2 on the SAME inputs as gen_like (same seeds), written to
  golden_like2.npz; the round-1 fixtures are left byte-identical."""
  from cases import LIKE_CASES2
  ev1, zg1, inj, N_inj = build_inputs(nev=10, ns=400, nz=60, ninj=3000, sky=False, seed=11)
  evp, zgp, _, _ = build_inputs(nev=6, ns=500, nz=50, ninj=10, sky=True, seed=21)
  gal = synth.make_galaxies(60_000, seed=31)
  rcatalog.load_galaxy_catalog = lambda fname, backend="numpy": dict(ra=gal["ra"], dec=gal["dec"], z=gal["z"])
  fid = CH.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  thp = ref_theta(evp, True)
  gcat = rcatalog.pixelated_catalog(completeness=dVdz_completeness([0.073, 1.3]), cosmo=fid,
                                    z_grids=J(zgp), fname_data_gal="synthetic", data_gw_pixelated=thp, z_err=0.001)
  sel = CH.selection_function(theta_inj_det(m1det=J(inj["m1det"]), m2det=J(inj["m2det"]), dL=J(inj["dL"]),
                                            p_draw=J(inj["p_draw"])), N_inj, 5.)
  out = {}
  for name, c in LIKE_CASES2.items():
    pix = c["kind"] is not None
    ev, zg = (evp, zgp) if pix else (ev1, zg1)
    th = thp if pix else ref_theta(ev, False)
    cosmo = getattr(CH.cosmo, c["cosmo"][0])(H0=70., Om0=0.25, z_max=5., **c["cosmo"][1])
    mass = getattr(CH.mass, c["mass"][0])(**c["mass"][1])
    rate = getattr(CH.rate, c["rate"][0])(**c["rate"][1])
    pop = CH.population(cosmo, mass, rate, gal_cat=gcat if pix else None)
    like = CH.hyperlikelihood(th, J(zg), pop, sel, kind_p_gw3d=c["kind"], kernel=c["kernel"], bw_method=c["bw"],
                              binning=c["binning"], num_bins=40, pe_neff=2.0, cut_grid=2.0)
    for h, hl in enumerate(c["hypers"]):
      lle, lnum, lnexp, lh = like.compute_all(**hl)
      out[f"{name}_h{h}_lle"] = np.asarray(lle)
      out[f"{name}_h{h}_tot"] = np.array([lnum, lnexp, lh], dtype=np.float64)
    out[f"{name}_nh"] = np.int64(len(c["hypers"]))
    print(name, [float(out[f"{name}_h{h}_tot"][2]) for h in range(len(c["hypers"]))])
  # selection function alone, bpl + mg_flrw + truncated rate (the fp32 selection kernel had never been compared with bpl)
  pop = CH.population(CH.cosmo.mg_flrw(H0=70., Om0=0.25, z_max=5.), CH.mass.bpl(), CH.rate.trunc_madau_dickinson(zmax=2.0))
  hs = [dict(H0=60., Xi0=0.7, n=1.9), dict(H0=70., Xi0=1.0, n=0.), dict(H0=82., Xi0=1.8, n=2.3, alpha_1=2.0, break_fraction=0.3)]
  out["sel_bpl_mg_nexp"] = np.array([float(sel.N_exp(pop.update(**hl))) for hl in hs])
  save("golden_like2.npz", **out)


def gen_setup():
  ev, zg, _, _ = build_inputs(nev=5, ns=300, nz=40, ninj=10, sky=True, seed=51)
  out = dict(dL=ev["dL"], ra=ev["ra"], dec=ev["dec"])
  th = theta_pe_det(dL=J(ev["dL"]))
  fid = CH.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  out["zgrid_default"] = CH.compute_z_grids(fid, th, cosmo_prior={"H0": [40., 120.]}, z_int_res=40)
  out["zgrid_sigma"] = CH.compute_z_grids(fid, th, cosmo_prior={"H0": [40., 120.], "Om0": [0.2, 0.4]},
                                          z_int_res=40, z_conf_range=3.)
  out["zgrid_pct"] = CH.compute_z_grids(fid, th, z_int_res=40, z_conf_range=[1., 99.])
  mg = CH.cosmo.mg_flrw(H0=70., Om0=0.25, z_max=5.)
  out["zgrid_mg"] = CH.compute_z_grids(mg, th, cosmo_prior={"H0": [50., 90.], "Xi0": [0.5, 2.], "n": [1., 3.]},
                                       z_int_res=40)
  # the reference's own pixelisation on the same samples (healpy := chimera_b200.healpix)
  thp = theta_pe_det(m1det=J(ev["m1det"]), m2det=J(ev["m2det"]), dL=J(ev["dL"]), ra=J(ev["ra"]), dec=J(ev["dec"]))
  pix = rdata.pixelize_gw_catalog(thp, nside_list=[64, 32, 16, 8], mean_npixels_event=6, sky_conf=0.9)
  out["pix_opt_nsides"] = np.asarray(pix.opt_nsides)
  out["pix_pixels"] = np.asarray(pix.pixels_opt_nsides)
  out["pix_ra"] = np.asarray(pix.ra_pix)
  out["pix_dec"] = np.asarray(pix.dec_pix)
  out["pix_pdf"] = np.asarray(pix.gw_loc2d_pdf)
  out["pix_pe"] = np.asarray(pix.pixels_pe_opt_nside)
  save("golden_setup.npz", **out)


if __name__ == "__main__":
  if "--like2" in sys.argv:      # round 2: only the new fixture (the round-1 files stay byte-identical)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    gen_like2()
    sys.exit(0)
  gen_models()
  gen_math()
  gen_like()
  gen_setup()
