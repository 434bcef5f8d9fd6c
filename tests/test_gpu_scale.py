"""GPU tests at the sizes BASELINE.json names (C3 / C4 / C5 shapes), through size-independent properties:
fp32 mode against the fp64 mode of the same library, invariance under permutation of the posterior samples
(the windowed KDE sorts them), exact shift under a rescaling of pe_prior, independence of the hyper-point
batching and of the event sharding, unit integral of every catalogue row, and the oracle on a subset of units."""
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


def _err(x, r):
  fin = np.isfinite(r) & (np.abs(r) < 1e300)
  assert np.array_equal(fin, np.isfinite(x) & (np.abs(x) < 1e300))
  return float(np.max(np.abs(x[fin] - r[fin]) / np.maximum(np.abs(r[fin]), 1.0)))


@pytest.fixture(scope="module")
def c3(cb):
  """C3 at full size: 1000 events x 5000 samples, pixelated catalogue ('approximate'), Gaussian KDE unbinned."""
  import bench
  w = bench.build_workload("C3", ninj=200_000)      # GPU pixelisation + precompute_p_cat from 1.6e6 galaxies (spiky rows)
  like = bench.build_likelihood(w, "fp32")
  return w, like


def test_c3_fp32_vs_fp64_and_oracle(cb, c3):
  import bench
  from oracle import chimera_oracle as orc
  w, like = c3
  idx = np.linspace(0, 255, 6).astype(int)
  hy = {k: v[idx] for k, v in w["hyper"].items()}
  l32 = like.compute_all(**hy)
  like64 = bench.build_likelihood(w, "fp64")
  l64 = like64.compute_all(**hy)
  assert _err(l32[0], l64[0]) < 1e-5            # per-event log-likelihoods, fp32 mode vs fp64 mode (budget 1e-3)
  np.testing.assert_allclose(l32[3], l64[3], rtol=1e-6)    # total log hyper-likelihood
  # oracle on the first 12 events, two hyper-points (fp64 mode 1e-9, fp32 mode 1e-5)
  ev = {k: np.asarray(w["ev"][k])[:12] for k in ("m1det", "m2det", "dL", "pe_prior", "pixels_opt_nsides", "gw_loc2d_pdf")}
  cat = dict(p_cat=w["p_cat"][:12], P_compl=w["P_compl"][:12][:, None, :], z_range=w["z_range"])
  pop0 = orc.make_pop(orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                      orc.make_rate("madau_dickinson"), catalog=cat)
  opts = orc.make_opts("approximate", "gauss", None, 2.0, False, 200, 2.0)
  for j in (0, 5):
    pop = orc.pop_update(pop0, H0=float(hy["H0"][j]), Om0=float(hy["Om0"][j]))
    with np.errstate(all="ignore"):
      ref = np.nan_to_num(np.log(orc.numlike_evs(pop, ev, w["zg"][:12], opts, w["ev"]["neff_pixels"][:12])), nan=-np.inf)
    assert _err(l64[0][j, :12], ref) < 1e-9
    assert _err(l32[0][j, :12], ref) < 1e-5


def test_c3_all_units_fp32_vs_fp64(cb, c3):
  """Every one of the 256 x 1000 (hyper-point, event) units of C3: fp32 mode against the fp64 mode of the same library
  (which the test above holds to the oracle at 1e-9).  Budget of the fp32 mode: 1e-3; asserted: 1e-5."""
  import bench
  w, like = c3
  l32 = like.compute_all(**w["hyper"])
  l64 = bench.build_likelihood(w, "fp64").compute_all(**w["hyper"])
  assert l32[0].shape == (256, 1000)
  assert _err(l32[0], l64[0]) < 1e-5
  np.testing.assert_allclose(l32[3], l64[3], rtol=1e-6)


def test_c3_invariances(cb, c3):
  import bench
  w, like = c3
  idx = np.array([3, 100, 250])
  hy = {k: v[idx] for k, v in w["hyper"].items()}
  base = like.compute_all(**hy)[0]
  # (1) batching: the same hyper-points inside a larger batch, in another order
  big = {k: np.concatenate([v[::17], v[idx][::-1]]) for k, v in w["hyper"].items()}
  out = like.compute_all(**big)[0]
  np.testing.assert_array_equal(out[-3:][::-1], base)
  # (2) permutation of the samples within every event + rescaled pe_prior: log L shifts by -log(c) exactly
  rng = np.random.default_rng(0)
  ev = dict(w["ev"])
  perm = np.argsort(rng.random(ev["dL"].shape), axis=1)
  for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "pixels_pe_opt_nside"):
    ev[k] = np.take_along_axis(np.asarray(ev[k]), perm, axis=1)
  c = 7.25
  ev["pe_prior"] = ev["pe_prior"] * c
  w2 = dict(w, ev=ev)
  like2 = bench.build_likelihood(w2, "fp32")
  out2 = like2.compute_all(**hy)[0]
  assert _err(out2 + np.log(c), base) < 2e-6
  # (3) sharding: two handles with half of the events each reproduce the per-event values bit for bit
  half = {k: (np.asarray(v)[:500] if np.ndim(v) and np.shape(v)[0] == 1000 else v) for k, v in w["ev"].items()}
  w3 = dict(w, ev=half, zg=w["zg"][:500], p_cat=w["p_cat"][:500], P_compl=w["P_compl"][:500])
  out3 = bench.build_likelihood(w3, "fp32").compute_all(**hy)[0]
  np.testing.assert_array_equal(out3, base[:, :500])


def test_c4_catalogue_rows_integrate_to_one(cb):
  """C4 shape: nside = 64 pixels, 10^7 galaxies, 500 events.  Every (event, pixel) row of p_cat that holds at
  least one galaxy is a weighted mean of Gaussians each normalised by the same trapezoid rule on the event grid
  (catalog.py:209-221), so its trapezoid integral is 1; rows without galaxies are 0; padded slots -100."""
  from chimera_b200 import synth
  nev = 500
  ev = synth.make_events(nev, 512, seed=41, sky=True)
  zg = synth.make_z_grids(ev["dL"], z_int_res=300, H0_prior=(40., 120.))
  ev = synth.pixelize(ev, nside_list=(64,), mean_npixels_event=15)
  gal = synth.make_galaxies(10_000_000, seed=42)
  th = cb.theta_pe_det(**{k: ev[k] for k in ("dL", "ra", "dec", "opt_nsides", "pixels_opt_nsides", "ra_pix", "dec_pix",
                                             "gw_loc2d_pdf", "pixels_pe_opt_nside")})
  fid = cb.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  gcat = cb.pixelated_catalog(cb.dVdz_completeness([0.073, 1.3]), cosmo=fid, z_grids=zg, data_gw_pixelated=th,
                              data_gal=dict(ra=gal["ra"], dec=gal["dec"], z=gal["z"]), z_err=0.01)
  p = gcat.p_cat
  assert p.shape == (nev, ev["pixels_opt_nsides"].shape[1], 300)
  pad = ev["pixels_opt_nsides"] == -100
  assert np.all(p[pad] == -100.) and np.all(p[~pad] >= 0.)
  integ = np.trapezoid(p, zg[:, None, :], axis=2)
  rows = (~pad) & (np.max(p, axis=2) > 0)
  assert rows.sum() > 1000 and gcat.N_gal.sum() > 1e5
  np.testing.assert_allclose(integ[rows], 1.0, rtol=1e-12)
  # galaxy counts: bucketing by pixel is exact (HEALPix ids on the GPU == host restatement)
  from chimera_b200 import healpix as hp_host
  gp = hp_host.find_pix_RAdec(gal["ra"][:2_000_000], gal["dec"][:2_000_000], 64)
  np.testing.assert_array_equal(cb.sky.find_pix_RAdec(gal["ra"][:2_000_000], gal["dec"][:2_000_000], 64), gp)


def test_c4_full_3d_kde_vs_oracle(cb):
  """C4 likelihood shape at reduced event count: full 3-D KDE, nside = 64, catalogue built on the GPU with an
  incompleteness correction (P_compl step + background term), fp64 and fp32 modes against the oracle."""
  from oracle import chimera_oracle as orc
  from chimera_b200 import synth, healpix as hp_host
  ev = synth.make_events(10, 2000, seed=51, sky=True)
  zg = synth.make_z_grids(ev["dL"], z_int_res=120, H0_prior=(40., 120.))
  ev = synth.pixelize(ev, nside_list=(64,), mean_npixels_event=12)
  gal = synth.make_galaxies(400_000, seed=52)
  th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides",
                                             "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf",
                                             "pixels_pe_opt_nside")})
  fid = cb.cosmo.flrw(H0=70., Om0=0.25, z_max=5.)
  z_range = np.array([0.05, 0.9])
  gcat = cb.pixelated_catalog(cb.dVdz_completeness(z_range), cosmo=fid, z_grids=zg, data_gw_pixelated=th,
                              data_gal=dict(ra=gal["ra"], dec=gal["dec"], z=gal["z"]), z_err=0.005)
  inj, N_inj = synth.make_injections(20000, seed=53)
  sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
  pop = cb.population(fid, cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=gcat)
  pop0 = orc.make_pop(orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                      orc.make_rate("madau_dickinson"), catalog=dict(p_cat=gcat.p_cat, P_compl=gcat.P_compl, z_range=z_range))
  opts = orc.make_opts("full", "gauss", None, 2.0, False, 200, 2.0)
  H0 = np.array([62., 70., 81.])
  for fp_mode, tol in (("fp64", 1e-9), ("fp32", 1e-4)):
    like = cb.hyperlikelihood(th, zg, pop, sel, kind_p_gw3d="full", kernel="gauss", fp_mode=fp_mode)
    lle = like.compute_all(H0=H0)[0]
    for j, h0 in enumerate(H0):
      ref = orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., ev["neff_pixels"], H0=float(h0))[0]
      assert _err(lle[j], ref) < tol, (fp_mode, h0)


def test_full_3d_windows_match_all_pairs(cb):
  """'full' 3-D KDE, fp32 mode: the sample-block windows (blocks of dL-sorted samples far from a tile of evaluation points
  in the z-only whitened coordinate are skipped when below 2^-30 of the largest term at every point of the tile) against
  the same kernel visiting every pair (`kde_win=0`): per-event log-likelihoods, and the p_gw arrays INCLUDING their far
  tails (relative agreement wherever the density is above 1e-30 of its maximum)."""
  from chimera_b200 import synth
  ev = synth.make_events(12, 4096, seed=71, sky=True)
  zg = synth.make_z_grids(ev["dL"], z_int_res=200, H0_prior=(30., 140.))      # a wide grid: many points far from the samples
  ev = synth.pixelize(ev, nside_list=(64, 32), mean_npixels_event=14)
  p_cat, P_compl = synth.smooth_p_cat(ev, zg, seed=72)
  th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides",
                                             "pixels_opt_nsides", "ra_pix", "dec_pix", "gw_loc2d_pdf",
                                             "pixels_pe_opt_nside")})
  gcat = cb.pixelated_catalog(cb.dVdz_completeness(np.array([0.073, 1.3])), p_cat=p_cat, P_compl=P_compl)
  pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=gcat)
  H0 = np.array([45., 70., 110.])
  out = {}
  for name, opt in (("win", {}), ("all", {"kde_win": 0})):
    like = cb.hyperlikelihood(th, zg, pop, None, kind_p_gw3d="full", kernel="gauss", fp_mode="fp32", options=opt)
    out[name] = (like.compute_all(H0=H0)[0], like.p_gw3dfull(pop.update(H0=70.)))
  assert _err(out["win"][0], out["all"][0]) < 2e-6
  pw, pa = out["win"][1], out["all"][1]
  big = pa > 1e-30 * np.max(pa)
  assert big.sum() > 1000
  np.testing.assert_allclose(pw[big], pa[big], rtol=2e-5)
  assert np.all(pw[~big] <= 1e-29 * np.max(pa))


def test_c5_modified_gravity_walker_batch(cb):
  """C5 shape at reduced event count: mg_flrw (Xi0, n) + mass + rate hyper-parameters, a 4096-point walker matrix
  through the sampler front end in ONE batched call; a strided subset against the oracle, the rest against
  chunked evaluation (batch independence)."""
  from oracle import chimera_oracle as orc
  from chimera_b200 import synth, sampling
  ev = synth.make_events(24, 1024, seed=61, sky=False)
  zg = synth.make_z_grids(ev["dL"], z_int_res=200, H0_prior=(30., 140.))
  inj, N_inj = synth.make_injections(50_000, seed=62)
  th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")})
  sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
  pop = cb.population(cb.cosmo.mg_flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson())
  like = cb.hyperlikelihood(th, zg, pop, sel, kernel="gauss", binning=False, fp_mode="fp32")
  keys = ["H0", "Xi0", "n", "alpha", "beta", "mu_g", "sigma_g", "lambda_peak", "gamma", "kappa"]
  lo = np.array([55., 0.6, 0.5, 2.5, 0.5, 28., 2.5, 0.01, 1.5, 2.0])
  hi = np.array([85., 1.9, 3.0, 4.2, 2.0, 38., 6.0, 0.10, 3.5, 4.5])
  rng = np.random.default_rng(7)
  walkers = lo + (hi - lo) * rng.random((4096, len(keys)))
  walkers[::500, 0] = 500.                               # outside the prior: never evaluated, -inf
  lp = sampling.log_prob_fn(like, keys, sampling.uniform_log_prior(np.stack([lo, hi], axis=1)))
  out = lp(walkers)
  # walkers may legitimately hit the N_eff gate (+inf) or a zero-likelihood event (-inf), exactly like the reference
  assert out.shape == (4096,) and np.all(np.isneginf(out[::500])) and np.mean(np.isfinite(out)) > 0.9
  # batch independence: the same walkers in chunks of 1000 (per-event values are bit-identical, test_c3_invariances;
  # the injection sums are tiled by batch size, so the totals agree to fp64 rounding)
  chunks = np.concatenate([lp(walkers[i:i + 1000]) for i in range(0, 4096, 1000)])
  fin = np.isfinite(out)
  np.testing.assert_array_equal(chunks[~fin], out[~fin])
  np.testing.assert_allclose(chunks[fin], out[fin], rtol=1e-12)
  # oracle on a few walkers
  pop0 = orc.make_pop(orc.make_cosmo("mg_flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  opts = orc.make_opts(None, "gauss", None, 2.0, False, 200, 2.0)
  for i in (1, 1234, 4095):
    ref = orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., None, **dict(zip(keys, walkers[i])))[3]
    assert abs(out[i] - ref) <= 1e-4 * max(abs(ref), 1.0), (i, out[i], ref)
