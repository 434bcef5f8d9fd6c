"""CPU tests of the host-side mirror of the reference interface (no GPU compute)."""
import numpy as np
import pytest


def test_struct_update_semantics():
  """update ignores unknown keys and returns self when nothing matches (cosmo.py:33-40 etc.)."""
  import chimera_b200 as cb
  for s in (cb.cosmo.flrw(), cb.cosmo.mg_flrw(), cb.mass.tpl(), cb.mass.bpl(), cb.mass.plp(),
            cb.rate.power_law(), cb.rate.madau_dickinson(), cb.rate.trunc_madau_dickinson(), cb.rate.trunc_power_law()):
    assert s.update(not_a_key=1.0) is s
    k = s.keys[-1]
    t = s.update(**{k: 0.123, "junk": 1})
    assert t is not s and getattr(t, k) == 0.123 and type(t) is type(s)
    assert s.as_dict == {kk: getattr(s, kk) for kk in s.keys}
  c = cb.cosmo.flrw(H0=60.)
  assert c.dH == 299792.458e-3 / 60. and c.Ode0 == 1.0 - 0.25
  assert cb.cosmo.flrw.default["z_grid_res"] == 1500 and cb.cosmo.mg_flrw().name == "mg_flrw"
  assert cb.mass.plp().name == "power_law_plus_peak" and cb.mass.tpl.default["alpha"] == 2.5


def test_population_routing_and_rows():
  import chimera_b200 as cb
  from chimera_b200 import _lib
  pop = cb.population(cb.cosmo.mg_flrw(), cb.mass.plp(), cb.rate.madau_dickinson(), R0=3.)
  rows, batched = pop.hyper_rows()
  assert rows.shape == (1, _lib.CHB_NPAR) and not batched
  p2 = pop.update(H0=[60., 70., 80.], Xi0=1.3, alpha=[3., 3.1, 3.2], zp=1.5, R0=[1., 2., 3.], nonsense=5)
  rows, batched = p2.hyper_rows()
  assert batched and rows.shape == (3, _lib.CHB_NPAR)
  S = _lib.SLOT
  np.testing.assert_array_equal(rows[:, S["H0"]], [60., 70., 80.])
  np.testing.assert_array_equal(rows[:, S["Xi0"]], 1.3)
  np.testing.assert_array_equal(rows[:, S["alpha"]], [3., 3.1, 3.2])
  np.testing.assert_array_equal(rows[:, S["zp"]], 1.5)
  np.testing.assert_array_equal(rows[:, S["R0"]], [1., 2., 3.])
  np.testing.assert_array_equal(rows[:, 27], np.power(10., np.log10(87.)))
  assert pop.cosmo.H0 == 70. and p2.gal_cat is pop.gal_cat and p2.scale_free is True
  with pytest.raises(ValueError):
    pop.update(H0=[1., 2.], Om0=[0.1, 0.2, 0.3]).hyper_rows()
  bpl = cb.population(cb.cosmo.flrw(), cb.mass.bpl(alpha_1=1.7, alpha_2=5.0), cb.rate.trunc_power_law(zmax=0.9))
  rows, _ = bpl.hyper_rows()
  assert rows[0, S["alpha_1"]] == 1.7 and rows[0, S["alpha_2"]] == 5.0 and rows[0, S["zmax"]] == 0.9


def test_theta_structs():
  import chimera_b200 as cb
  t = cb.theta_pe_det(dL=np.ones((2, 3)))
  assert t.pe_prior.shape == (2, 3) and np.all(t.pe_prior == 1.) and t.pixels_opt_nsides is None
  u = t.update(m1det=np.zeros((2, 3)))
  assert u is not t and t.m1det is None and u.m1det.shape == (2, 3)
  with pytest.raises(AttributeError):
    t.update(bogus=1)
  with pytest.raises(TypeError):
    cb.theta_inj_det(foo=1)


def test_shard_bounds_reference_rule():
  """n // R per rank, the first n % R ranks get one more (CHIMERA/parallel.py:94-99)."""
  from chimera_b200.parallel import shard_bounds
  for n in (0, 1, 7, 300, 1000, 10007):
    for world in (1, 2, 3, 8):
      bounds = [shard_bounds(n, r, world) for r in range(world)]
      assert bounds[0][0] == 0 and bounds[-1][1] == n
      for (a, b), (c, d) in zip(bounds[:-1], bounds[1:]):
        assert b == c
      sizes = [b - a for a, b in bounds]
      chunk, rem = divmod(n, world)
      assert sizes == [chunk + 1] * rem + [chunk] * (world - rem)


def test_completeness_and_catalog_objects():
  import chimera_b200 as cb
  c = cb.dVdz_completeness([0.1, 1.0])
  zg = np.array([[0.05, 0.1, 0.5, 1.0, 1.2]])
  np.testing.assert_array_equal(c.P_compl(zg), [[0., 0., 1., 0., 0.]])
  with pytest.raises(ValueError):
    cb.dVdz_completeness([0.1, 1.0], kind="step_smooth")
  p_cat = np.full((2, 3, 4), -100.)
  p_cat[0, :2] = 1.0
  p_cat[1, :1] = 2.0
  g = cb.pixelated_catalog(c, p_cat=p_cat, P_compl=np.ones((2, 4)))
  assert g.max_npixels == 3 and list(g.neff_pixels) == [2, 1] and g.P_compl.shape == (2, 1, 4)
  with pytest.raises(ValueError):
    cb.empty_catalog(p_bkg=lambda c, z: z)


def test_likelihood_constructor_checks_without_gpu():
  """Option validation happens before any device work (likelihood.py:85, math.py:75)."""
  import chimera_b200 as cb
  th = cb.theta_pe_det(m1det=np.ones((2, 8)), m2det=np.ones((2, 8)), dL=np.ones((2, 8)),
                       pixels_opt_nsides=np.zeros((2, 3), dtype=np.int64))
  c = cb.dVdz_completeness([0.1, 1.0])
  g = cb.pixelated_catalog(c, p_cat=np.zeros((2, 3, 6)), P_compl=np.ones((2, 6)))
  pop = cb.population(cb.cosmo.flrw(), cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=g)
  with pytest.raises(AssertionError):
    cb.hyperlikelihood(th, np.ones((2, 6)), pop, None, kind_p_gw3d=None)
  with pytest.raises(ValueError):
    cb.hyperlikelihood(th, np.ones((2, 6)), pop, None, kind_p_gw3d="full", bw_method="nope")
  with pytest.raises(ValueError):
    cb.hyperlikelihood(th, np.ones((2, 6)), pop, None, kind_p_gw3d="full", kernel="box")


# ------------------------------------------------------------------------------------------ HEALPix
@pytest.mark.parametrize("nside", [1, 2, 8, 64, 512])
def test_healpix_roundtrip_and_ranges(nside):
  from chimera_b200 import healpix as hp
  npix = hp.nside2npix(nside)
  assert npix == 12 * nside * nside
  pix = np.arange(npix) if npix <= 50000 else np.random.default_rng(1).integers(0, npix, 50000)
  th, ph = hp.pix2ang(nside, pix)
  assert np.all((th > 0) & (th < np.pi)) and np.all((ph >= 0) & (ph < 2 * np.pi))
  np.testing.assert_array_equal(hp.ang2pix(nside, th, ph), pix)          # centres map back to their pixel
  ra, dec = hp.find_ra_dec(pix, nside)
  np.testing.assert_array_equal(hp.find_pix_RAdec(ra, dec, nside), pix)


def test_healpix_equal_area_and_ring_structure():
  from chimera_b200 import healpix as hp
  nside = 16
  rng = np.random.default_rng(2)
  n = 2_000_000
  th = np.arccos(rng.uniform(-1, 1, n))
  ph = rng.uniform(0, 2 * np.pi, n)
  pix = hp.ang2pix(nside, th, ph)
  assert pix.dtype == np.int64 and pix.min() >= 0 and pix.max() < hp.nside2npix(nside)
  counts = np.bincount(pix, minlength=hp.nside2npix(nside))
  mean = n / hp.nside2npix(nside)
  assert np.all(np.abs(counts - mean) < 6 * np.sqrt(mean))            # equal-area pixels
  # RING ordering: colatitude of pixel centres is non-decreasing with the index
  tc, _ = hp.pix2ang(nside, np.arange(hp.nside2npix(nside)))
  assert np.all(np.diff(tc) >= -1e-15)
  # known values for nside=1 (12 base pixels): north cap pixels at phi = pi/4 + k pi/2
  t1, p1 = hp.pix2ang(1, np.arange(12))
  np.testing.assert_allclose(p1[:4], np.pi / 4 + np.arange(4) * np.pi / 2)
  np.testing.assert_allclose(np.cos(t1[:4]), 2. / 3.)
  np.testing.assert_allclose(np.cos(t1[4:8]), 0., atol=1e-15)
  assert hp.ang2pix(1, 0.0, 0.0) == 0 and hp.ang2pix(1, np.pi, 0.0) == 8
  with pytest.raises(ValueError):
    hp.ang2pix(3, 0.1, 0.1)


def test_pixelize_matches_reference_procedure(golden_setup):
  """synth.pixelize restates data.py:262-392; the fixture holds the reference's own output on the
  same samples (with healpy := chimera_b200.healpix).  Pixel ids must be bit-exact."""
  from chimera_b200 import synth
  g = golden_setup
  ev = synth.pixelize(dict(ra=g["ra"], dec=g["dec"]), nside_list=(64, 32, 16, 8), mean_npixels_event=6, sky_conf=0.9)
  np.testing.assert_array_equal(ev["opt_nsides"], g["pix_opt_nsides"])
  np.testing.assert_array_equal(ev["pixels_opt_nsides"], g["pix_pixels"])
  np.testing.assert_array_equal(ev["pixels_pe_opt_nside"], g["pix_pe"])
  np.testing.assert_allclose(ev["ra_pix"], g["pix_ra"], rtol=1e-15)
  np.testing.assert_allclose(ev["dec_pix"], g["pix_dec"], rtol=1e-15, atol=1e-15)
  np.testing.assert_allclose(ev["gw_loc2d_pdf"], g["pix_pdf"], rtol=1e-9)


def test_synthetic_workload_shapes():
  from chimera_b200 import synth
  ev = synth.make_events(5, 64, seed=1, sky=True)
  assert ev["dL"].shape == (5, 64) and np.all(ev["m1det"] >= ev["m2det"]) and np.all(ev["pe_prior"] > 0)
  inj, N = synth.make_injections(1000, seed=2)
  assert all(v.shape == (1000,) for v in inj.values()) and N >= 1000 and np.all(inj["p_draw"] > 0)
  zg = synth.make_z_grids(ev["dL"], 30)
  assert zg.shape == (5, 30) and np.all(np.diff(zg, axis=1) > 0)


def test_sampler_front_end_host_logic():
  """generate_dict (emcee_utils.py:54-64) and the vectorised log-probability wrapper: walkers outside the
  prior are never sent to the likelihood, the rest go in ONE batched call."""
  from chimera_b200 import sampling
  pos = np.array([[70., 0.3], [10., 0.3], [65., 0.25], [80., 0.9]])
  d = sampling.generate_dict(pos, ["H0", "Om0"])
  np.testing.assert_array_equal(d["H0"], pos[:, 0])
  d = sampling.generate_dict(pos, ["H0", "Om0"], to_calc=np.array([0, 2]))
  np.testing.assert_array_equal(d["Om0"], [0.3, 0.25])
  assert sampling.generate_dict(pos[0], ["H0", "Om0"]) == {"H0": 70., "Om0": 0.3}
  calls = []

  def fake_like(**kw):
    calls.append({k: np.array(v) for k, v in kw.items()})
    return -0.5 * ((np.asarray(kw["H0"]) - 70.) / 5.) ** 2
  lp = sampling.log_prob_fn(fake_like, ["H0", "Om0"], sampling.uniform_log_prior([[20., 140.], [0.05, 0.6]]))
  out = lp(pos)
  assert len(calls) == 1 and calls[0]["H0"].tolist() == [70., 65.]
  np.testing.assert_allclose(out, [0., -np.inf, -0.5, -np.inf])
  assert lp(pos[1]) == -np.inf and lp(pos[2]) == -0.5


def test_sky_conf_sparse_equals_dense():
  """sky.compute_sky_conf_event (sparse counts) == the reference's dense-map procedure (data.py:246-260)."""
  from chimera_b200 import sky
  rng = np.random.default_rng(5)
  for nside in (8, 64, 256):
    npix = 12 * nside * nside
    for _ in range(20):
      centre = rng.integers(0, npix)
      pe = np.clip(centre + np.round(rng.normal(0, rng.uniform(0.5, 30), 700)).astype(np.int64), 0, npix - 1)
      for level in (0.3, 0.5, 0.9, 0.99, 0.999):
        np.testing.assert_array_equal(sky.compute_sky_conf_event(pe, level, nside),
                                      sky._compute_sky_conf_event_dense(pe, level, nside))


def test_builtin_ensemble_sampler_recovers_a_gaussian():
  """The vectorised stretch-move sampler (stand-in for emcee, absent offline): half-ensembles go through ONE call,
  and a correlated 3-D Gaussian target is recovered (mean and covariance) from 64 walkers x 1500 steps."""
  from chimera_b200 import sampling
  rng = np.random.default_rng(42)
  mean = np.array([70., 0.3, 2.5])
  A = np.array([[4.0, 0.0, 0.0], [0.01, 0.03, 0.0], [0.3, 0.005, 0.4]])
  cov = A @ A.T
  icov = np.linalg.inv(cov)
  calls = []

  def log_prob(p):
    calls.append(p.shape)
    d = p - mean
    return -0.5 * np.einsum("ij,jk,ik->i", d, icov, d)
  prior = sampling.uniform_log_prior([[0., 200.], [-5., 5.], [-50., 50.]])
  p0 = sampling.get_initial_state(64, 3, prior, "gaussian", gaussian_bests=mean, gaussian_sigmas=[1., 0.01, 0.1], rng=rng)
  assert p0.shape == (64, 3) and np.all(np.isfinite(prior(p0)))
  s = sampling.EnsembleSampler(64, 3, log_prob, rng=rng)
  s.run_mcmc(p0, 1500)
  assert all(c == (32, 3) for c in calls[1:]) and calls[0] == (64, 3) and len(calls) == 1 + 2 * 1500
  flat = s.chain[500:].reshape(-1, 3)
  sig = np.sqrt(np.diag(cov))
  assert np.all(np.abs(flat.mean(axis=0) - mean) < 0.2 * sig)
  assert np.all(np.abs(np.cov(flat.T) - cov) < 0.25 * np.outer(sig, sig))
  assert 0.2 < s.acceptance_fraction.mean() < 0.9
  tg = sampling.get_initial_state(10, 3, prior, "truncgauss", priors=[[69., 71.], [0.2, 0.4], [2., 3.]],
                                  gaussian_bests=[80., 0.3, 2.5], gaussian_sigmas=[1., 0.01, 0.1], rng=rng)
  assert np.all((tg[:, 0] >= 69.) & (tg[:, 0] <= 71.))
  with pytest.raises(ValueError):
    sampling.EnsembleSampler(5, 3, log_prob)


def test_bench_arms_share_one_config_object():
  """bench.py: the `config` object of the JSON line is built by ONE function for both arms (`--impl ours` and
  `--impl reference`), for every N the driver launches -- it names the workload, not the implementation."""
  import argparse
  import bench
  for gpus in (1, 2, 8):
    a = argparse.Namespace(config="C3", scaling="strong", nev=0, ninj=0, hyper_groups=1, fp_mode="fp32", options="")
    c = bench.line_config(a, gpus)
    assert c["events_total"] == 1000 and c["n_hyper"] == 256 and c["samples_per_event"] == 5000
    assert c["events_per_gpu"] == 1000 // gpus and c["injections_total"] == 1_000_000
    assert "workload" in c and "l2" in c and not any(k in c for k in ("model", "global_batch", "seq_len"))
    assert c == bench.line_config(a, gpus)
  w = bench.line_config(argparse.Namespace(config="C3", scaling="weak", nev=0, ninj=0, hyper_groups=1, fp_mode="fp32", options=""), 4)
  assert w["events_total"] == 4000 and w["events_per_gpu"] == 1000
