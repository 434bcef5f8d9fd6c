"""GPU tests of the alternative code paths of the fp32 mode, selected per handle with chb_set_option (`options=`):
the fused one-kernel form of the 1-D kinds (default) with and without windows, the round-1 split / MODE-0 kernels
(`fused=0`), odd sample counts, hyper-point batching of the stage buffer, one-MUFU-per-pair sums -- all against the
NumPy oracle on the same seeded inputs; plus two handles of different shapes interleaved in one process."""
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


def _case(cb, ns, nz, nev=10, seed=400):
  from oracle import chimera_oracle as orc
  from chimera_b200 import synth
  ev = synth.make_events(nev, ns, seed=seed, sky=False)
  zg = synth.make_z_grids(ev["dL"], z_int_res=nz, H0_prior=(40., 120.))
  inj, N_inj = synth.make_injections(20000, seed=seed + 1)
  th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")})
  sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
  pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson())
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  opts = orc.make_opts(None, "gauss", None, 2.0, False, 200, 2.0)
  H0 = np.array([55., 70., 88.])
  ref = np.array([orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., None, H0=float(h))[0] for h in H0])
  return th, zg, pop, sel, H0, ref


def _check(lle, ref, tol=2e-5):
  fin = np.isfinite(ref) & (np.abs(ref) < 1e300)
  assert np.array_equal(fin, np.isfinite(lle) & (np.abs(lle) < 1e300))
  err = np.max(np.abs(lle[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0))
  assert err < tol, err


@pytest.mark.parametrize("ns,nz,opt", [
  (4096, 300, {}),                                   # default: fused kernel, windowed recurrence on the raw stage
  (4096, 300, {"kde_win_t2": 30}),                   # round 1's window threshold (default: 24 bits)
  (4096, 300, {"kde_win": 0}),                       # fused kernel, direct pair sums (no window plan)
  (4096, 48, {}),                                    # z grid too short for a window plan
  (700, 300, {}),                                    # few samples: ragged last block, 64-sample chunks
  (300, 300, {}),                                    # too few samples for 8 chunks -> direct sums
  (4096, 300, {"fused_nt": 128}),                    # the 128-thread instantiation (default for Ns <= 2048) on a long event
  (700, 300, {"fused_nt": 256}),                     # ... and the 256-thread one on a short event
  (700, 300, {"fused_nt": 64}),                      # 64-thread instantiation
  (4097, 300, {}),                                   # odd sample count: round-1 MODE-0 kernel, recurrence without windows
  (4096, 300, {"fused": 0}),                         # round-1 split kernels + windows
  (4096, 300, {"fused": 0, "split": 0}),             # round-1 MODE-0 kernel with the windowed KDE
  (4096, 300, {"fused": 0, "kde_win": 0}),           # split kernels, full-grid recurrence
  (4096, 300, {"fused": 0, "kde_direct": 1}),        # one MUFU.EX2 per pair
  (4096, 300, {"fused": 0, "stage_gb": 0.0003}),     # stage buffer for ONE hyper-point at a time: three batches
  (4096, 300, {"zterms_gb": 0.00004}),               # z-grid terms for ONE hyper-point at a time: the fused kernel in three launches
])
def test_fast_path_variants_match_oracle(cb, ns, nz, opt):
  th, zg, pop, sel, H0, ref = _case(cb, ns, nz)
  like = cb.hyperlikelihood(th, zg, pop, sel, kernel="gauss", binning=False, fp_mode="fp32", options=opt)
  lle = like.compute_all(H0=H0)[0]
  lle2 = like.compute_all(H0=H0)[0]
  _check(lle, ref)
  np.testing.assert_array_equal(lle, lle2)      # bit-reproducible from call to call


@pytest.mark.parametrize("kernel,binning,bw", [("epan", True, None), ("gauss", True, "silverman"), ("epan", False, 0.3),
                                               ("gauss", False, "silverman"), ("gauss", False, 0.25)])
def test_fused_kernel_kde_options(cb, kernel, binning, bw):
  """Every KDE option of the 1-D kinds through the fused kernel and through the round-1 kernels, against the oracle."""
  from oracle import chimera_oracle as orc
  th, zg, pop, sel, H0, _ = _case(cb, 2048, 200, nev=8, seed=431)
  from chimera_b200 import synth
  ev = synth.make_events(8, 2048, seed=431, sky=False)
  inj, N_inj = synth.make_injections(20000, seed=432)
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  opts = orc.make_opts(None, kernel, bw, 2.0, binning, 200, 2.0)
  ref = np.array([orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., None, H0=float(h))[0] for h in H0])
  for opt in ({}, {"fused": 0}, {"epan_blocks": 0}):
    like = cb.hyperlikelihood(th, zg, pop, sel, kernel=kernel, bw_method=bw, binning=binning, num_bins=200,
                              fp_mode="fp32", options=opt)
    _check(like.compute_all(H0=H0)[0], ref, tol=1e-4)


@pytest.mark.parametrize("bw,ns", [(None, 4096), ("silverman", 5000), (0.3, 700), (0.05, 4096)])
def test_epanechnikov_unbinned_block_moments(cb, bw, ns):
  """Unbinned Epanechnikov KDE of the fused kernel (block moments of the sorted samples, fu_epan_blocks) against the
  oracle's pair sums and against the direct pair sums of the same kernel (`epan_blocks=0`); wide and narrow bandwidths,
  a ragged last block."""
  from oracle import chimera_oracle as orc
  from chimera_b200 import synth
  nev = 8
  ev = synth.make_events(nev, ns, seed=451, sky=False)
  zg = synth.make_z_grids(ev["dL"], z_int_res=300, H0_prior=(40., 120.))
  inj, N_inj = synth.make_injections(20000, seed=452)
  th = cb.theta_pe_det(**{k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")})
  sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
  pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson())
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  opts = orc.make_opts(None, "epan", bw, 2.0, False, 200, 2.0)
  H0 = np.array([55., 70., 88.])
  ref = np.array([orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., None, H0=float(h))[0] for h in H0])
  out = {}
  for name, opt in (("blocks", {}), ("direct", {"epan_blocks": 0})):
    like = cb.hyperlikelihood(th, zg, pop, sel, kernel="epan", bw_method=bw, binning=False, fp_mode="fp32", options=opt)
    out[name] = like.compute_all(H0=H0)[0]
    _check(out[name], ref, tol=2e-5)
  _check(out["blocks"], out["direct"], tol=5e-6)


def test_two_handles_interleaved(cb):
  """Handles with different shared-memory footprints alternate in one process (the kernels' dynamic shared-memory
  attribute is process-wide): neither may invalidate the other's launches, in either path."""
  big = _case(cb, 4096, 300)
  small = _case(cb, 512, 64, nev=6, seed=77)
  for opt in ({}, {"fused": 0}):
    la = cb.hyperlikelihood(big[0], big[1], big[2], big[3], kernel="gauss", binning=False, fp_mode="fp32", options=opt)
    lb = cb.hyperlikelihood(small[0], small[1], small[2], small[3], kernel="gauss", binning=False, fp_mode="fp32", options=opt)
    for _ in range(2):
      _check(la.compute_all(H0=big[4])[0], big[5])
      _check(lb.compute_all(H0=small[4])[0], small[5])


@pytest.mark.parametrize("kind,kernel,binning", [("approximate", "gauss", False), ("marginalized", "epan", False),
                                                 ("marginalized", "epan", True), ("full", "gauss", False)])
def test_fused_kernel_pixelated_kinds(cb, kind, kernel, binning):
  """`split=0` routes the pixelated kinds through the one-kernel MODE-0 form (numerator_f32_kernel<KG, 0>); it must
  agree with the oracle like the split form does (tests/test_gpu_parity.py)."""
  from oracle import chimera_oracle as orc
  from test_gpu_parity import _synthetic
  ev, zg, inj, N_inj, cat = _synthetic(12, 1500, 120, 20000, True, seed=107)
  kw = {k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides", "pixels_opt_nsides", "ra_pix",
                           "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside")}
  gcat = cb.pixelated_catalog(cb.dVdz_completeness(cat["z_range"]), p_cat=cat["p_cat"], P_compl=cat["P_compl"])
  pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=gcat)
  sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"), catalog=cat)
  opts = orc.make_opts(kind, kernel, None, 2.0, binning, 200, 2.0)
  H0 = np.array([60., 75.])
  ref = np.array([orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., ev["neff_pixels"], H0=float(h))[0] for h in H0])
  like = cb.hyperlikelihood(cb.theta_pe_det(**kw), zg, pop, sel, kind_p_gw3d=kind, kernel=kernel, binning=binning,
                            num_bins=200, fp_mode="fp32", options={"fused": 0, "split": 0})
  lle = like.compute_all(H0=H0)[0]
  _check(lle, ref, tol=1e-4)


@pytest.mark.parametrize("bw", [None, "silverman", 0.35])
def test_marginalized_binned_fused_vs_split_and_oracle(cb, bw):
  """'marginalized' + binning (the reference's default options): the fused warp-per-pixel kernel with the prefix-sum
  Epanechnikov (default) and the round-1 split kernels (`fused=0`) against the oracle, p_gw arrays included."""
  from oracle import chimera_oracle as orc
  from test_gpu_parity import _synthetic
  ev, zg, inj, N_inj, cat = _synthetic(12, 2000, 160, 20000, True, seed=211)
  kw = {k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior", "ra", "dec", "opt_nsides", "pixels_opt_nsides", "ra_pix",
                           "dec_pix", "gw_loc2d_pdf", "pixels_pe_opt_nside")}
  gcat = cb.pixelated_catalog(cb.dVdz_completeness(cat["z_range"]), p_cat=cat["p_cat"], P_compl=cat["P_compl"])
  pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson(), gal_cat=gcat)
  sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"), catalog=cat)
  opts = orc.make_opts("marginalized", "epan", bw, 2.0, True, 200, 2.0)
  H0 = np.array([58., 70., 83.])
  ref = np.array([orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., ev["neff_pixels"], H0=float(h))[0] for h in H0])
  pgw_ref = orc.p_gw3dmarg(orc.pop_update(pop0, H0=70.), ev, zg, opts)
  for opt in ({}, {"fused": 0}):
    like = cb.hyperlikelihood(cb.theta_pe_det(**kw), zg, pop, sel, kind_p_gw3d="marginalized", kernel="epan", bw_method=bw,
                              binning=True, num_bins=200, fp_mode="fp32", options=opt)
    _check(like.compute_all(H0=H0)[0], ref, tol=1e-4)
    pgw = like.p_gw3dmarg(pop.update(H0=70.))
    scale = np.nanmax(np.abs(pgw_ref))
    for e in range(pgw_ref.shape[0]):
      n = int(ev["neff_pixels"][e])
      # (fp32 mode: a sample whose fp32 redshift sits on a bin edge may land in the neighbouring bin; in a sparsely
      # populated pixel that moves a visible fraction of the weight by one bin width -- hence per-element 2 %, while the
      # integrated likelihoods above agree to 1e-4)
      np.testing.assert_allclose(pgw[e, :n], pgw_ref[e, :n], rtol=2e-2, atol=2e-4 * scale)
