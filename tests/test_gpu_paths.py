"""GPU tests of the alternative code paths of the fp32 fast path: fused kernel (odd sample counts, CHB_SPLIT=0),
hyper-point batching of the stage buffer (CHB_STAGE_GB), windows switched off (CHB_KDE_WIN=0, short z grids) and
the one-MUFU-per-pair pair sums (CHB_KDE_DIRECT=1) -- all against the NumPy oracle on the same seeded inputs."""
import os
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


@pytest.mark.parametrize("ns,nz,env", [
  (4097, 300, {}),                          # odd sample count: fused kernel, recurrence without windows
  (4096, 300, {"CHB_SPLIT": "0"}),          # fused kernel with the windowed KDE
  (4096, 300, {"CHB_KDE_WIN": "0"}),        # split kernels, full-grid recurrence
  (4096, 300, {"CHB_KDE_DIRECT": "1"}),     # one MUFU.EX2 per pair
  (4096, 300, {"CHB_STAGE_GB": "0.0003"}),  # stage buffer for ONE hyper-point at a time: three batches
  (4096, 48, {}),                           # z grid too short for the chunk tables: no windows
  (4096, 300, {}),                          # default: split + windows + packed FP32
])
def test_fast_path_variants_match_oracle(cb, ns, nz, env):
  th, zg, pop, sel, H0, ref = _case(cb, ns, nz)
  old = {k: os.environ.get(k) for k in env}
  os.environ.update(env)
  try:
    like = cb.hyperlikelihood(th, zg, pop, sel, kernel="gauss", binning=False, fp_mode="fp32")
    lle = like.compute_all(H0=H0)[0]
    lle2 = like.compute_all(H0=H0)[0]
  finally:
    for k, v in old.items():
      if v is None:
        os.environ.pop(k, None)
      else:
        os.environ[k] = v
  _check(lle, ref)
  np.testing.assert_array_equal(lle, lle2)      # bit-reproducible from call to call


@pytest.mark.parametrize("kind,kernel,binning", [("approximate", "gauss", False), ("marginalized", "epan", False),
                                                 ("marginalized", "epan", True), ("full", "gauss", False)])
def test_fused_kernel_pixelated_kinds(cb, kind, kernel, binning):
  """CHB_SPLIT=0 routes the pixelated kinds through the fused kernel (numerator_f32_kernel<KG, 0>); it must agree
  with the oracle like the split form does (tests/test_gpu_parity.py)."""
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
  old = os.environ.get("CHB_SPLIT")
  os.environ["CHB_SPLIT"] = "0"
  try:
    like = cb.hyperlikelihood(cb.theta_pe_det(**kw), zg, pop, sel, kind_p_gw3d=kind, kernel=kernel, binning=binning,
                              num_bins=200, fp_mode="fp32")
    lle = like.compute_all(H0=H0)[0]
  finally:
    if old is None:
      os.environ.pop("CHB_SPLIT", None)
    else:
      os.environ["CHB_SPLIT"] = old
  _check(lle, ref, tol=1e-4)
