"""Independent checks of the oracle: besides the fixtures written by the reference's own source
(test_oracle_golden.py), the restatement is held to closed-form results and normalisations that do not depend on
any implementation: Einstein-de Sitter distances, the inverse dL -> z, the 1/H0 scaling, unit integrals of the mass
model and of both KDE kernels, and a selection function whose Monte-Carlo weights are constant by construction."""
import numpy as np
import pytest

from oracle import chimera_oracle as orc

_trapz = np.trapezoid if hasattr(np, "trapezoid") else np.trapz


def test_einstein_de_sitter_distances():
  """Om0 = 1, flat: dC = 2 dH (1 - (1+z)^-1/2), dL = (1+z) dC, dVc/dz = 4 pi dH dC^2 / E, E = (1+z)^1.5."""
  c = orc.make_cosmo("flrw", H0=70., Om0=1.0, z_max=5.)
  z = np.array([1e-3, 0.05, 0.3, 1.0, 2.5, 4.9])
  dH = 299792.458e-3 / 70.
  dC = 2 * dH * (1 - (1 + z) ** -0.5)
  np.testing.assert_allclose(orc.E_at_z(c, z), (1 + z) ** 1.5, rtol=1e-14)
  np.testing.assert_allclose(orc.dL_at_z(c, z), (1 + z) * dC, rtol=2e-5)          # 1500-knot trapezoid table
  np.testing.assert_allclose(orc.dVcdz_at_z(c, z), 4 * np.pi * dH * dC ** 2 / (1 + z) ** 1.5, rtol=5e-5)
  np.testing.assert_allclose(orc.ddLdz_at_z(c, z), dC + dH * (1 + z) / (1 + z) ** 1.5, rtol=2e-5)


@pytest.mark.parametrize("model,kw", [("flrw", dict(H0=67., Om0=0.31)), ("flrw", dict(H0=80., Om0=0.3, Ok0=0.05)),
                                      ("mg_flrw", dict(H0=70., Om0=0.25, Xi0=1.5, n=2.0))])
def test_distance_inverse_and_h0_scaling(model, kw):
  c = orc.make_cosmo(model, z_max=5., **kw)
  z = np.geomspace(1e-3, 4.5, 60)
  # both directions interpolate linearly on the 1500 log-spaced knots (1.6 % apart): second-order error ~3e-5
  np.testing.assert_allclose(orc.z_from_dGW(c, orc.dL_at_z(c, z)), z, rtol=5e-5)
  c2 = orc.make_cosmo(model, z_max=5., **{**kw, "H0": 2 * kw["H0"]})
  np.testing.assert_allclose(orc.dL_at_z(c2, z), 0.5 * orc.dL_at_z(c, z), rtol=1e-13)


@pytest.mark.parametrize("model", ["tpl", "bpl", "plp"])
def test_mass_model_integrates_to_one(model):
  m = orc.make_mass(model)
  m1 = np.geomspace(m["m_low"] * (1 + 1e-6), m["m_high"], 1600)      # p(m2|m1) is 0/0 at m1 = m_low exactly
  inner = np.empty_like(m1)
  for i, a in enumerate(m1):
    m2 = np.linspace(m["m_low"], a, 400)
    with np.errstate(all="ignore"):
      inner[i] = _trapz(np.nan_to_num(orc.p_m1m2(m, np.full_like(m2, a), m2)), m2)
  assert abs(_trapz(inner, m1) - 1.0) < 5e-3


@pytest.mark.parametrize("kernel", ["gauss", "epan"])
@pytest.mark.parametrize("bw", [None, "silverman", 0.3])
def test_kde1d_is_a_density(kernel, bw):
  rng = np.random.default_rng(3)
  x = rng.normal(0.4, 0.05, 3000)
  w = rng.random(3000)
  grid = np.linspace(0.0, 0.8, 4001)
  assert abs(_trapz(orc.kde1d(x, grid, w, kernel=kernel, bw_method=bw), grid) - 1.0) < 1e-4


def test_selection_function_with_constant_weights():
  """If p_draw is the population's own detector-frame rate divided by a constant C, every Monte-Carlo weight equals
  C: xi = C n_det / N_inj exactly, the variance term vanishes up to rounding and the N_eff gate stays open."""
  from chimera_b200 import synth
  inj, N_inj = synth.make_injections(20000, seed=11)
  pop = orc.make_pop(orc.make_cosmo("flrw", H0=70., Om0=0.25, z_max=5.), orc.make_mass("plp"),
                     orc.make_rate("madau_dickinson"))
  with np.errstate(all="ignore"):
    rate = orc.pop_rate_det_inj(pop, inj)
  keep = np.isfinite(rate) & (rate > 0)
  inj = {k: v[keep] for k, v in inj.items()}
  C = 3.7
  inj["p_draw"] = rate[keep] / C
  Nexp, xi, neff = orc.N_exp(pop, inj, N_inj, 5.)
  np.testing.assert_allclose(xi, C * keep.sum() / N_inj, rtol=1e-12)
  assert Nexp == pop["Tobs"] * xi and neff > 5.
