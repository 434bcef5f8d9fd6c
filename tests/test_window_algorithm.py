"""CPU checks of the windowed recurrence KDE ALGORITHM (tests/window_model.py mirrors csrc/kde_win.cuh): the windows
never drop a term above 2^-30 of the largest term at a grid point, and the density keeps its RELATIVE accuracy at
every grid point -- bulk, far tails, gaps between modes, zero-weight samples -- against exact fp64 sums."""
import numpy as np
import pytest

from window_model import kde_window


def _exact(z, w, lb, step, G, bw):
  """fp64 terms on the exact grid: log2-terms matrix (G, n) and the sums."""
  s = 0.8493218002880191 / bw
  g = lb + step * np.arange(G)
  with np.errstate(divide="ignore"):
    lt = np.log2(w / w.sum())[None, :] - ((g[:, None] - z[None, :]) * s) ** 2
  return lt, np.sum(np.exp2(lt), axis=1)


def _case(rng, kind, n=5000):
  if kind == "bulk":
    z = rng.normal(0.5, 0.05, n)
  elif kind == "bimodal":
    z = np.concatenate([rng.normal(0.30, 0.02, n // 2), rng.normal(0.55, 0.02, n - n // 2)])   # gap of ~7 bandwidths
  elif kind == "heavy_tail":
    z = 0.4 + 0.03 * rng.standard_t(3, n)
    z = z[(z > 0.01) & (z < 2.0)]
  else:  # sparse weights: most samples outside the mass support
    z = rng.normal(0.5, 0.05, n)
  z = np.sort(z)
  w = rng.random(z.size) ** 3 * 10 ** rng.uniform(-3, 0, z.size)
  if kind == "sparse":
    w = w * (rng.random(z.size) < 0.15)
  return z, w


def _run(z, w, lb, ub, G=150):
  neff = w.sum() ** 2 / (w ** 2).sum()
  bw = neff ** -0.2 * z.std()
  step = (ub - lb) / (G - 1)
  out = kde_window(z, w, lb, step, G, bw)
  if out is None:
    return None
  dens, info = out
  lt, exact = _exact(z, w, lb, step, G, bw)
  return dens, info, lt, exact, lb + step * np.arange(G)


@pytest.mark.parametrize("kind", ["bulk", "bimodal", "heavy_tail", "sparse"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_window_model_accuracy_and_containment(kind, seed):
  """The guarantees of the design (kde_win.cuh header): (1) no term within 2^-30 of the largest term at a grid point
  is left outside its chunk's window; (2) absolute error below 2e-6 of the peak everywhere; (3) relative error at fp32
  level wherever the density is above 1e-8 of the peak; (4) the same relative accuracy at every grid point OUTSIDE the
  span of the samples (the tails that decide the likelihood of an event sitting beyond a catalogue or rate edge)."""
  rng = np.random.default_rng(100 * seed + len(kind))
  z, w = _case(rng, kind)
  out = _run(z, w, max(z.min() - 2 * z.std(), 1e-8), z.max() + 2 * z.std())
  if out is None:
    pytest.skip("the plan refuses windows for this bandwidth (window ~ whole grid)")
  dens, info, lt, exact, g = out
  chunk = info["chunk"]
  big = lt >= (lt.max(axis=1, keepdims=True) - 30.0)
  for gi, j in zip(*np.nonzero(big)):
    ia, ib = info["win"][j // chunk]
    assert ia <= gi <= ib
  assert np.abs(dens - exact).max() < 2e-6 * exact.max()
  core = exact > 1e-8 * exact.max()
  assert (np.abs(dens[core] - exact[core]) / exact[core]).max() < 3e-5
  live = w > 0
  outside = ((g < z[live].min()) | (g > z[live].max())) & (exact > 2.0 ** -1000)
  if outside.any():
    assert (np.abs(dens[outside] - exact[outside]) / exact[outside]).max() < 1e-4
  if kind != "bimodal":                                 # (wide bandwidth relative to the grid: little to save there)
    assert info["pairs"] < 0.85 * z.size * 150         # and the windows do save work


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_window_model_deep_tails(seed):
  """A grid that reaches 7 sigma beyond the samples on both sides: densities down to 1e-280 of the peak keep their
  relative accuracy (reversed runs + integer exponent offsets), where a plain fp32 sum returns exactly 0."""
  rng = np.random.default_rng(7 + seed)
  z = np.sort(rng.normal(0.6, 0.03, 5000))
  w = rng.random(5000) ** 2
  out = _run(z, w, z.min() - 7 * z.std(), z.max() + 7 * z.std(), G=300)
  assert out is not None
  dens, info, lt, exact, g = out
  outside = ((g < z.min()) | (g > z.max())) & (exact > 2.0 ** -1000)
  assert exact[outside].min() < 1e-250 * exact.max()
  assert (np.abs(dens[outside] - exact[outside]) / exact[outside]).max() < 2e-4
  core = exact > 1e-8 * exact.max()
  assert (np.abs(dens[core] - exact[core]) / exact[core]).max() < 3e-5


@pytest.mark.parametrize("mult", [0.05, 0.08, 0.12, 0.18, 0.25, 0.35, 0.5])
def test_window_model_all_tilings(mult):
  """Scalar bandwidths from 0.05 to 0.5 sigma walk through the tilings (R, LPS) = (4,4) (4,8) (8,4) (12,4) (16,4) (12,8)
  (16,8).  Along a run the terms fall by q = 2^-(2 h d + h^2) per grid point, so fp32 carries a run's far end only
  while 2 h d (R-1) stays below ~200 bits: the tail guarantee is asserted down to 1e-120 of the peak, which every
  tiling reaches (the fine grids of the default bandwidth reach 1e-280, test_window_model_deep_tails)."""
  rng = np.random.default_rng(int(mult * 1000))
  seen = set()
  for n in (5000, 2048, 20000):
    z = np.sort(rng.normal(0.5, 0.05, n))
    w = rng.random(n) ** 2 * (rng.random(n) > 0.1)
    bw = mult * z.std()
    lb, ub = z.min() - 2 * z.std(), z.max() + 2 * z.std()
    step = (ub - lb) / 149
    out = kde_window(z, w, lb, step, 150, bw)
    if out is None:
      continue
    dens, info = out
    seen.add((info["R"], info["LPS"]))
    lt, exact = _exact(z, w, lb, step, 150, bw)
    g = lb + step * np.arange(150)
    live = w > 0
    assert np.abs(dens - exact).max() < 3e-6 * exact.max()
    core = exact > 1e-8 * exact.max()
    assert (np.abs(dens[core] - exact[core]) / exact[core]).max() < 1e-4
    outside = ((g < z[live].min()) | (g > z[live].max())) & (exact > 1e-120 * exact.max())
    if outside.any():
      # (fp32 rounding of d^2 at d ~ 20 scaled units: 2 d ulp(d) ln 2 ~ 1e-4)
      assert (np.abs(dens[outside] - exact[outside]) / exact[outside]).max() < 5e-4
  assert seen


@pytest.mark.parametrize("n", [1000, 700, 512, 449])
@pytest.mark.parametrize("kind", ["bulk", "heavy_tail", "sparse"])
def test_window_model_short_events(n, kind):
  """Events with few samples (walker batches with ~1000 samples per event): the plan halves the loop iterations per pass
  (32 -> 16 -> 8) until there are at least 8 chunks instead of refusing windows; the same guarantees hold with the
  short chunks.  Below 449 samples (8 chunks of 64) the plan still refuses and the kernel sums directly."""
  from window_model import plan
  rng = np.random.default_rng(n + len(kind))
  z, w = _case(rng, kind, n=n)
  out = _run(z, w, max(z.min() - 2 * z.std(), 1e-8), z.max() + 2 * z.std())
  if out is None:
    pytest.skip("the plan refuses windows for this bandwidth (window ~ whole grid)")
  dens, info, lt, exact, g = out
  assert info["nchunks"] >= 8 and info["chunk"] <= 128
  big = lt >= (lt.max(axis=1, keepdims=True) - 30.0)
  for gi, j in zip(*np.nonzero(big)):
    ia, ib = info["win"][j // info["chunk"]]
    assert ia <= gi <= ib
  assert np.abs(dens - exact).max() < 2e-6 * exact.max()
  core = exact > 1e-8 * exact.max()
  assert (np.abs(dens[core] - exact[core]) / exact[core]).max() < 3e-5
  assert plan(150, 300, 0.3) is None                   # 300 samples: no window plan
