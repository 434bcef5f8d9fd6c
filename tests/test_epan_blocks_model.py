"""CPU check of the algebra behind the fused kernel's unbinned Epanechnikov KDE (csrc/numerator_fused.cu, fu_epan_blocks):
a NumPy float32 model of the block-moment form -- moments {M0, M1, M2} of every 64-sample block about its own centre,
blocks inside the kernel's support through (1 - D^2) M0 + 2 D M1 - M2, straddling blocks summed directly -- against the
float64 pair sums of kde1d (utils/math.py:52-85, kernel 'epan').  Sorted and unsorted samples (the hull decides, not the
order), sparse tails, zero weights."""
import numpy as np
import pytest

f32 = np.float32


def epan_blocks_model(x, w, grid, bw, block=64):
  """float32 throughout, like the kernel: x' = x / bw, grid points gp = g / bw; returns sum_j w_j max(1 - (gp - x'_j)^2, 0)."""
  sf = f32(1.0 / bw)
  xs = (x.astype(f32) * sf).astype(f32)
  ws = w.astype(f32)
  nb = (len(xs) + block - 1) // block
  lo, hi, c, m0, m1, m2 = (np.zeros(nb, f32) for _ in range(6))
  for b in range(nb):
    xb, wb = xs[b * block:(b + 1) * block], ws[b * block:(b + 1) * block]
    lo[b], hi[b] = xb.min(), xb.max()
    c[b] = f32(0.5) * (lo[b] + hi[b])
    d = (xb - c[b]).astype(f32)
    m0[b], m1[b], m2[b] = wb.sum(dtype=f32), (wb * d).sum(dtype=f32), (wb * d * d).sum(dtype=f32)
  out = np.zeros(len(grid), f32)
  for i, g in enumerate(grid):
    gp = f32(g * float(sf))
    acc = f32(0)
    for b in range(nb):
      if hi[b] < gp - f32(1) or lo[b] > gp + f32(1):
        continue
      if lo[b] >= gp - f32(1) and hi[b] <= gp + f32(1):
        D = gp - c[b]
        acc += (f32(1) - D * D) * m0[b] + f32(2) * D * m1[b] - m2[b]
      else:
        xb, wb = xs[b * block:(b + 1) * block], ws[b * block:(b + 1) * block]
        d = gp - xb
        acc += (wb * np.maximum(f32(1) - d * d, f32(0))).sum(dtype=f32)
    out[i] = acc
  return out


def epan_pairs_f64(x, w, grid, bw):
  u = (grid[:, None] - x[None, :]) / bw
  return np.sum(w[None, :] * np.maximum(1.0 - u * u, 0.0), axis=1)


@pytest.mark.parametrize("seed,sort", [(0, True), (1, True), (2, False), (3, True)])
def test_block_moments_match_pair_sums(seed, sort):
  rng = np.random.default_rng(seed)
  n = int(rng.integers(300, 3000))
  x = np.concatenate([rng.normal(0.0, 0.05, n), rng.normal(0.25, 0.02, n // 5), rng.uniform(-0.6, 0.9, 20)])   # bulk, peak, sparse tails
  w = rng.lognormal(0.0, 1.5, x.size)
  w[rng.random(x.size) < 0.2] = 0.0
  if sort:
    o = np.argsort(x)
    x, w = x[o], w[o]
  bw = float(rng.uniform(0.008, 0.05))
  grid = np.linspace(x.min() - 0.1, x.max() + 0.1, 150)
  ref = epan_pairs_f64(x, w, grid, bw)
  got = epan_blocks_model(x, w, grid, bw).astype(np.float64)
  assert np.max(np.abs(got - ref)) < 2e-6 * ref.max()
  assert np.all(got[ref == 0.0] == 0.0)          # outside every sample's support: exactly zero
