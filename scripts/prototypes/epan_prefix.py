"""Prototype (NumPy, CPU) of the Epanechnikov KDE without pair sums, for DESIGN.md section 7 item 4.

kde1d with the Epanechnikov kernel (CHIMERA/utils/math.py:52-85) is
    dens(g) = sum_j w_j 3/4 (1 - ((g - x_j)/bw)^2) [|g - x_j| <= bw] / bw .
The kernel is a quadratic on a compact support, so for SORTED samples the sum over the samples inside the window
[g - bw, g + bw] is a difference of prefix sums of {w, w x, w x^2}:
    dens(g) = 3/(4 bw) [ S0 - (g^2 S0 - 2 g S1 + S2) / bw^2 ],   S_k = P_k[hi(g)] - P_k[lo(g)],
with lo/hi found by binary search: O(Ns + G log Ns) instead of O(Ns G).  x is centred on the grid midpoint before
the sums so that the cancellation in g^2 S0 - 2 g S1 + S2 costs ~(range/bw)^2 ~ 1e3 ulp of fp64, nothing more.
`python scripts/prototypes/epan_prefix.py` checks it against the pair-sum form on random weighted samples."""
import numpy as np


def kde1d_epan_prefix(x, grid, w, bw):
  o = np.argsort(x, kind="stable")
  x, w = x[o], w / np.sum(w)
  w = w[o]
  c = 0.5 * (grid[0] + grid[-1])
  xc, gc = x - c, grid - c
  P0 = np.concatenate([[0.0], np.cumsum(w)])
  P1 = np.concatenate([[0.0], np.cumsum(w * xc)])
  P2 = np.concatenate([[0.0], np.cumsum(w * xc * xc)])
  lo = np.searchsorted(xc, gc - bw, side="left")        # |u| <= 1 is inclusive on both sides
  hi = np.searchsorted(xc, gc + bw, side="right")
  S0, S1, S2 = P0[hi] - P0[lo], P1[hi] - P1[lo], P2[hi] - P2[lo]
  return 0.75 / bw * (S0 - (gc * gc * S0 - 2.0 * gc * S1 + S2) / (bw * bw))


def kde1d_epan_pairs(x, grid, w, bw):
  w = w / np.sum(w)
  u = (grid[:, None] - x) / bw
  return np.sum(w * np.where(np.abs(u) <= 1, 0.75 * (1 - u ** 2), 0.0), axis=-1) / bw


if __name__ == "__main__":
  rng = np.random.default_rng(0)
  worst = 0.0
  for _ in range(200):
    n = int(rng.integers(50, 6000))
    x = rng.normal(rng.uniform(0.05, 1.5), rng.uniform(0.005, 0.2), n)
    w = rng.random(n) * (rng.random(n) > 0.2)
    neff = np.sum(w) ** 2 / np.sum(w ** 2)
    bw = neff ** -0.2 * np.std(x) * rng.uniform(0.3, 3.0)
    grid = np.linspace(x.min() - 2 * np.std(x), x.max() + 2 * np.std(x), 150)
    a, b = kde1d_epan_prefix(x, grid, w, bw), kde1d_epan_pairs(x, grid, w, bw)
    worst = max(worst, float(np.max(np.abs(a - b)) / np.max(b)))
  print(f"max |prefix - pairs| / peak over 200 random cases: {worst:.2e}")
  assert worst < 1e-10
