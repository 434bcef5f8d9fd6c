"""CPU check of the algebra behind the sample-block windows of the 'full' 3-D KDE (csrc/numerator_f32.cu, KG = 2, MODE 2):
a NumPy model of (1) the whitening with the variables ordered (dec, ra, z), whose third coordinate depends on z alone and
whose quadratic form equals the one of any other order, and (2) the skip rule -- per tile of evaluation points a LOWER
bound M of the largest single term at each point (from every 32nd sample, minimum over the tile), per 64-sample block an
UPPER bound from its heaviest weight and its nearest approach in that one coordinate; blocks below 2^-24 of M are
skipped.  Checked: what is dropped is below nblk * 2^-24 of the exact sum at EVERY point, far tails included."""
import numpy as np


def whiten(cov_inv, order):
  """Upper factor U with x^T cov_inv x = |U x|^2 for the variables taken in `order` (Cholesky of the permuted matrix)."""
  P = np.eye(3)[list(order)]
  L = np.linalg.cholesky(P @ cov_inv @ P.T)
  return L.T @ P


def test_quadratic_form_is_order_independent_and_last_coordinate_is_z():
  rng = np.random.default_rng(0)
  A = rng.normal(size=(3, 3))
  cov_inv = np.linalg.inv(A @ A.T + 0.1 * np.eye(3))
  x = rng.normal(size=(100, 3))                      # columns (z, ra, dec)
  U0, U1 = whiten(cov_inv, (0, 1, 2)), whiten(cov_inv, (2, 1, 0))
  q = np.einsum("ni,ij,nj->n", x, cov_inv, x)
  np.testing.assert_allclose(np.sum((x @ U0.T) ** 2, axis=1), q, rtol=1e-12)
  np.testing.assert_allclose(np.sum((x @ U1.T) ** 2, axis=1), q, rtol=1e-12)
  assert abs(U1[2, 1]) < 1e-14 and abs(U1[2, 2]) < 1e-14 and U1[2, 0] > 0      # y2 = l22 * z


def test_block_windows_drop_less_than_their_threshold_everywhere():
  rng = np.random.default_rng(1)
  ns, sb, t2 = 4096, 64, 24.0
  z = np.sort(rng.normal(0.4, 0.05, ns))             # sorted by dL <=> sorted by z
  ra, dec = rng.normal(1.0, 0.02, ns) + 0.3 * (z - 0.4), rng.normal(-0.3, 0.03, ns)
  w = rng.lognormal(0.0, 1.0, ns)
  w[rng.random(ns) < 0.3] = 0.0
  w /= w.sum()
  X = np.stack([z, ra, dec], axis=1)
  mu = w @ X
  cov = (w[:, None] * (X - mu)).T @ (X - mu) / (1.0 - np.sum(w ** 2))
  f = (1.0 / np.sum(w ** 2)) ** (-1.0 / 7.0)
  U = whiten(np.linalg.inv(cov) / f ** 2, (2, 1, 0)) * np.sqrt(np.log2(np.e) / 2.0)    # terms are w 2^-|y - q|^2
  Y = (X - mu) @ U.T
  assert np.all(np.diff(Y[:, 2]) >= 0)               # sorted in the z-only coordinate
  # evaluation points: 20 pixels x 150 z values, pixel-fastest, z grid far wider than the samples (far tails)
  zk = np.linspace(0.05, 0.9, 150)
  # (pixels inside the event's sky localisation, as the credible-region pixelisation gives them)
  pix = np.stack([rng.normal(1.0, 0.008, 20), rng.normal(-0.3, 0.012, 20)], axis=1)
  pts = np.array([[zz, p[0], p[1]] for zz in zk for p in pix])
  Q = (pts - mu) @ U.T
  lw = np.where(w > 0, np.log2(np.where(w > 0, w, 1.0)), -np.inf)
  nblk = ns // sb
  blo, bhi = Y[::sb, 2], Y[sb - 1::sb, 2]
  bw = np.array([lw[b * sb:(b + 1) * sb].max() for b in range(nblk)])
  worst, skipped = 0.0, 0
  for t in range(0, len(Q), 128):
    q = Q[t:t + 128]
    e_sub = np.sum((Y[None, ::32, :] - q[:, None, :]) ** 2, axis=2)
    M = np.min(np.max(lw[None, ::32] - e_sub, axis=1))
    dist = np.maximum(np.maximum(blo - q[:, 2].max(), q[:, 2].min() - bhi), 0.0)
    keep = bw + 6.0 - dist ** 2 >= M - t2
    skipped += int(np.sum(~keep))
    e = np.sum((Y[None, :, :] - q[:, None, :]) ** 2, axis=2)
    terms = w[None, :] * np.exp2(-e)
    exact = terms.sum(axis=1)
    mask = np.repeat(keep, sb)
    got = terms[:, mask].sum(axis=1)
    ok = exact > 0
    worst = max(worst, float(np.max((exact[ok] - got[ok]) / exact[ok])))
  assert skipped > 0.05 * nblk * (len(Q) // 128)     # the windows do skip work (37 % of the blocks on the C4 bench workload) ...
  assert worst < nblk * 2.0 ** -t2                  # ... and what they drop is below the stated bound at every point
