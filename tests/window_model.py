"""NumPy model of the windowed recurrence KDE of chimera_b200/csrc/kde_win.cuh (float32 arithmetic where the kernel
uses it), used by tests/test_window_algorithm.py to check the ALGORITHM on the CPU: chunk summaries, the lower bound
M(g), the per-chunk hull of needed grid points, runs walked away from the samples, integer exponent offsets, and the
q, q^2, q^3, q^4 recurrence.  It mirrors the kernel step by step; it is not used by the product."""
import numpy as np

f32 = np.float32
T2 = f32(30.0)


def plan(G, n, h, iters=32, span=5, maxr=16):
  if not (h > 0) or h > 1.8:
    return None
  wn = 2 * int(np.ceil(6.2 / h)) + span
  if 10 * wn > 7 * G:
    return None
  rmax = min(maxr, 1 + int(5.5 / h))
  R = L = 0
  for r, l in zip([4, 8, 4, 12, 16, 8, 4, 12, 16, 8, 12, 16], [2, 2, 4, 2, 2, 4, 8, 4, 4, 8, 8, 8]):
    if r > rmax:
      continue
    if R == 0 or R * L < wn:
      if R == 0 or r * l > R * L or (r * l == R * L and r > R):
        R, L = r, l
    elif r * l == R * L and r > R:
      R, L = r, l
  if R == 0:
    return None
  q = 4 * (32 // L)
  it = iters
  while True:                # few samples: shorter chunks rather than no windows (kde_win.cuh win_plan)
    chunk = max(it * (32 // L), ((n + 31) // 32 + q - 1) // q * q)
    nch = (n + chunk - 1) // chunk
    if nch >= 8 or it <= 8:
      break
    it //= 2
  return (R, L, chunk, nch) if nch >= 8 else None


def ex2(a):
  """ex2.approx.ftz.f32: flush-to-zero below 2^-126."""
  a = np.asarray(a, dtype=f32)
  with np.errstate(over="ignore", under="ignore"):
    r = np.exp2(a.astype(np.float64))
  r = np.where(a < -126, 0.0, r)
  return r.astype(f32)


def kde_window(z, w, lb, step, G, bw, iters=32, stats=None):
  """dens[g] = sum_j (w_j / W) 2^-((g' - x'_j)^2) on the grid lb + g step, as the kernel computes it.
  Returns (dens (G,) float64, info dict) or None when the plan refuses windows."""
  n = z.size
  s = 0.8493218002880191 / bw
  sf = f32(s)
  c = 0.5 * (lb + (lb + (G - 1) * step))
  gfirst, hd = (lb - c) * float(sf), step * float(sf)
  h = f32(hd)
  pl = plan(G, n, float(h), iters)
  if pl is None:
    return None
  R, LPS, chunk, nch = pl
  # ---- phase A
  c_hi = f32(c); c_lo = f32(c - float(c_hi))
  W = float(np.sum(w, dtype=np.float64))
  lg2invW = -np.log2(f32(W)).astype(f32)
  x = ((z.astype(f32) - c_hi) - c_lo) * sf
  live = w.astype(f32) > 0
  with np.errstate(divide="ignore"):
    lw = np.where(live, np.log2(w.astype(f32)).astype(f32) + lg2invW, f32(-np.inf)).astype(f32)
  summ = []
  for ck in range(nch):
    sl = slice(ck * chunk, min(n, (ck + 1) * chunk))
    lv = live[sl]
    if lv.any():
      xs, ls = x[sl][lv], lw[sl][lv]
      k = int(np.argmax(ls))
      summ.append((xs.min(), xs.max(), ls[k], xs[k]))
    else:
      summ.append((f32(np.inf), f32(-np.inf), f32(-np.inf), f32(0)))
  summ = np.array(summ, dtype=f32)
  # ---- phase B
  lgchunk = f32(np.log2(f32(chunk)))
  gp = (np.arange(G, dtype=f32) * h + f32(gfirst)).astype(f32)
  d = gp[:, None] - summ[None, :, 3]
  M = np.max(-d * d + summ[None, :, 2], axis=1)
  dist = np.maximum(np.maximum(summ[None, :, 0] - gp[:, None], gp[:, None] - summ[None, :, 1]), 0)
  need = (-dist * dist + (summ[None, :, 2] + lgchunk) >= (M - T2)[:, None]) & (summ[None, :, 2] > -np.inf)
  win = []
  for ck in range(nch):
    g = np.flatnonzero(need[:, ck])
    win.append((int(g.min()), int(g.max())) if g.size else (G, -1))
  # ---- phase C
  cr = np.array([2.0 ** (-(float(h) * float(h)) * (r * (r - 1))) for r in range(16)], dtype=f32)
  dens = np.zeros(G)
  pairs = 0
  for ck in range(nch):
    ia, ib = win[ck]
    if ia > ib:
      continue
    sl = slice(ck * chunk, min(n, (ck + 1) * chunk))
    xs, ls = x[sl], lw[sl]
    lo_c, hi_c, lwmax = summ[ck, 0], summ[ck, 1], summ[ck, 2]
    for gb in range(ia, ib + 1, LPS * R):
      for gl in range(LPS):
        g0 = gb + gl * R
        run_lo = f32(gfirst + g0 * hd); run_hi = f32(run_lo + f32(R - 1) * h)
        rev = run_hi < lo_c
        gs = g0 + R - 1 if rev else g0
        hs = -h if rev else h
        gpp = f32(gfirst + gs * hd)
        dr = (lo_c - run_hi) if rev else max(f32(run_lo - hi_c), f32(0))
        top = f32(lwmax - dr * dr)
        Kf = f32(0) if top > -64 else f32(min(np.floor(100.0 - float(top)), 1900.0))
        dd = (gpp - xs).astype(f32)
        e0 = ex2((-dd * dd + ls).astype(f32) + Kf)
        q = ex2(np.minimum((dd * (f32(-2) * hs) + (-h * h)).astype(f32), f32(31)))
        q2 = (q * q).astype(f32); q3 = (q2 * q).astype(f32); q4 = (q2 * q2).astype(f32)
        p = e0.copy()
        acc = np.zeros(R, dtype=f32)
        with np.errstate(over="ignore", invalid="ignore"):
          for b in range(0, R, 4):
            if b:
              p = (p * q4).astype(f32)
            acc[b] += np.sum(p, dtype=f32)
            acc[b + 1] += np.sum((p * q).astype(f32), dtype=f32)
            acc[b + 2] += np.sum((p * q2).astype(f32), dtype=f32)
            acc[b + 3] += np.sum((p * q3).astype(f32), dtype=f32)
        for r in range(R):
          g = gs + (-r if rev else r)
          if g <= ib and 0 <= g < G:
            dens[g] += float(f32(acc[r] * cr[r])) * 2.0 ** (-float(Kf))
        pairs += xs.size * R
  return dens, dict(R=R, LPS=LPS, chunk=chunk, nchunks=nch, pairs=pairs, x=x, lw=lw, gp=gp, win=win, M=M)
