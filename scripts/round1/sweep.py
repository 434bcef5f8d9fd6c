"""Tuning sweep on the C3 workload: python scripts/sweep.py ENVVAR v1,v2,... [--fp64-ref]
Prints kernel times per setting and the deviation of the per-event log-likelihoods from the first setting
(and from the GPU fp64 mode with --fp64-ref)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

class A: pass
a = A()
a.nev, a.hyper_side, a.ns, a.ninj, a.nz = int(os.environ.get("SWEEP_NEV", 1000)), 16, 5000, 1_000_000, 300
var, vals = sys.argv[1], sys.argv[2].split(",")
w = bench.build_workload(a, 0)
ref64 = None
if "--fp64-ref" in sys.argv:
  like64 = bench.build_likelihood(w, "fp64", False)
  ref64 = like64.compute_all(**w["hyper"])[0]
  print("fp64 mode timings", like64.engine.timings(), flush=True)
  del like64
like = bench.build_likelihood(w, "fp32", False)

def dev(x, r):
  fin = np.isfinite(r) & (np.abs(r) < 1e300)
  e = np.abs(x[fin] - r[fin]) / np.maximum(np.abs(r[fin]), 1.0)
  cls = np.array_equal(np.isfinite(x) & (np.abs(x) < 1e300), fin)
  return f"max {e.max():.2e} p99.9 {np.quantile(e, 0.999):.2e} median {np.median(e):.2e} classes_equal {cls}"

ref = None
for v in vals:
  os.environ[var] = v
  for i in range(3):
    out = like.compute_all(**w["hyper"])
  t = like.engine.timings()
  lle = out[0]
  if ref is None:
    ref = lle
  msg = f"{var}={v}: numerator {t['numerator_ms']:.3f} ms  selection {t['selection_ms']:.3f} ms | vs first: {dev(lle, ref)}"
  if ref64 is not None:
    msg += f" | vs fp64: {dev(lle, ref64)}"
  print(msg, flush=True)
