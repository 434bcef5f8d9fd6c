"""fp32-mode vs fp64-mode per-event log-likelihoods on the C3 workload: list the worst units."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

class A: pass
a = A()
a.nev, a.hyper_side, a.ns, a.ninj, a.nz = int(os.environ.get("SWEEP_NEV", 1000)), 16, 5000, 100_000, 300
w = bench.build_workload(a, 0)
like64 = bench.build_likelihood(w, "fp64", False)
r64 = like64.compute_all(**w["hyper"])[0]
raw64 = like64.engine.last_numlike_evs() if hasattr(like64.engine, "last_numlike_evs") else None
del like64
out = {}
for tag, env in (("win", os.environ.get("CHB_KDE_WIN", "16")), ("nowin", "0")):
  os.environ["CHB_KDE_WIN"] = env
  like = bench.build_likelihood(w, "fp32", False)
  r32 = like.compute_all(**w["hyper"])[0]
  fin = np.isfinite(r64) & (np.abs(r64) < 1e300)
  err = np.where(fin, np.abs(r32 - r64) / np.abs(r64), 0.0)
  idx = np.argsort(err.ravel())[::-1][:30]
  rows = []
  for i in idx:
    h, e = np.unravel_index(i, err.shape)
    rows.append(dict(h=int(h), ev=int(e), err=float(err[h, e]), l32=float(r32[h, e]), l64=float(r64[h, e])))
  out[tag] = rows
  print(tag, "n(err>1e-3) =", int((err > 1e-3).sum()), "n(err>1e-4) =", int((err > 1e-4).sum()), "of", err.size)
  for r in rows[:12]:
    print(r)
  del like
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/worst_units.json", "w"))
