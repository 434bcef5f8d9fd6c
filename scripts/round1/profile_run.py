"""Small fixed workload for ncu captures and the phase profile: python scripts/profile_run.py [nev] [side] [fp_mode]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

class A: pass
a = A()
a.nev = int(sys.argv[1]) if len(sys.argv) > 1 else 148
a.hyper_side = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fp = sys.argv[3] if len(sys.argv) > 3 else "fp32"
a.ns, a.ninj, a.nz = 5000, 200_000, 300
w = bench.build_workload(a, 0)
like = bench.build_likelihood(w, fp, False)
like.engine.phase_profile(True)
for i in range(3):
  t0 = time.perf_counter(); out = like(**w["hyper"]); dt = time.perf_counter() - t0
print("step s", dt, "timings", like.engine.timings())
prof = like.engine.phase_profile(True)
tot = sum(prof.values())
print("phase cycles per CTA:", {k: (round(v), f"{100*v/tot:.1f}%") for k, v in prof.items()})
