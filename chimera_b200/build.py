"""Build libchimera_b200.so in-tree with nvcc for sm_100a (B200).  No GPU needed to compile."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("CHB_BUILD_OUT") or os.path.join(HERE, "libchimera_b200.so")
SOURCES = ["tables.cu", "selection.cu", "numerator.cu", "numerator_f32.cu", "numerator_fused.cu", "numerator_fused_nt128.cu", "numerator_fused_nt64.cu", "api.cu", "microbench.cu", "setup.cu"]
# setup.cu holds the HEALPix index arithmetic: no FMA contraction, so that it rounds like the host libraries
EXTRA_FLAGS = {"setup.cu": ["-fmad=false"]}
HEADERS = ["numerator_fused.cu", "models.cuh", "models_f32.cuh", "kde_f32.cuh", "kde_win.cuh", "common.cuh", "devguard.cuh", "stage.cuh", os.path.join("..", "..", "include", "chimera_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
  """Compile every CUDA source into chimera_b200/libchimera_b200.so; returns the path."""
  if not force and not _stale():
    return LIB
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + os.environ.get("CHB_BUILD_DEFS", "").split()
  objs = []
  bdir = os.path.join(HERE, "build" + ("_" + os.path.basename(LIB) if os.environ.get("CHB_BUILD_OUT") else ""))
  os.makedirs(bdir, exist_ok=True)
  procs = []
  for s in SOURCES:
    o = os.path.join(bdir, s.replace(".cu", ".o"))
    cmd = [nvcc] + flags + EXTRA_FLAGS.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
    procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(o)
  for s, p in procs:
    out, _ = p.communicate()
    if verbose or p.returncode != 0:
      sys.stderr.write(out)
    if p.returncode != 0:
      raise RuntimeError(f"nvcc failed on {s}")
  cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
  subprocess.check_call(cmd)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
