#!/usr/bin/env python
"""Warp-stall samples of an .ncu-rep aggregated per CUDA source line (nvdisasm -g line table joined with the SASS page).
usage: python scripts/ncu_lines.py rep.ncu-rep lib.so <mangled kernel name> [main_file_basename] [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
main = sys.argv[4] if len(sys.argv) > 4 else None
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}      # offset -> (file, line, main_line)
for f in os.listdir(tmp):
  if not f.endswith(".cubin"):
    continue
  out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
  if f"\n{kern}:" not in out:
    continue
  inside, cur, cur_main = False, ("?", 0), 0
  for ln in out.splitlines():
    if ln.startswith(kern + ":"):
      inside = True; continue
    if inside and re.match(r"^[_A-Za-z][\w$.]*:$", ln) and not ln.startswith(".") and not ln.startswith(kern):
      if not ln.startswith(".L_") and not ln.startswith(".text"):
        break
    if not inside:
      continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
      cur = (os.path.basename(m.group(1)), int(m.group(2)))
      if main and cur[0] == main:
        cur_main = cur[1]
      continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
      line_of[int(m.group(1), 16)] = (cur[0], cur[1], cur_main, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels in one report: keep the block whose "Kernel Name" row matches NCU_KERNEL_INDEX (default 0)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if starts:
  ki = int(os.environ.get("NCU_KERNEL_INDEX", "0"))
  lo = starts[ki]
  hi_ = starts[ki + 1] if ki + 1 < len(starts) else len(rows)
  rows = rows[lo:hi_]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
per_main = collections.defaultdict(lambda: [0, 0])
tot = 0
for r in rows[hi + 1:]:
  if len(r) <= isamp or not r[ia].startswith("0x"):
    continue
  a = int(r[ia], 16)
  if base is None:
    base = a
  info = line_of.get(a - base, ("?", 0, 0, ""))
  s, ex = int(r[isamp] or 0), int(r[iex] or 0)
  tot += s
  e = per_line[(info[0], info[1])]
  e[0] += s; e[1] += ex
  for c in stall_cols:
    v = int(r[c] or 0)
    if v:
      e[2][hdr[c]] += v
  m = per_main[info[2]]
  m[0] += s; m[1] += ex
print(f"total samples {tot}")
print("\n== by innermost source line (top) ==")
for (f, l), (s, ex, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:topn]:
  top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(4))
  print(f"{100*s/tot:5.1f}%  {s:7d} smp  {ex:10d} inst  {f}:{l}   [{top}]")
if main:
  print(f"\n== by enclosing line of {main} ==")
  for l, (s, ex) in sorted(per_main.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*s/tot:5.1f}%  {s:7d} smp  {ex:10d} inst  {main}:{l}")

# optional region totals: NCU_REGIONS="name:lo-hi,name:lo-hi" over the enclosing lines of the main file
reg = os.environ.get("NCU_REGIONS")
if main and reg:
  print(f"\n== regions of {main} (enclosing lines) ==")
  texe = sum(ex for _, ex in per_main.values())
  for item in reg.split(","):
    nm, rng = item.split(":")
    lo_, hi_ = (int(x) for x in rng.split("-"))
    s_ = sum(v[0] for l, v in per_main.items() if lo_ <= l <= hi_)
    e_ = sum(v[1] for l, v in per_main.items() if lo_ <= l <= hi_)
    print(f"{nm:>16}: {100*s_/tot:5.1f}% samples  {100*e_/texe:5.1f}% instructions  ({e_} warp inst)")
