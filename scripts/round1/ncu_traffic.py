#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full capture of the numerator launches of ONE `bench.py` step
(reweighting + KDE/z-integral kernels of the split fast path): DRAM bytes read + written, summed over the launches.
usage: python scripts/ncu_traffic.py rep.ncu-rep nev ns nz hyper_side fp_mode"""
import csv, io, json, subprocess, sys, os
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def get(vals, name):
  i = hdr.index(name)
  return float(vals[i]) * scale[units[i]]
kernels = []
for vals in rows[2:]:
  kernels.append(dict(kernel=vals[hdr.index("Kernel Name")], dram_bytes_read=get(vals, "dram__bytes_read.sum"),
                      dram_bytes_write=get(vals, "dram__bytes_write.sum"),
                      duration_ms_under_ncu=float(vals[hdr.index("gpu__time_duration.sum")]) *
                      {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]))
d = dict(kernels=kernels, dram_bytes_per_launch=sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in kernels),
         nev=int(sys.argv[2]), ns=int(sys.argv[3]), nz=int(sys.argv[4]), hyper_side=int(sys.argv[5]), fp_mode=sys.argv[6],
         source=os.path.basename(rep),
         note="one step = one reweighting launch + one KDE/z-integral launch (split fast path); the 2 x 10.2 GB are the "
              "stage buffer {z, w} written by the first and read by the second")
json.dump(d, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
