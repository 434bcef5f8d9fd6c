#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full capture of the numerator launch of `bench.py` (DRAM bytes per launch).
usage: python scripts/ncu_traffic.py rep.ncu-rep nev ns nz hyper_side fp_mode"""
import csv, io, json, subprocess, sys, os
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
  i = hdr.index(name)
  v = float(vals[i]); u = units[i]
  return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
d = dict(kernel=vals[hdr.index("Kernel Name")], dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr,
         nev=int(sys.argv[2]), ns=int(sys.argv[3]), nz=int(sys.argv[4]), hyper_side=int(sys.argv[5]), fp_mode=sys.argv[6],
         source=os.path.basename(rep), duration_ms_under_ncu=float(vals[hdr.index("gpu__time_duration.sum")]))
json.dump(d, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(d)
