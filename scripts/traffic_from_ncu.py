"""profiles/traffic.json entry from an `ncu --set full` capture of the two X passes:
    python scripts/traffic_from_ncu.py REP.ncu-rep KEY "note"
KEY = "<workload>/<dtype>/<world>" as bench.py looks it up (dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, key, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ent = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = "h_pass" if "h_pass" in r[ik] else "w_pass" if "w_pass" in r[ik] else None
    if name:
        ent[name] = int(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
ent["capture"] = os.path.basename(rep) + ((" -- " + note) if note else "")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
t = json.load(open(path))
t[key] = ent
json.dump(t, open(path, "w"), indent=1)
print(key, ent)
