"""Extract the physical constants the synthetic EDXS generator needs from the reference's tables
(espm/tables/200keV_xrays.json: x-ray line energies [keV] and emission cross sections at 200 keV;
espm/tables/SDD_efficiency.txt: detector efficiency curve) into espm_b200/data/edxs_tables.json, so that
espm_b200.synth can follow SURVEY.md section 8d on a box without the reference tree.  Data only, no code.

    python scripts/gen_xray_table.py [/root/reference]
"""
import json
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
tab = json.load(open(os.path.join(ref, "espm", "tables", "200keV_xrays.json")))["table"]
eff = np.loadtxt(os.path.join(ref, "espm", "tables", "SDD_efficiency.txt"))
# 25 elements that occur in STEM-EDXS samples of minerals / alloys / oxides (the first 9 are the C1 / C2 set)
Z = [8, 12, 13, 14, 20, 22, 26, 28, 29, 11, 15, 16, 19, 24, 25, 27, 30, 38, 40, 47, 50, 56, 57, 73, 79]
out = {"source": "adriente/espm v1.1.3 espm/tables/200keV_xrays.json + SDD_efficiency.txt (physical constants)",
       "elements": Z, "lines": {}, "sdd_efficiency": {"energy_keV": [round(float(v), 6) for v in eff[:, 0]],
                                                      "efficiency": [round(float(v), 6) for v in eff[:, 1]]}}
for z in Z:
    lines = tab[str(z)]
    out["lines"][str(z)] = [[name, float(v["energy"]), float(v["cs"])] for name, v in sorted(lines.items())
                            if 0.2 <= float(v["energy"]) <= 41.0 and float(v["cs"]) > 0.0]
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "espm_b200", "data", "edxs_tables.json")
json.dump(out, open(dst, "w"), separators=(",", ":"))
print(dst, os.path.getsize(dst), "bytes;", sum(len(v) for v in out["lines"].values()), "lines")
