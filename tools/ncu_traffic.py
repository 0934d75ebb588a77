#!/usr/bin/env python3
"""Write profiles/residual_traffic.json from an `ncu --set full --page raw --csv` export of the default residual kernel:

    ncu --set full --clock-control none -k regex:k_residual_fast_bulk -c 1 -o gpurun_out/res_full python tools/res_one.py 8192x2048 0 2
    ncu -i gpurun_out/res_full.ncu-rep --page raw --csv > profiles/rX_residual_fast_full_raw.csv
    python tools/ncu_traffic.py profiles/rX_residual_fast_full_raw.csv 8192 2048

bench.py reads the sidecar and reports roofline.traffic only when the grid and the hash of the kernel sources match the code
that is running (bench.kernel_source_hash)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

path, im, jm = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr, units = rows[0], rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
pick = [r for r in data if "k_residual_fast_bulk" in r[hdr.index("Kernel Name")]] or \
       [r for r in data if "k_residual_fast" in r[hdr.index("Kernel Name")] and "tma" not in r[hdr.index("Kernel Name")]]
r = pick[-1]


def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)
    return v * mult


out = {"kernel": r[hdr.index("Kernel Name")], "grid": [im, jm], "dram_bytes_read": val("dram__bytes_read.sum"),
       "dram_bytes_write": val("dram__bytes_write.sum"), "source_hash": bench.kernel_source_hash(), "capture": os.path.relpath(path, ROOT)}
json.dump(out, open(os.path.join(ROOT, "profiles", "residual_traffic.json"), "w"), indent=1)
print(out)
