#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of a bench.py step (raw page as CSV): DRAM bytes read + written per
launch of the four stage kernels, the capture it came from and the fingerprint of the kernel sources it was taken with
(bench.py quotes `roofline.traffic` only when the sources it runs still have that fingerprint).
usage: traffic_from_ncu.py raw.csv profiles/<summary>.json"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def gb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = "decode" if "k_decode" in name else "scan" if "k_scan" in name else "crc" if "k_crc" in name else "parse" if "k_parse" in name else None
    if not key:
        continue
    b = gb(r[ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_read.sum"]]) + gb(r[ix["dram__bytes_write.sum"]], rows[1][ix["dram__bytes_write.sum"]])
    out[key] = max(out.get(key, 0), int(b))          # (the serial k_parse launch of a clean stream is the larger one)
out["source"] = sys.argv[2]
out["kernels_fingerprint"] = bench.kernels_fingerprint()
out["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, one ncu --set full capture of `bench.py --steps 1 --warmup 1 --no-configs --no-e2e --no-cpu` (cfg2)"
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(out)
