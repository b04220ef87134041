#!/usr/bin/env python
"""Adds one capture to profiles/scan_traffic.json from an exported `ncu --page raw --csv` file (one kernel launch):
       python tools/ncu_traffic.py <raw.csv> <M> <rows> <queries> <k> <source note>
bench.py reports `roofline.traffic` = dram bytes of ONE launch of exactly the shape it timed, else null."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, M, rows, queries, k, note = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
r = list(csv.reader(open(path)))
hdr, units, vals = r[0], r[1], r[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def get(name):
    i = hdr.index(name)
    return float(vals[i].replace(",", "")) * scale[units[i]]


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
t = float(vals[hdr.index("gpu__time_duration.sum")].replace(",", ""))
tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
doc = {"captures": []}
if os.path.exists(tp):
    try:
        old = json.load(open(tp))
        if isinstance(old.get("captures"), list):
            doc = old
    except Exception:
        pass
doc["captures"] = [c for c in doc["captures"] if not (c["M"] == M and c["rows"] == rows and c["queries"] == queries and c["k"] == k)]
doc["captures"].append({"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "adc_scan_topk_kernel", "M": M, "rows": rows,
                        "queries": queries, "k": k, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                        "dram_bytes_per_launch": int(rd + wr), "gpu_time_duration": t,
                        "gpu_time_unit": units[hdr.index("gpu__time_duration.sum")], "source": note})
json.dump(doc, open(tp, "w"), indent=1)
print("traffic", M, rows, queries, k, int(rd + wr))
