"""Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration averaged per launch from
`ncu --page raw --csv` exports; writes the JSON bench.py reads for `roofline.traffic`.
Usage: python tools/ncu_traffic.py profiles/ncu_traffic.json <raw.csv> [<raw.csv> ...]"""
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
out = {}
for path in sys.argv[2:]:
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h, u = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(h)}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).split("::")[-1].strip()
        name = re.sub(r"<.*", "", name)
        def val(m):
            return float(r[ix[m]].replace(",", "")) * SCALE.get(u[ix[m]], 1.0)
        rec = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "dur_us": 0.0})
        rec["launches"] += 1
        rec["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        rec["dur_us"] += val("gpu__time_duration.sum")
res = {k: {"launches": v["launches"], "dram_bytes_per_launch": round(v["dram_bytes"] / v["launches"]),
           "dur_us_per_launch": round(v["dur_us"] / v["launches"], 2)} for k, v in sorted(out.items())}
res["_source"] = [p for p in sys.argv[2:]]
json.dump(res, open(sys.argv[1], "w"), indent=1)
print(json.dumps(res, indent=1))
