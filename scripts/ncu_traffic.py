"""profiles/traffic.json from the ncu CSV of scripts/ncu_traffic.sh: DRAM bytes (read + write) and device time per launch of
the dominant kernels on the loop state, keyed by the rule whose iteration launches them (bench.py reads `dram_bytes_per_launch`
of the symbolic kernel as roofline.traffic when `parents` matches its own)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_traffic.csv")
parents = int(sys.argv[2]) if len(sys.argv) > 2 else 10**7
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    name, metric, unit, value = r[col["Kernel Name"]], r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6, "second": 1e3}.get(unit, 1)
    per.setdefault((r[col["ID"]], name), {})[metric] = value * scale
out = {}
for (_, name), m in per.items():
    rule = "erase_create" if "flip_rule" in name or "table_compact" in name else "split_merge"
    key = "symbolic" if "symbolic" in name else ("dedup" if "bin_dedup" in name else "compact")
    entry = {"kernel": name.split("(")[0], "dram_bytes_per_launch": m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
             "dram_read": m.get("dram__bytes_read.sum", 0), "dram_write": m.get("dram__bytes_write.sum", 0), "ncu_ms": m.get("gpu__time_duration.sum")}
    out.setdefault(rule, {"parents": parents})
    if key == "symbolic":
        out[rule].update(entry)
    else:
        out[rule][key] = entry
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
