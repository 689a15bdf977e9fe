"""Developer helper: fills the @PLACEHOLDERS@ of DESIGN.md section 4.1 from a bench.py line (python scripts/fill_design.py profiles/bench_r2_final.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
sm, ec = d["per_rule"]["split_merge"], d["per_rule"]["erase_create"]
names = {"num_child": "num_child", "pre_truncate": "group + sort of the work items", "symbolic": "symbolic", "insert": "dedup", "compact": "compact", "truncate": "truncate", "finalize": "finalize"}


def phases(r):
    return ", ".join(f"{names[k]} {v:.1f}" for k, v in sorted(r["phase_ms"].items(), key=lambda kv: -kv[1]) if k in names and v >= 0.15)


sub = {"@SM_MS@": f"{sm['ms_per_call']:.1f}", "@EC_MS@": f"{ec['ms_per_call']:.1f}", "@SM_PHASES@": phases(sm), "@EC_PHASES@": phases(ec),
       "@SM_FRAC@": f"{sm['whole_iteration']['frac']:.2f}", "@EC_FRAC@": f"{ec['whole_iteration']['frac']:.2f}", "@STEP_MS@": f"{d['ms_per_step']:.1f}",
       "@VALUE@": f"{d['value']:.2e}", "@STEP_FRAC@": f"{d['roofline']['whole_step']['frac']:.2f}", "@SM_SYM@": f"{sm['phase_ms']['symbolic']:.1f}"}
path = os.path.join(ROOT, "DESIGN.md")
text = open(path).read()
for k, v in sub.items():
    text = text.replace(k, v)
open(path, "w").write(text)
print(sub)
