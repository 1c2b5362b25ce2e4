"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python scripts/launch_shares.py gpurun_out/launches.csv [out.json]"""
import csv
import json
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, float(r["Metric Value"].replace(",", "")) * 1e-6))
agg = defaultdict(lambda: [0, 0.0])
for n, ms in rows:
    agg[n][0] += 1
    agg[n][1] += ms
tot = sum(v[1] for v in agg.values())
out = [dict(kernel=k, launches=v[0], ms=round(v[1], 3), share_pct=round(100 * v[1] / tot, 3))
       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
print(f"{len(rows)} launches, {tot:.1f} ms")
for o in out[:12]:
    print(o)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
