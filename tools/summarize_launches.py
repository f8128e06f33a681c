"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |dcb::\(anonymous namespace\)::|dcb::", "", name)
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, u in rows:
    agg[n][0] += 1; agg[n][1] += u
print("%d launches, %.2f ms total (ncu per-launch times are cold-cache and serialised: compare SHARES)" % (len(rows), tot / 1e3))
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  %5d x %9.1f us  %s" % (100 * u / tot, c, u / c, n[:110]))
