"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the top launches."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "nsecond": 1, "msecond": 1e6}.get(unit, 1)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    rows.append((int(row["ID"]), name, ns, row.get("Grid Size", ""), row.get("Block Size", "")))
tot = sum(r[2] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for _, n, ns, _, _ in rows:
    agg[n][0] += 1; agg[n][1] += ns
print(f"{len(rows)} launches, total {tot/1e6:.3f} ms")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns/1e6:9.3f} ms {100*ns/tot:5.1f}%  x{c:4d}  avg {ns/c/1e3:9.1f} us  {n[:110]}")
if len(sys.argv) > 2:
    print("--- top launches")
    for i, n, ns, g, b in sorted(rows, key=lambda r: -r[2])[: int(sys.argv[2])]:
        print(f"{ns/1e3:10.1f} us  id {i:5d} grid {g:>16s} block {b:>12s} {n[:90]}")
