"""ncu launch list (gpu__time_duration.sum CSV) → per-kernel summary (markdown).
usage: python scripts/summarize_launches.py gpurun_out/launches_updown.csv > profiles/r01_launches_updown.md"""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    key = (r["Kernel Name"].split("(")[0][:70], r["Grid Size"], r["Block Size"])
    agg.setdefault(key, []).append(float(r["Metric Value"].replace(",", "")))
total = sum(sum(v) for v in agg.values())
print(f"# ncu launch list summary: {path}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES)\n")
print("| kernel | grid | block | launches | avg µs | total µs | share |")
print("|---|---|---|---:|---:|---:|---:|")
for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/1e3:.1f} | {100*sum(v)/total:.1f}% |")
print(f"\ntotal {total/1e3:.1f} µs over {len(rows)} launches")
