"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the listed launches)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
grid = collections.defaultdict(set)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("q3::", "").replace("(anonymous namespace)::", "")
    agg[k][0] += 1
    agg[k][1] += v
    grid[k].add(row.get("Grid Size", ""))
tot = sum(v[1] for v in agg.values())
print(f"{'total':>10}  {tot/1e6:9.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={n:6d}  avg={t/n/1e3:8.2f} us  {k[:70]}  grids={sorted(grid[k])[:4]}")
