"""Developer tool: per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections, csv, sys
for fn in sys.argv[1:]:
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][-40:]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg.setdefault(name, []).append(v)
    print(fn)
    for k, v in agg.items():
        print(f"  {k:42s} n={len(v):4d} total={sum(v)/1e3:9.3f} ms  mean={sum(v)/len(v):9.1f} us  min={min(v):.1f} max={max(v):.1f}")
