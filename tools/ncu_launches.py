"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    t = t / 1e3 if u in ("ns", "nsecond") else (t * 1e3 if u in ("ms", "msecond") else t)
    agg.setdefault((row["Kernel Name"][:58], row.get("Grid Size", "")), []).append(t)
tot = sum(sum(v) for v in agg.values())
print("%-60s %-16s %4s %10s %11s %6s" % ("kernel", "grid", "n", "avg us", "sum us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-60s %-16s %4d %10.1f %11.1f %5.1f%%" % (k[0], k[1], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
print("total us %.1f" % tot)
