"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py list.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
by = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r[ix["Kernel Name"]].split("(")[0][:64]
    n, s = by.get(name, (0, 0.0))
    by[name] = (n + 1, s + us)
total = sum(s for _, s in by.values())
print(f"# {sys.argv[1]}: {sum(n for n, _ in by.values())} launches, {total:.1f} us of kernel time (serialised, cold cache)")
print(f"{'kernel':64s} {'n':>5s} {'sum_us':>10s} {'avg_us':>9s} {'share':>6s}")
for name, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:64s} {n:5d} {s:10.1f} {s / n:9.2f} {s / total:6.3f}")
