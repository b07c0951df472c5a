"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv) for the LAST solve in it:
python tools/summarize_launches.py launches.csv [header comment ...]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
rows = []
for row in r:
    if len(row) <= vi:
        continue
    v = float(row[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(row[ui], 1.0)
    rows.append((row[ki], v))
starts = [i for i, (k, _) in enumerate(rows) if "first_touch" in k]
last = rows[starts[-1]:] if starts else rows
agg = collections.OrderedDict()
for k, v in last:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print("#", c)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:6d} launches {t / 1e3:10.3f} ms {100 * t / tot:5.1f}%  avg {t / n:9.1f} us")
print(f"total {tot / 1e3:.3f} ms over {len(last)} launches")
