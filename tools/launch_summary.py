"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total, average, share.
usage: launch_summary.py launches.csv ['command line that produced it']"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv or not r[0].isdigit():
        continue
    a = agg.setdefault(r[ik][:64], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("(cold-cache serialised per-launch times from ncu: compare SHARES, not absolutes)  total %.1f us" % tot)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s n=%4d total=%10.1f us avg=%8.1f us share=%5.1f%%" % (k, n, t, t / n, 100 * t / tot))
