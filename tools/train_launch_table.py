"""Per-kernel table of ONE training step from an ncu csv launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]):
launches, time, DRAM bytes and achieved DRAM bandwidth.  usage: train_launch_table.py CSV [--list PATTERN]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
pat = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--list" else None
lines = [l for l in open(path) if l.startswith('"')]
rec = collections.OrderedDict()
for x in csv.DictReader(lines):
    r = rec.setdefault(int(x["ID"]), {"k": x["Kernel Name"], "g": x["Grid Size"]})
    r[x["Metric Name"]] = float(x["Metric Value"].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1e3, "ms": 1e6, "ns": 1, "msecond": 1e6, "usecond": 1e3,
                                                                          "nsecond": 1}.get(x["Metric Unit"], 1)
L = list(rec.values())
ends = [i for i, r in enumerate(L) if "adam_kernel" in r["k"]]
seg = L[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else L
short = lambda k: re.sub(r"\(.*", "", k).replace("void ", "").replace("hft::", "").replace("<unnamed>::", "").replace("tc::", "")
if pat:
    for r in seg:
        if re.search(pat, r["k"]):
            t = r["gpu__time_duration.sum"] / 1e3
            b = r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
            print("%8.1f us %-14s %7.1f MB %6.0f GB/s  %s" % (t, r["g"], b / 1e6, b / t / 1e3, short(r["k"])))
    sys.exit(0)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in seg:
    a = agg[short(r["k"])]
    a[0] += 1
    a[1] += r["gpu__time_duration.sum"] / 1e3
    a[2] += r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print("one step: %d launches, %.0f us summed kernel time, %.2f GB of DRAM traffic" % (len(seg), tot, sum(a[2] for a in agg.values()) / 1e9))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%% %4d launches %8.1f MB %6.0f GB/s  %s" % (a[1], 100 * a[1] / tot, a[0], a[2] / 1e6, a[2] / max(a[1], 1e-9) / 1e3, k))
