"""Per-kernel DRAM bytes and time of ONE 16-segment forward out of an ncu csv
(--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum): launches from the 2nd front kernel to the 3rd.
usage: dram_summary.py dram.csv 'command line' [hbm_peak_GBs] [segments_per_forward=16]"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, im, iv, iid = (hdr.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID'))
d = collections.OrderedDict()
for r in rows:
    if not r[iid].isdigit(): continue
    d.setdefault(int(r[iid]), {'name': r[ik]})[r[im]] = float(r[iv].replace(',', ''))
ids = sorted(d)
fronts = [i for i in ids if 'front' in d[i]['name'] and 'collapse' not in d[i]['name']]
a, b = fronts[1], fronts[2]
sel = [d[i] for i in ids if a <= i < b]
R = sum(k['dram__bytes_read.sum'] for k in sel); W = sum(k['dram__bytes_write.sum'] for k in sel); T = sum(k['gpu__time_duration.sum'] for k in sel) / 1e3
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6536.7
nseg = int(sys.argv[4]) if len(sys.argv) > 4 else 16
print(sys.argv[2])
print('one %d-segment forward = launches %d..%d (%d kernels): DRAM read %.2f GB + write %.2f GB = %.2f GB (%.3f GB / segment), %.2f ms serialised -> %.2f TB/s = %.0f%% of the measured HBM peak %.1f GB/s'
      % (nseg, a, b - 1, len(sel), R / 1e9, W / 1e9, (R + W) / 1e9, (R + W) / (nseg * 1e9), T / 1e3, (R + W) / T / 1e6, 100 * (R + W) / T / 1e3 / peak, peak))
agg = collections.OrderedDict()
for k in sel:
    x = agg.setdefault(k['name'][:56], [0, 0, 0, 0]); x[0] += 1; x[1] += k['dram__bytes_read.sum']; x[2] += k['dram__bytes_write.sum']; x[3] += k['gpu__time_duration.sum'] / 1e3
for n, x in sorted(agg.items(), key=lambda y: -y[1][3]):
    print('%-58s n=%3d read %7.2f GB write %7.2f GB time %8.1f us (%4.1f%%) -> %5.2f TB/s' % (n, x[0], x[1] / 1e9, x[2] / 1e9, x[3], 100 * x[3] / T, (x[1] + x[2]) / x[3] / 1e6))
