"""Aggregate an `ncu --csv --metrics ...` log by (kernel, grid, block): launches, total time, mean pipe / issue utilisation, DRAM bytes."""
import csv, collections, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, im, iv, iid, ig, ib = (hdr.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Grid Size', 'Block Size'))
d = collections.OrderedDict()
for r in rows:
    if not r[iid].isdigit(): continue
    k = d.setdefault(r[iid], {'name': r[ik][:70], 'grid': r[ig], 'block': r[ib]})
    k[r[im]] = float(r[iv].replace(',', ''))
agg = collections.OrderedDict()
for k in d.values():
    key = (k['name'].split('(')[0][-38:], k['grid'], k['block'])
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += k['gpu__time_duration.sum']
    a[2] += k.get('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 0)
    a[3] += k.get('dram__bytes_read.sum', 0) + k.get('dram__bytes_write.sum', 0)
    a[4] += k.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)
tot = sum(a[1] for a in agg.values())
sc = 1e3 if tot > 1e6 else 1.0
print('total %.1f us' % (tot / sc))
for key, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-40s grid %-16s blk %-12s n=%3d t=%8.1f us fma=%4.1f%% issue=%4.1f%% dram=%6.1f MB/launch' % (key[0], key[1], key[2], a[0], a[1] / sc, a[2] / a[0], a[4] / a[0], a[3] / a[0] / 1e6))
