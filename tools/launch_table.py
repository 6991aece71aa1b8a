"""Per-kernel table of the LAST MSM in an ncu launch list (csv written with --metrics gpu__time_duration.sum,dram__bytes_*,
sm__pipe_fmaheavy_cycles_active...): launches, total us, DRAM bytes, time-weighted fmaheavy %.
usage: python tools/launch_table.py gpurun_out/launches.csv [--seq]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ix['ID']], {'name': r[ix['Kernel Name']], 'grid': r[ix['Grid Size']]})
    d[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
ids = list(per)
last = [i for i in ids if 'k_recode' in per[i]['name']][-1]
sel = ids[ids.index(last):]
def short(n):
    n = n.replace('void ', '').replace('zk::', '')
    return n.split('(')[0][:44]
if '--seq' in sys.argv:
    for i in sel:
        d = per[i]
        print(i, short(d['name']), d['grid'], '%.1f us' % (d['gpu__time_duration.sum'] / 1e3), '%.0f MB' % ((d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6),
              '%.0f%%' % d.get('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 0))
agg = collections.OrderedDict()
for i in sel:
    d = per[i]; a = agg.setdefault(short(d['name']), [0, 0.0, 0.0, 0.0])
    t = d['gpu__time_duration.sum'] / 1e3
    a[0] += 1; a[1] += t; a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    a[3] += t * d.get('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 0)
tot = sum(a[1] for a in agg.values())
print('%-46s %4s %9s %9s %6s' % ('kernel', 'n', 'us', 'MB', 'fma%'))
for k, a in agg.items():
    print('%-46s %4d %9.1f %9.0f %6.1f' % (k, a[0], a[1], a[2] / 1e6, a[3] / a[1] if a[1] else 0))
print('%-46s %4s %9.1f %9.0f' % ('total', '', tot, sum(a[2] for a in agg.values()) / 1e6))
