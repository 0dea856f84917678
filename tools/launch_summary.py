"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
per-kernel launches, time, DRAM bytes and share of the step.  usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = {}
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        i = int(d['ID'])
        k = re.sub(r'\(.*', '', d['Kernel Name']).split('::')[-1]
        e = per.setdefault(i, {'k': k})
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        if d['Metric Name'] == 'gpu__time_duration.sum':
            e['ms'] = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
        else:
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            e[d['Metric Name']] = v * mult
agg = collections.OrderedDict()
for i in sorted(per):
    e = per[i]
    a = agg.setdefault(e['k'], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += e.get('ms', 0.0)
    a[2] += e.get('dram__bytes_read.sum', 0.0)
    a[3] += e.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
print('%-58s %5s %9s %6s %9s %9s %7s' % ('kernel', 'n', 'ms', 'share', 'rd GB', 'wr GB', 'TB/s'))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-58s %5d %9.3f %5.1f%% %9.2f %9.2f %7.2f' % (k[:58], a[0], a[1], 100 * a[1] / tot, a[2] / 1e9, a[3] / 1e9,
                                                      (a[2] + a[3]) / max(a[1], 1e-9) / 1e9))
print('%-58s %5d %9.3f %6s %9.2f %9.2f' % ('TOTAL', sum(a[0] for a in agg.values()), tot, '', sum(a[2] for a in agg.values()) / 1e9,
                                        sum(a[3] for a in agg.values()) / 1e9))
