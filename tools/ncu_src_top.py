"""Print the hottest SASS instructions (warp-stall samples) of an `ncu --page source --csv` export, with a window of context.
usage: python tools/ncu_src_top.py export.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[h]
iS = hdr.index('Warp Stall Sampling (All Samples)')
iSrc = hdr.index('Source')
iX = hdr.index('Instructions Executed')
iW = hdr.index('L1 Wavefronts Shared')
iWi = hdr.index('L1 Wavefronts Shared Ideal')
data = []
for k, r in enumerate(rows[h + 1:]):
    if len(r) != len(hdr):
        continue
    try:
        data.append((int(r[iS] or 0), k, r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
top = sorted(data, key=lambda x: -x[0])[:topn]
for s, k, r in sorted(top, key=lambda x: x[1]):
    print('%5d %5d (%4.1f%%) %-80s exec %s wf %s/%s' % (k, s, 100.0 * s / max(tot, 1), r[iSrc].strip()[:80], r[iX], r[iW], r[iWi]))
