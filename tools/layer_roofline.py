"""Per-launch roofline of one DenseNet-121 forward from an ncu launch list (time + DRAM bytes per launch, the capture of
profiles/r1_launches_final2.csv): algorithmic conv FLOPs of each launch, the time it would take at the measured tensor peak
and at the measured HBM peak for the bytes it actually moved, and the time it took.
usage: python tools/layer_roofline.py launches.csv N_FRAMES [MEASURED_PEAKS.json]"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_frames = int(sys.argv[2])
peaks = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else {"bf16_tflops_sustained": 1394.6, "hbm_gbs": 6548.8}
PF, BW = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
hdr, per = None, {}
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        e = per.setdefault(int(d['ID']), {'k': re.sub(r'\(.*', '', d['Kernel Name']).split('::')[-1]})
        v, u = float(d['Metric Value'].replace(',', '')), d['Metric Unit']
        if d['Metric Name'] == 'gpu__time_duration.sum':
            e['s'] = v * {'ns': 1e-9, 'nsecond': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3}.get(u, 1.0)
        else:
            e['bytes'] = e.get('bytes', 0.0) + v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)

# the launch order of densenet_forward (tn_backbone.cu): algorithmic MACs per frame of every conv launch, in order
cfg, hw = (6, 12, 24, 16), (56 * 56, 28 * 28, 14 * 14, 7 * 7)
plan = [("stem 7x7/2 + pool", 112 * 112 * 64 * 147)]
c = 64
for b, nl in enumerate(cfg):
    for l in range(nl):
        plan.append(("block%d 1x1 K=%d" % (b + 1, c), hw[b] * 128 * c))
        plan.append(("block%d 3x3" % (b + 1), hw[b] * 32 * 1152))
        c += 32
    if b < 3:
        plan.append(("trans%d 1x1 K=%d" % (b + 1, c), hw[b] * (c // 2) * c))
        c //= 2
conv = [per[i] for i in sorted(per) if any(t in per[i]['k'] for t in ('conv_gemm_kernel', 'conv1x1_ts_kernel', 'conv3x3_halo_kernel', 'stem_pool_kernel'))]
assert len(conv) == len(plan), (len(conv), len(plan))
groups = {}
for (name, macs), e in zip(plan, conv):
    key = re.sub(r' K=\d+', '', name)
    g = groups.setdefault(key, [0, 0.0, 0.0, 0.0])
    g[0] += 1
    g[1] += 2.0 * macs * n_frames
    g[2] += e.get('bytes', 0.0)
    g[3] += e['s']
other = sum(per[i]['s'] for i in per) - sum(e['s'] for e in conv)
print("| layer group | launches | GFLOP | DRAM GB | t @ tensor peak (ms) | t @ HBM peak (ms) | actual (ms) | actual / max(bound) |")
print("|---|---|---|---|---|---|---|---|")
tt = tb = ta = tm = 0.0
for k, (n, fl, by, s) in groups.items():
    a, b = fl / PF * 1e3, by / BW * 1e3
    tt, tb, ta, tm = tt + a, tb + b, ta + s * 1e3, tm + max(a, b)
    print("| %s | %d | %.0f | %.2f | %.2f | %.2f | %.2f | %.2fx |" % (k, n, fl / 1e9, by / 1e9, a, b, s * 1e3, s * 1e3 / max(a, b)))
print("| **all conv launches** | %d | | | %.2f | %.2f | %.2f | %.2fx of sum of per-group max(bound) = %.2f ms |" % (len(conv), tt, tb, ta, ta / tm, tm))
print("\nnon-conv launches (input conversion, transition pre-pass, border zeroing, tail pool): %.2f ms" % (other * 1e3))
