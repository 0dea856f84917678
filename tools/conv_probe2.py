"""Does L2 residency pay for the 1x1 conv -> 3x3 conv pair?  Times back-to-back launches of (1x1 into the padded bottleneck,
3x3 halo conv out of it) on a frame chunk small enough to stay in L2 vs. a DRAM-sized one.  Development aid."""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import ops  # noqa: E402
from tennis_b200 import _lib  # noqa: E402


def run(frames, H, K, cs, reps):
    g = torch.Generator().manual_seed(1)
    w = torch.randn(128, K, 1, 1, generator=g) * 0.05
    ps = (torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.1)
    es = (torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1)
    conv = ops.Conv(w, pro_scale=ps[0], pro_shift=ps[1], epi_scale=es[0], epi_shift=es[1])
    x = torch.randn(frames, H, H, cs, device="cuda").to(torch.bfloat16)
    out = torch.empty(frames, H, H, 128, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        conv(x, epi_relu=True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        conv(x, epi_relu=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = frames * H * H * (K + 128) * 2
    print("1x1 frames=%5d HxW=%dx%d K=%3d cs=%4d: %.1f us/launch  %.2f TB/s  (%.2f us per frame), footprint %.0f MB" % (
        frames, H, H, K, cs, ms * 1e3, by / ms / 1e9, ms * 1e3 / frames, (x.numel() + out.numel()) * 2 / 1e6), flush=True)


for K in (64, 224):
    for frames in (24, 30, 48, 96, 512):
        run(frames, 56, K, 256, reps=40 if frames < 200 else 6)
for K in (256, 512, 992):
    for frames in (96, 193, 1024):
        run(frames, 14, K, 1024, reps=40 if frames < 500 else 6)
