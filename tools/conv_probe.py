"""Probe the TMA-fed 1x1 conv GEMM: time vs channel stride (DRAM access pattern), K, M (L2 residency), prologue on/off.
Development aid; prints achieved TB/s (activation read + output write) and TFLOP/s."""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import ops  # noqa: E402


def run(M, K, cs, Cout=128, pro=True, out_cs=128, iters=10, flush=None):
    g = torch.Generator().manual_seed(1)
    w = torch.randn(Cout, K, 1, 1, generator=g) * 0.05
    ps = (torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.1) if pro else (None, None)
    es = (torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1)
    conv = ops.Conv(w, pro_scale=ps[0], pro_shift=ps[1], epi_scale=es[0], epi_shift=es[1])
    assert M % 196 == 0
    x = torch.randn(M // 196, 14, 14, cs, device="cuda").to(torch.bfloat16)
    out = torch.empty(M // 196, 14, 14, out_cs, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        conv(x, epi_relu=True, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        conv(x, epi_relu=True, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = M * K * 2 + M * Cout * 2
    fl = 2.0 * M * K * Cout
    print("M=%7d K=%4d cs=%4d pro=%d flush=%d: %.3f ms  %.2f TB/s  %.0f TFLOP/s" % (
        M, K, cs, pro, flush is not None, ms, by / ms / 1e9, fl / ms / 1e9), flush=True)


flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
M3 = 1024 * 196
for K, cs in [(512, 1024), (512, 512), (992, 1024), (256, 1024), (256, 256)]:
    run(M3, K, cs, flush=flush)
    run(M3, K, cs, pro=False, flush=flush)
# L2-resident input (25 MB) without flushing
run(25088, 512, 512)
run(25088, 512, 1024)
run(25088, 512, 512, pro=False)
# block-1-like: small K, large M
M1 = 512 * 3136
for K, cs in [(64, 256), (64, 64), (224, 256)]:
    run(M1, K, cs, flush=flush)
    run(M1, K, cs, pro=False, flush=flush)
