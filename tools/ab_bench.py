"""A/B timing of the DenseNet-121 forward under environment switches (development aid).
usage: python tools/ab_bench.py NAME=VAL[,NAME=VAL...] [more variants...]   ('base' = no switches)"""
import os
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402  (seeded synthetic weights)
from tennis_b200 import ops  # noqa: E402

n = int(os.environ.get("AB_FRAMES", "2048"))
p = O.synthetic_params("densenet121", seed=1234)
bb = ops.Backbone("densenet121", O.flatten_params("densenet121", p))
x = torch.randn(n, 3, 224, 224, device="cuda")
variants = sys.argv[1:] or ["base"]
touched = set()
ref = None
for rep in range(2):
    for v in variants:
        for k in touched:
            os.environ.pop(k, None)
        if v != "base":
            for kv in v.split(","):
                k, val = kv.split("=")
                os.environ[k] = val
                touched.add(k)
        for _ in range(2):
            out = bb(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 5
        for _ in range(iters):
            out = bb(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if ref is None:
            ref = out.clone()
        d = (out - ref).abs().max().item()
        print("[%d] %-40s %.2f ms  %.0f frames/s   max|out-first|=%.4g (max|first|=%.3g)" % (rep, v, ms, n / ms * 1e3, d, ref.abs().max().item()), flush=True)
