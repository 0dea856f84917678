"""Sweep the per-block frame-chunk sizes of the DenseNet schedule (TN_CHUNK_B1..4) on the GPU and print ms per 2048 frames.
Development aid: the chosen defaults go into tn_backbone.cu::chunk_frames."""
import os
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402  (seeded synthetic weights)
from tennis_b200 import ops  # noqa: E402

n = 2048
p = O.synthetic_params("densenet121", seed=1234)
bb = ops.Backbone("densenet121", O.flatten_params("densenet121", p))
x = torch.randn(n, 3, 224, 224, device="cuda")


def run(env, iters=4):
    for k in ("TN_CHUNK_B1", "TN_CHUNK_B2", "TN_CHUNK_B3", "TN_CHUNK_B4"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    out = None
    for _ in range(2):
        out = bb(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = bb(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


base, ref = run({})
print("baseline (no chunking): %.2f ms" % base, flush=True)
sweeps = {
    "TN_CHUNK_B1": [24, 30, 36, 48, 60, 72, 96, 145, 290],
    "TN_CHUNK_B2": [48, 72, 96, 121, 145, 193, 290, 580],
    "TN_CHUNK_B3": [97, 145, 193, 290, 386, 580, 773],
    "TN_CHUNK_B4": [193, 386, 773],
}
best = {}
for k, vals in sweeps.items():
    for v in vals:
        ms, out = run({k: v})
        same = bool(torch.equal(out, ref))
        print("%s=%d: %.2f ms (delta %+.2f) bit-identical=%s" % (k, v, ms, ms - base, same), flush=True)
        if k not in best or ms < best[k][1]:
            best[k] = (v, ms)
combo = {k: v for k, (v, ms) in best.items() if ms < base - 0.05}
ms, out = run(combo, iters=8)
print("combined %s: %.2f ms  -> %.0f frames/s; bit-identical=%s" % (combo, ms, n / ms * 1e3, bool(torch.equal(out, ref))), flush=True)
