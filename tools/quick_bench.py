"""Quick device-side timing of the backbone forward (development aid, not the bench of record)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402  (seeded synthetic weights)
from tennis_b200 import ops  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "densenet121"
ns = [int(a) for a in sys.argv[2:]] or [256, 1024]
p = O.synthetic_params(arch, seed=1234)
bb = ops.Backbone(arch, O.flatten_params(arch, p))
for n in ns:
    x = torch.randn(n, 3, 224, 224, device="cuda")
    for _ in range(2):
        bb(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 3
    for _ in range(iters):
        bb(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%s n=%d: %.2f ms  -> %.0f frames/s  (%.1f%% of 240.8k roofline)" % (arch, n, ms, n / ms * 1e3, n / ms * 1e3 / 240.8e3 * 100), flush=True)
