"""DenseNet-121 `features` in the split-bf16 mode on 2048 frames @224: ms per forward.  TN_PRECISE_CHUNK is read once per process.
usage: TN_PRECISE_CHUNK=256 python tools/bench_precise.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tennis_b200 import ops  # noqa: E402
from tennis_b200 import synthetic as O  # noqa: E402

p = O.synthetic_params("densenet121", seed=1234)
bb = ops.Backbone("densenet121", O.flatten_params("densenet121", p), precision="split_bf16")
x = torch.randn(2048, 3, 224, 224, device="cuda")
bb(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = bb(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
out = out[0] if isinstance(out, tuple) else out
print("chunk %s: %.1f ms  %.0f frames/s  checksum %.6f" % (os.environ.get("TN_PRECISE_CHUNK", "128"), ms, 2048 / ms * 1e3,
                                                         float(out.double().sum())), flush=True)
