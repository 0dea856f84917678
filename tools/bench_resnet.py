"""ResNet-18 v2 `features` on 2048 frames @224: ms per forward with and without the stage-1 halo kernel, max |difference| of the
features between the two paths.  usage: python tools/bench_resnet.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tennis_b200 import ops  # noqa: E402
from tennis_b200 import synthetic as O  # noqa: E402

p = O.synthetic_params("resnet18_v2", seed=1234)
bb = ops.Backbone("resnet18_v2", O.flatten_params("resnet18_v2", p))
x = torch.randn(2048, 3, 224, 224, device="cuda")
outs = {}
for tag, env in (("halo", None), ("gather", "1")):
    if env:
        os.environ["TN_RESNET_NO_HALO"] = env
    else:
        os.environ.pop("TN_RESNET_NO_HALO", None)
    for _ in range(2):
        out = bb(x)
    out = out[0] if isinstance(out, tuple) else out
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        bb(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    outs[tag] = out.float().clone()
    print("%-7s %.2f ms  %.0f frames/s" % (tag, ms, 2048 / ms * 1e3), flush=True)
print("max |halo - gather| = %.4g on max |feature| %.4g" % ((outs["halo"] - outs["gather"]).abs().max().item(),
                                                            outs["gather"].abs().max().item()))
