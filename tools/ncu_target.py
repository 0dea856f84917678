"""One warm-up forward, then ONE DenseNet-121 forward between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
usage: python tools/ncu_target.py [n_frames] [arch]"""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402  (seeded synthetic weights)
from tennis_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
arch = sys.argv[2] if len(sys.argv) > 2 else "densenet121"
p = O.synthetic_params(arch, seed=1234)
bb = ops.Backbone(arch, O.flatten_params(arch, p))
x = torch.randn(n, 3, 224, 224, device="cuda")
bb(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
bb(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
