"""One DenseNet-121 forward in the split-bf16 (fp32-grade) mode between cudaProfilerStart/Stop.
usage: python tools/ncu_precise_target.py [n_frames]"""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402
from tennis_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = O.synthetic_params("densenet121", seed=1234)
bb = ops.Backbone("densenet121", O.flatten_params("densenet121", p), precision="split_bf16")
x = torch.randn(n, 3, 224, 224, device="cuda")
bb(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
bb(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
