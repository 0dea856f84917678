"""Bi-GRU head forward (projection GEMM + scan, (256,32,1024) bf16 features) between cudaProfilerStart/Stop, for
`ncu --set full --profile-from-start off -k regex:rnn_scan|conv_gemm`.  usage: python tools/ncu_head_target.py"""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402
from tennis_b200 import ops  # noqa: E402

B, T, D, H = 256, 32, 1024, 128
p = O.synthetic_rnn_params("gru", D, H, seed=4321)
rnn = ops.BiRNN("gru", D, H, p)
x = torch.randn(B, T, D, device="cuda").relu().to(torch.bfloat16)
for _ in range(2):
    rnn(x, want_y=False, want_max=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rnn(x, want_y=False, want_max=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
