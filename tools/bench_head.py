"""Times the Bi-GRU head alone (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
from tennis_b200 import synthetic as O
from tennis_b200 import ops
B, T, D, H = 256, 32, 1024, 128
p = O.synthetic_rnn_params("gru", D, H, seed=4321)
rnn = ops.BiRNN("gru", D, H, p)
x = torch.randn(B, T, D, device="cuda").relu().to(torch.bfloat16)
for _ in range(3):
    rnn(x, want_y=False, want_max=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    rnn(x, want_y=False, want_max=True)
e1.record()
torch.cuda.synchronize()
print("head us/call", e0.elapsed_time(e1) * 1e3 / 20)
