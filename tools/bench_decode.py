"""Times one beam-search translate call at BASELINE configs[3] shapes (development aid; TN_GNMT_CELL_V1=1 selects the old cell kernel)."""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.argv = [sys.argv[0], "150"]
exec(open("tools/ncu_gnmt_decode_target.py").read().split("tr.translate(x.to(dev), vl.to(dev))")[0])
for _ in range(2):
    s, sc, v = tr.translate(x.to(dev), vl.to(dev))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    s, sc, v = tr.translate(x.to(dev), vl.to(dev))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
toks = int((v[:, 0].cpu() - 2).clamp(min=0).sum())
print("translate: %.2f ms per call, %d steps, %d best-beam tokens -> %.0f tokens/s" % (dt * 1e3, s.shape[2] - 1, toks, toks / dt))
