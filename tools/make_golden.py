"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle (run here, commit the output).

The reference ships no golden vectors and its runtime (MXNet/GluonCV/GluonNLP) cannot be imported offline, so these
fixtures pin the ORACLE (so it cannot drift silently) rather than the reference itself; the oracle in turn is pinned
against torchvision's DenseNet-121 graph and torch.nn.GRU/LSTM in tests/test_oracle_cpu.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vision as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fix = {}
    with torch.no_grad():
        # config 1 flavour: one 224x224 frame through each backbone (seeded weights 1234, pixels seed 100)
        u8, x = O.synthetic_frames(2, 224, seed=100)
        for arch in ("densenet121", "resnet18_v2"):
            p = O.synthetic_params(arch, seed=1234)
            f = O.FEATURES[arch](x, p)
            fix[arch + "_feats"] = f.numpy().astype(np.float32)
        # CNN+GRU head on features (published 0042 layout): B=3, T=5, D=64, H=128
        g = torch.Generator().manual_seed(3)
        feats = torch.randn(3, 5, 64, generator=g).relu()
        for cell in ("gru", "lstm"):
            rp = O.synthetic_rnn_params(cell, 64, 128, seed=4321)
            fix["birnn_%s_y" % cell] = O.birnn_layer(feats, rp, cell, 128).numpy().astype(np.float32)
        gg = torch.Generator().manual_seed(77)
        cw = (torch.rand(11, 256, generator=gg) * 2 - 1) * 0.07
        cb = torch.zeros(11)
        rp = O.synthetic_rnn_params("gru", 64, 128, seed=4321)
        fix["cnnrnn_feats_logits"] = O.cnnrnn(feats, None, rp, "gru", 128, cw, cb, feats=True).numpy().astype(np.float32)
        fix["temporal_pool_mean"] = O.temporal_pooling(feats, None, "mean", feats=True).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "vision_oracle.npz"), **fix)
    for k, v in fix.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
