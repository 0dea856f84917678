"""Generate tests/golden/captioning_oracle.npz from the CPU oracle (run here, commit the output): teacher-forced logits, the
masked-CE loss, and the beam-search token ids / valid lengths / scores of seeded GNMT models (LSTM and GRU).  Like
tools/make_golden.py these pin the ORACLE against silent drift; the reference ships no vectors of its own."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import captioning as C  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "captioning_oracle.npz")
CFG = dict(H=32, D=48, E=20, V=37, B=4, T=9, Tt=7, beam=4, max_length=12)


def compute(cell):
    c = CFG
    p = C.synthetic_gnmt_params(seed=10000, scale=0.3, cell=cell, H=c["H"], D_src=c["D"], E=c["E"], V=c["V"])
    x, vl = C.synthetic_sources(c["B"], c["T"], c["D"], seed=3)
    g = torch.Generator().manual_seed(5)
    tgt = torch.randint(0, c["V"], (c["B"], c["Tt"]), generator=g).float()
    tvl = torch.tensor([7., 6., 4., 2.])
    with torch.no_grad():
        logits = C.nmt_forward(p, x, tgt[:, :-1], vl, tvl - 1, cell=cell, H=c["H"])
        loss = C.masked_softmax_ce(logits, tgt[:, 1:], tvl - 1)
        samples, scores, vlen = C.translate(p, x, vl, cell=cell, H=c["H"], beam=c["beam"], max_length=c["max_length"])
    return {"logits": logits.numpy().astype(np.float32), "loss": loss.numpy().astype(np.float32),
            "samples": samples.numpy().astype(np.int32), "scores": scores.numpy().astype(np.float32),
            "valid_length": vlen.numpy().astype(np.int32)}


def main():
    torch.manual_seed(0)
    fix = {}
    for cell in ("lstm", "gru"):
        for k, v in compute(cell).items():
            fix["%s_%s" % (cell, k)] = v
    np.savez_compressed(OUT, **fix)
    for k, v in fix.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
