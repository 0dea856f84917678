"""One captioner training step (bench.py::bench_training shapes) between cudaProfilerStart/Stop, for an ncu launch list."""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as C  # noqa: E402
from tennis_b200 import autograd  # noqa: E402
from tennis_b200.gluon import Dropout, Embedding, HybridSequential, MaskedSoftmaxCELoss, Trainer  # noqa: E402
from tennis_b200.models.captioning.gnmt import NMTModel, get_gnmt_encoder_decoder  # noqa: E402
from tennis_b200.vocab import Vocab, count_tokens  # noqa: E402

dev = torch.device("cuda", 0)
Bc, Ts, Dc, Hc, E, V, Tt = 128, 224, 1024, 128, 100, 254, 30
p = C.synthetic_gnmt_params(seed=10000, scale=0.1, cell="lstm", H=Hc, D_src=Dc, E=E, V=V)
vocab = Vocab(count_tokens(["w%03d" % i for i in range(V - 4)]))
src_embed = HybridSequential()
src_embed.add(Dropout(0.0))
enc, dec = get_gnmt_encoder_decoder(cell_type="lstm", hidden_size=Hc, dropout=0.2, num_layers=2, num_bi_layers=1)
model = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=enc, decoder=dec, embed_size=E, prefix="gnmt_", src_embed=src_embed,
                 tgt_embed=Embedding(V, E))
params = model.collect_params()
for k, v in p.items():
    params[k].shape, params[k]._data = tuple(v.shape), v.clone().to(dev)
    params[k]._version += 1
g = torch.Generator().manual_seed(11)
x, vl = C.synthetic_sources(Bc, Ts, Dc, seed=100, min_len=64)
tgt = torch.randint(4, V, (Bc, Tt), generator=g).float()
tvl = torch.randint(6, Tt + 1, (Bc,), generator=g).float()
xs, vls, ts, tvls = x.to(dev), vl.to(dev), tgt.to(dev), tvl.to(dev)
scale = float((Tt - 1) / (tvl - 1).mean())
mce = MaskedSoftmaxCELoss()
trc = Trainer(model.collect_params(), 'adam', {'learning_rate': 1e-3})


def step():
    with autograd.record():
        o, _ = model(xs, ts[:, :-1], vls, tvls - 1)
        lv = mce(o, ts[:, 1:], tvls - 1)
    autograd.backward([lv], [torch.full_like(lv, scale / Bc)])
    trc.step(1)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
