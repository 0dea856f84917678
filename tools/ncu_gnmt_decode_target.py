"""One beam-search translate call at BASELINE configs[3] shapes (B=32, T_src<=224, 1024-d, LSTM H=128, beam 5, 30 steps) between
cudaProfilerStart/Stop, for an ncu launch list / `--set full -k regex:cell_kernel|attn_kernel|proj_kernel|beam_kernel`."""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as S  # noqa: E402
from tennis_b200.gluon import Dropout, Embedding, HybridSequential  # noqa: E402
from tennis_b200.models.captioning.gnmt import BeamSearchScorer, NMTModel, get_gnmt_encoder_decoder  # noqa: E402
from tennis_b200.utils.translation import BeamSearchTranslator  # noqa: E402
from tennis_b200.vocab import Vocab, count_tokens  # noqa: E402

dev = torch.device("cuda", 0)
B, Tsrc, D, H, E, V, beam, max_len = 32, 224, 1024, 128, 100, 254, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 30
p = S.synthetic_gnmt_params(seed=10000, scale=0.35, cell="lstm", H=H, D_src=D, E=E, V=V)
vocab = Vocab(count_tokens(["w%03d" % i for i in range(V - 4)]))
src_embed = HybridSequential()
src_embed.add(Dropout(0.0))
enc, dec = get_gnmt_encoder_decoder(cell_type="lstm", hidden_size=H, dropout=0.0, num_layers=2, num_bi_layers=1)
model = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=enc, decoder=dec, embed_size=E, prefix="gnmt_", src_embed=src_embed,
                 tgt_embed=Embedding(V, E))
params = model.collect_params()
for k, v in p.items():
    params[k].shape, params[k]._data = tuple(v.shape), v.to(dev)
    params[k]._version += 1
x, vl = S.synthetic_sources(B, Tsrc, D, seed=100, min_len=64)
tr = BeamSearchTranslator(model, beam_size=beam, scorer=BeamSearchScorer(alpha=1.0, K=5), max_length=max_len)
tr.translate(x.to(dev), vl.to(dev))
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.translate(x.to(dev), vl.to(dev))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
