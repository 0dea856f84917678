"""One head-training step (CNNRNN on features: forward-with-save, softmax CE, backward, SGD) between cudaProfilerStart/Stop."""
import sys

import torch

sys.path.insert(0, ".")
from tennis_b200 import synthetic as O  # noqa: E402
from tennis_b200 import autograd  # noqa: E402
from tennis_b200.gluon import SoftmaxCrossEntropyLoss, Trainer  # noqa: E402
from tennis_b200.models.vision.definitions import CNNRNN  # noqa: E402

dev = torch.device("cuda", 0)
B, T, D, H = 256, 32, 1024, 128
g = torch.Generator().manual_seed(11)
feats = torch.randn(B, T, D, generator=g).relu().to(dev)
labels = torch.randint(0, 11, (B,), generator=g).to(dev)
head = CNNRNN(None, 11, hidden_size=H, type="gru")
head.initialize(ctx=dev)
for k, v in O.synthetic_rnn_params("gru", D, H, seed=4321).items():
    prm = head.rnn._reg_params[k]
    prm.shape, prm._data = tuple(v.shape), v.to(dev)
    prm._version += 1
head(feats)
loss_fn = SoftmaxCrossEntropyLoss()
tr = Trainer(head.collect_params(), 'sgd', {'learning_rate': 1e-3, 'momentum': 0.9, 'wd': 1e-4})


def step():
    with autograd.record():
        loss = loss_fn(head(feats), labels)
    autograd.backward([loss])
    tr.step(B)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
