"""One training step (forward in training mode + backward) of the trainable DenseNet-121 between cudaProfilerStart/Stop, for
`ncu --metrics gpu__time_duration.sum,... --profile-from-start off`.  usage: python tools/ncu_cnn_train_target.py [arch] [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tennis_b200 import autograd, model_zoo  # noqa: E402
from tennis_b200 import synthetic as O  # noqa: E402
from tennis_b200.gluon import SoftmaxCrossEntropyLoss  # noqa: E402
from tennis_b200.models.vision.definitions import FrameModel  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "densenet121"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
p = O.synthetic_params(arch, seed=1234)
model = FrameModel(model_zoo.get_model(arch).features, 11)
model.initialize(ctx=dev)
for k, v in p.items():
    prm = model.backbone._reg_params[k]
    prm.shape, prm._data = tuple(v.shape), v.clone().to(dev)
    prm._version += 1
x = torch.randn(frames, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
y = (torch.arange(frames) % 11).to(dev)
loss_fn = SoftmaxCrossEntropyLoss()


def step():
    with autograd.record():
        loss = loss_fn(model(x), y)
    autograd.backward([loss])


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
