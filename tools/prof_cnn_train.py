"""Per-kernel GPU time of one CNN training step through torch.profiler (CUPTI activity records: no replay, seconds to run).
usage: python tools/prof_cnn_train.py [arch] [frames] [mode]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
arch = sys.argv[1] if len(sys.argv) > 1 else "densenet121"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.environ["TN_TRAIN_GEMM"] = sys.argv[3] if len(sys.argv) > 3 else "x3"
from tennis_b200 import autograd, model_zoo  # noqa: E402
from tennis_b200 import synthetic as O  # noqa: E402
from tennis_b200.gluon import SoftmaxCrossEntropyLoss  # noqa: E402
from tennis_b200.models.vision.definitions import FrameModel  # noqa: E402

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
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0].split("<")[0].split("::")[-1][:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("%-62s %6s %10s %6s" % ("kernel", "n", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-62s %6d %10.3f %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
print("TOTAL %d launches %.3f ms" % (sum(v[0] for v in agg.values()), tot / 1e3))
