"""Time one training step (forward in training mode + backward, no optimiser) of the trainable per-frame CNN
(reference train.py:415-421 without --freeze_backbone) for each GEMM mode of the training path:
    python tools/bench_cnn_train.py [arch] [frames] [modes...]      e.g.  densenet121 64 fp32 x3 bf16
CUDA events on the launch stream, 1 warm-up + 3 timed steps per mode; prints ms/step, frames/s and the conv-GEMM share
(tn_profile_* device timers: kind 0 = tensor-core GEMM kernels, kind 1 = everything else incl. the SIMT SGEMM)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402


def run(arch, frames, gemm, steps=3):
    os.environ["TN_TRAIN_GEMM"] = gemm
    from tennis_b200 import _lib, autograd, model_zoo
    from tennis_b200 import synthetic as O
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss
    from tennis_b200.models.vision.definitions import FrameModel
    dev = torch.device("cuda", 0)
    p = O.synthetic_params(arch, seed=1234)
    model = FrameModel(model_zoo.get_model(arch).features, 11)
    model.initialize(ctx=dev)
    for k, v in p.items():
        prm = model.backbone._reg_params[k]
        prm.shape, prm._data = tuple(v.shape), v.clone().to(dev)
        prm._version += 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(frames, 3, 224, 224, generator=g).to(dev)
    y = (torch.arange(frames) % 11).to(dev)
    loss_fn = SoftmaxCrossEntropyLoss()

    def step():
        with autograd.record():
            loss = loss_fn(model(x), y)
        autograd.backward([loss])
        return loss
    step()
    torch.cuda.synchronize()
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    print("%-12s %4d frames  gemm=%-4s  %9.2f ms/step  %8.1f frames/s  loss %.4f  peak mem %.1f GB  prof %s" %
          (arch, frames, gemm, ms, frames / ms * 1e3, float(loss.float().mean()), torch.cuda.max_memory_allocated() / 2**30,
           {k: (round(v, 2) if isinstance(v, float) else v) for k, v in prof.items()} if isinstance(prof, dict) else prof), flush=True)
    return ms


if __name__ == "__main__":
    arch = sys.argv[1] if len(sys.argv) > 1 else "densenet121"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    modes = sys.argv[3:] or ["fp32", "x3", "bf16"]
    for m in modes:
        run(arch, frames, m)
