"""CNN backward (V7 with a trainable backbone): training-mode forward, loss, backward and running-statistics update through the
C ABI against torch.autograd over the CPU oracle in training mode (batch statistics), same seeded weights and frames.

Tolerances.  Every primitive matches torch to 2e-5 (tests/test_gpu_cnn_train_ops.py) and the gradient entering the network
(d loss / d features) to 1e-7.  Through 20-120 layers two fp32 implementations do not take identical ReLU decisions: an
activation within ~1e-5 of zero is kept by one and dropped by the other, which changes single entries of the gradients below
it by whole units (torch's own fp32 and fp64 runs of DenseNet-121 differ by > 1e-3 of max|g| in 347 of 362 tensors on this
input).  The test therefore pins the direction of every gradient tensor (cosine >= 0.995 against the fp64 oracle), bounds the
largest single-entry difference at 25 % of the tensor's largest entry, and holds the tensors above the last ReLU layers and the
loss to 1e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _load(block, p, dev):
    for k, v in p.items():
        prm = block._reg_params[k]
        prm.shape, prm._data = tuple(v.shape), v.clone().to(dev)
        prm._version += 1


@pytest.mark.parametrize("gemm", ["x3", "fp32"])
@pytest.mark.parametrize("arch", ["resnet18_v2", "densenet121"])
def test_frame_model_gradients_match_oracle_autograd(arch, gemm, monkeypatch):
    """train.py's default flow: FrameModel(backbone, 11) on single frames, SoftmaxCrossEntropyLoss, ag.backward.
    gemm = fp32: the SIMT SGEMM (24-bit operands, the parity anchor); x3: the default, split-bf16 on tcgen05 -- operands carry 18
    significant bits (hi + lo, 2^-18 relative), measured: features within 8e-5 of the fp64 oracle (2.5e-6 for fp32), which moves the
    tensors above the last ReLU by up to 3e-3 of their largest entry."""
    monkeypatch.setenv("TN_TRAIN_GEMM", gemm)
    from oracle import vision as O
    from tennis_b200 import autograd, model_zoo
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss
    from tennis_b200.models.vision.definitions import FrameModel
    dev = torch.device("cuda", 0)
    N, C = 3, 11
    p = O.synthetic_params(arch, seed=1234)
    _, x = O.synthetic_frames(N, 224, seed=100)
    g = torch.Generator().manual_seed(3)
    D = 1024 if arch == "densenet121" else 512
    cw = (torch.rand(C, D, generator=g) * 2 - 1) * 0.05
    cb = torch.randn(C, generator=g) * 0.1
    y = torch.tensor([1, 7, 4])
    # oracle: training-mode forward + autograd
    q = {k: v.clone().double().requires_grad_(not k.endswith(("running_mean", "running_var"))) for k, v in p.items()}
    q["_update_running"] = True
    cwr, cbr = cw.clone().double().requires_grad_(True), cb.clone().double().requires_grad_(True)
    feats_ref = O.FEATURES[arch](x.double(), q, training=True)
    loss_ref = torch.nn.functional.cross_entropy(feats_ref @ cwr.t() + cbr, y, reduction="none")
    loss_ref.sum().backward()
    # ours
    model = FrameModel(model_zoo.get_model(arch).features, C)
    model.initialize(ctx=dev)
    _load(model.backbone, p, dev)
    model.classes.weight.shape, model.classes.weight._data = tuple(cw.shape), cw.to(dev)
    model.classes.bias._data = cb.to(dev)
    loss_fn = SoftmaxCrossEntropyLoss()
    with autograd.record():
        out = model(x.to(dev))
        loss = loss_fn(out, y.to(dev))
    autograd.backward([loss])
    torch.cuda.synchronize()
    assert (loss.cpu().double() - loss_ref.detach()).abs().max().item() < 1e-3
    params = model.collect_params()
    worst = ("", 0.0)
    errs, coss = [], []
    for k, v in q.items():
        if k.startswith("_") or not v.requires_grad:
            continue
        if arch == "resnet18_v2" and k.startswith("bn_data."):
            continue  # BatchNorm(scale=False, center=False): fixed gamma/beta, grad_req null
        got = params["backbone." + k].grad().cpu().double()
        ref = v.grad
        rel = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        cos = float((got * ref).sum() / (got.norm() * ref.norm()).clamp(min=1e-30))
        if rel > worst[1]:
            worst = (k, rel)
        errs.append((k, rel))
        coss.append((k, cos))
    for k, ref in (("classes.weight", cwr.grad), ("classes.bias", cbr.grad)):
        got = params[k].grad().cpu().double()
        assert (got - ref).abs().max().item() < 1e-3 * max(ref.abs().max().item(), 1e-12), k
    print("%s: worst relative gradient error %.2e at %s" % (arch, worst[1], worst[0]))
    import os
    if os.environ.get("TN_DUMP_ERRS"):
        with open(os.environ["TN_DUMP_ERRS"] + "." + arch, "w") as f:
            f.write("\n".join("%s %.3e" % e for e in errs) + "\n")
    print("  first layers:", ", ".join("%s %.1e" % e for e in errs[:6]))
    print("  last layers: ", ", ".join("%s %.1e" % e for e in errs[-6:]))
    worst_cos = min(coss, key=lambda e: e[1])
    print("  lowest cosine %.6f at %s" % (worst_cos[1], worst_cos[0]))
    final_gamma = "bn5.gamma" if arch == "densenet121" else "bn_final.gamma"
    assert dict(errs)[final_gamma] < (1e-3 if gemm == "fp32" else 8e-3)  # above every ReLU decision that can differ: tight
    assert worst_cos[1] > 0.995, worst_cos
    assert worst[1] < 0.25, worst
    # running statistics: 0.9 * old + 0.1 * batch (biased variance)
    for k in ("bn0.running_mean", "bn0.running_var"):
        got = params["backbone." + k].data().cpu().double()
        assert (got - q[k]).abs().max().item() < 1e-3 * max(1.0, q[k].abs().max().item()), k


def test_cnn_rnn_end_to_end_gradients_reach_the_backbone():
    """CNNRNN with a trainable backbone (train.py --temp_pool gru without --freeze_backbone): the gradient flows from the
    classifier through max-over-time, the bi-GRU, TimeDistributed's unfold and the whole CNN."""
    from oracle import vision as O
    from tennis_b200 import autograd, model_zoo
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss
    from tennis_b200.models.vision.definitions import CNNRNN, FrameModel
    dev = torch.device("cuda", 0)
    arch, B, T, H, C = "resnet18_v2", 2, 2, 128, 11
    p = O.synthetic_params(arch, seed=1234)
    rp = O.synthetic_rnn_params("gru", 512, H, seed=4321)
    _, frames = O.synthetic_frames(B * T, 224, seed=5)
    clips = frames.reshape(B, T, 3, 224, 224)
    g = torch.Generator().manual_seed(3)
    cw = (torch.rand(C, 2 * H, generator=g) * 2 - 1) * 0.3
    cb = torch.zeros(C)
    y = torch.tensor([2, 9])
    q = {k: v.clone().double().requires_grad_(not k.endswith(("running_mean", "running_var"))) for k, v in p.items()}
    rq = {k: v.clone().double().requires_grad_(True) for k, v in rp.items()}
    out_ref = O.cnnrnn(clips.double(), lambda f: O.FEATURES[arch](f, q, training=True), rq, "gru", H, cw.double(), cb.double())
    torch.nn.functional.cross_entropy(out_ref, y, reduction="sum").backward()
    model = CNNRNN(FrameModel(model_zoo.get_model(arch).features, C), C, hidden_size=H, type="gru")
    model.initialize(ctx=dev)
    _load(model.td.model, p, dev)
    _load(model.rnn, rp, dev)
    model.classes.weight.shape, model.classes.weight._data = tuple(cw.shape), cw.to(dev)
    model.classes.bias._data = cb.to(dev)
    with autograd.record():
        out = model(clips.to(dev))
        loss = SoftmaxCrossEntropyLoss()(out, y.to(dev))
    autograd.backward([loss])
    torch.cuda.synchronize()
    assert (out.cpu().double() - out_ref.detach()).abs().max().item() < 2e-2
    params = model.collect_params()
    for k in ("conv0.weight", "stage1.block1.conv1.weight", "stage4.block2.bn2.gamma", "bn_final.beta"):
        got, ref = params["td.model." + k].grad().cpu().double(), q[k].grad
        cos = float((got * ref).sum() / (got.norm() * ref.norm()).clamp(min=1e-30))
        assert cos > 0.99, "%s: cosine %.5f" % (k, cos)  # ReLU/max-over-time decisions may differ on near-ties (see module docstring)
