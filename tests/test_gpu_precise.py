"""fp32-grade ("precise", split-bf16) CNN path and the parity claims of BASELINE.json's north_star: event logits within 1e-3 of
the reference arithmetic (fp32 oracle) on configs[1] shapes, argmax class ids bit-exact -- on clips whose logits actually differ."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CLASSES, HIDDEN = 11, 128


def _detector(precision, device="cuda"):
    """CNNRNN(FrameModel(DenseNet121.features, 11), 11, 128, 'gru') with the seeded synthetic weights (train.py:204-236)."""
    from tennis_b200 import model_zoo
    from tennis_b200 import synthetic as S
    from tennis_b200.models.vision.definitions import CNNRNN, FrameModel
    dev = torch.device(device, 0)
    backbone = model_zoo.get_model("DenseNet121", pretrained=False).features
    backbone.precision = precision
    model = CNNRNN(FrameModel(backbone, CLASSES), CLASSES, hidden_size=HIDDEN, type="gru")
    model.initialize(ctx=dev)
    p = S.synthetic_params("densenet121", seed=1234)
    for k, v in p.items():
        model.td.model._reg_params[k].set_data(v)
    rp = S.synthetic_rnn_params("gru", 1024, HIDDEN, seed=4321)
    for k, v in rp.items():
        prm = model.rnn._reg_params[k]
        prm.shape = tuple(v.shape)
        prm._data = v.to(dev).contiguous()
        prm._version += 1
    g = torch.Generator().manual_seed(77)
    cw = (torch.rand(CLASSES, 2 * HIDDEN, generator=g) * 2 - 1) * 0.5  # wide head: logits spread over the classes
    model.classes.weight.shape = (CLASSES, 2 * HIDDEN)
    model.classes.weight._data = cw.to(dev)
    model.classes.bias._data = torch.zeros(CLASSES, device=dev)
    model.collect_params().reset_ctx(dev)
    return model, p, rp, cw


@pytest.mark.parametrize("size,n", [(224, 3), (256, 1)])
def test_precise_backbone_features_match_oracle(size, n):
    from oracle import vision as O
    from tennis_b200 import ops
    p = O.synthetic_params("densenet121", seed=1234)
    u8, x = O.synthetic_frames(n, size, seed=100)
    with torch.no_grad():
        ref = O.FEATURES["densenet121"](x, p)
    bb = ops.Backbone("densenet121", O.flatten_params("densenet121", p), precision="split_bf16")
    out = bb(x.cuda())
    out8 = bb(u8.cuda())
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs()
    scale = ref.abs().max().item()
    print("precise densenet121 %d: max|ref|=%.4f max err=%.6f mean err=%.7f" % (size, scale, err.max().item(), err.mean().item()))
    assert err.max().item() < 2e-4 * max(1.0, scale)  # 16-bit mantissas + fp32 accumulation: ~50x tighter than the bf16 path
    assert (out8.cpu() - ref).abs().max().item() < 2e-4 * max(1.0, scale)
    bb.set_precision("bf16")  # the same handle switches back
    e_bf = (bb(x.cuda()).cpu() - ref).abs().max().item()
    assert 20 * err.max().item() < e_bf < 1e-2 * scale


def test_precise_rejects_resnet():
    from oracle import vision as O
    from tennis_b200 import _lib, ops
    p = O.synthetic_params("resnet18_v2", seed=1234)
    with pytest.raises(_lib.TennisB200Error):
        ops.Backbone("resnet18_v2", O.flatten_params("resnet18_v2", p), precision="split_bf16")


def test_config1_logits_1e3_and_argmax_bit_exact_over_64_clips():
    """BASELINE.json configs[1]: 64 clips x 32 frames @224, DenseNet-121 -> BiGRU(128) -> max -> Dense(11).
      * precise mode: |logit - oracle| <= 1e-3 on every one of the 64 x 11 logits (the north_star tolerance);
      * argmax class ids bit-exact over the 64 clips in BOTH modes, on structured clips whose predicted classes differ and
        whose top-1/top-2 margins are printed (the bf16 path is only required to agree where the margin exceeds its error);
      * bf16 speed mode: stated tolerance 2e-2 x max|uncentred logit| (about 2x the measured 0.96 %).
    Measured on B200 (round 2): 11 distinct classes, margins 1e-3 .. 0.5 (median 0.12); max |logit - oracle| = 2.7e-4 (precise),
    7.5e-2 on logits of magnitude 7.8 (bf16)."""
    from oracle import vision as O
    from tennis_b200 import synthetic as S
    B, T = 64, 32
    model, p, rp, cw = _detector("split_bf16")
    u8, clips = S.structured_clips(B, T, 224, seed=300)
    with torch.no_grad():
        ref0 = O.cnnrnn(clips, lambda x: O.FEATURES["densenet121"](x, p), rp, "gru", HIDDEN, cw, torch.zeros(CLASSES))
    # a random-weight network puts most of the logit mass into an input-independent component (every clip -> the same class with
    # a margin > 1); the classifier BIAS centres the logits over this batch so that the decision rests on what differs between
    # clips: 8+ distinct classes with margins from 1e-2 to 0.5 -- a comparison that can fail
    bias = -ref0.mean(dim=0)
    model.classes.bias._data = bias.cuda()
    model.classes.bias._version += 1
    ref = ref0 + bias
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    classes_hit = ref.argmax(1).unique().numel()
    out_p = torch.cat([model(clips[i:i + 16].cuda()).cpu() for i in range(0, B, 16)])
    model.td.model.precision = "bf16"
    out_b = torch.cat([model(clips[i:i + 16].cuda()).cpu() for i in range(0, B, 16)])
    torch.cuda.synchronize()
    e_p = (out_p - ref).abs().max().item()
    e_b = (out_b - ref).abs().max().item()
    print("configs[1] parity: %d distinct argmax classes over %d clips, top-1/top-2 margin min %.4f median %.4f, max|logit| %.3f; "
          "max |logit - oracle|: precise %.2e, bf16 %.2e" % (classes_hit, B, margin.min().item(), margin.median().item(),
                                                            ref.abs().max().item(), e_p, e_b))
    assert classes_hit >= 6, "the test inputs must exercise several classes"
    assert e_p <= 1e-3
    assert torch.equal(out_p.argmax(1), ref.argmax(1))
    assert e_b <= 2e-2 * max(1.0, ref0.abs().max().item())
    safe = margin > 2 * e_b  # where the bf16 error cannot flip the decision it must not
    assert torch.equal(out_b.argmax(1)[safe], ref.argmax(1)[safe]) and int(safe.sum()) >= 8
    print("bf16 mode: %d / %d clips have a margin above twice its error; argmax equal on all of them (and on %d / %d overall)"
          % (int(safe.sum()), B, int((out_b.argmax(1) == ref.argmax(1)).sum()), B))
