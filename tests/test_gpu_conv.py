"""Parity of the tcgen05 implicit-GEMM convolution (through the C ABI) against fp32 torch-CPU convolution
on identical bf16-rounded operands.  Every geometry the CNN plans use is covered."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


def _mk(n, H, W, C, cs=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    cs = cs or C
    x = _bf(torch.randn(n, H, W, cs, generator=g))
    return x


def _ref_conv(x_nhwc, w, stride, pad, pro=None, pro_relu=True, epi=None, epi_relu=False, residual=None, pool2=False):
    Cin = w.shape[1]
    a = x_nhwc[..., :Cin]
    if pro is not None:
        a = a * pro[0] + pro[1]
        if pro_relu:
            a = a.relu()
        if not pool2:
            a = _bf(a)
    a = a.permute(0, 3, 1, 2)
    if pool2:
        a = _bf(F.avg_pool2d(a, 2, 2))
    # the library folds the epilogue scale (BN gamma/sigma) into the weights BEFORE rounding them to bf16
    wq = _bf(w if epi is None else w * epi[0].reshape(-1, 1, 1, 1))
    y = F.conv2d(a, wq, stride=1 if pool2 else stride, padding=0 if pool2 else pad).permute(0, 2, 3, 1)
    if epi is not None:
        y = y + epi[1]
    if residual is not None:
        y = y + residual
    if epi_relu:
        y = y.relu()
    return y


CASES = [
    # name, n, H, W, Cin, in_cs, Cout, R, stride, pad, pro, epi, epi_relu
    ("1x1_plain", 2, 12, 12, 64, 64, 128, 1, 1, 0, False, False, False),
    ("1x1_k96_pro_epi", 3, 9, 7, 96, 256, 128, 1, 1, 0, True, True, True),
    ("1x1_k992", 1, 14, 14, 992, 1024, 128, 1, 1, 0, True, True, True),
    ("3x3_bott", 2, 14, 14, 128, 128, 32, 3, 1, 1, False, False, False),
    ("3x3_s2", 2, 15, 13, 64, 64, 64, 3, 2, 1, True, True, True),
    ("3x3_c256", 1, 10, 10, 128, 128, 256, 3, 1, 1, True, False, False),
    ("1x1_c512", 1, 16, 16, 256, 256, 512, 1, 2, 0, True, False, False),
    ("1x1_n768", 1, 1, 300, 1024, 1024, 768, 1, 1, 0, False, True, False),
    # >= 4*148 M tiles and K > 256: two M tiles share each streamed weight chunk (odd tile count -> ragged last pair)
    ("1x1_k320_paired", 25, 56, 56, 320, 320, 128, 1, 1, 0, True, True, True),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_modes(case):
    from tennis_b200 import ops
    name, n, H, W, Cin, cs, Cout, R, stride, pad, pro, epi, epi_relu = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = _mk(n, H, W, Cin, cs, seed=1)
    w = torch.randn(Cout, Cin, R, R, generator=g) * (2.0 / (Cin * R * R)) ** 0.5
    ps = (torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.1) if pro else None
    es = (torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1) if epi else None
    conv = ops.Conv(w, pro_scale=ps[0] if pro else None, pro_shift=ps[1] if pro else None,
                    epi_scale=es[0] if epi else None, epi_shift=es[1] if epi else None)
    y = conv(x.cuda().to(torch.bfloat16), stride=stride, pad=pad, epi_relu=epi_relu, out_fp32=True)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, stride, pad, ps, True, es, epi_relu)
    assert y.shape == ref.shape
    err = (y.cpu() - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), "%s: max abs err %g" % (name, err)


def test_conv_inplace_concat_bf16_out():
    """DenseNet concat: 32 new channels written at a channel offset of a wider buffer; neighbours untouched."""
    from tennis_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = _mk(2, 14, 14, 128, seed=2)
    w = torch.randn(32, 128, 3, 3, generator=g) * 0.03
    buf = torch.full((2, 14, 14, 256), 7.0, dtype=torch.bfloat16).cuda()
    conv = ops.Conv(w)
    conv(x.cuda().to(torch.bfloat16), stride=1, pad=1, out=buf, out_coff=96)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, 1, 1)
    out = buf.float().cpu()
    assert (out[..., :96] == 7.0).all() and (out[..., 128:] == 7.0).all()
    assert (out[..., 96:128] - ref).abs().max().item() < 2e-2


def test_conv_residual():
    from tennis_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = _mk(2, 8, 8, 64, seed=3)
    r = _mk(2, 8, 8, 64, seed=4)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    conv = ops.Conv(w)
    y = conv(x.cuda().to(torch.bfloat16), stride=1, pad=1, out_fp32=True, residual=r.cuda().to(torch.bfloat16))
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, 1, 1, residual=r)
    assert (y.cpu() - ref).abs().max().item() < 2e-3


def test_conv_pool2_transition():
    from tennis_b200 import _lib, ops
    g = torch.Generator().manual_seed(7)
    x = _mk(2, 14, 14, 256, seed=5)
    w = torch.randn(128, 256, 1, 1, generator=g) * 0.08
    ps = (torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1)
    conv = ops.Conv(w, mode=_lib.MODE_POOL2, pro_scale=ps[0], pro_shift=ps[1])
    y = conv(x.cuda().to(torch.bfloat16), out_fp32=True)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, 1, 0, ps, True, pool2=True)
    assert y.shape == ref.shape == (2, 7, 7, 128)
    assert (y.cpu() - ref).abs().max().item() < 3e-3
    # and it equals avgpool(conv1x1(relu(bn(x)))) -- the reference order of operations -- up to bf16 rounding
    a = (x * ps[0] + ps[1]).relu().permute(0, 3, 1, 2)
    ref2 = F.avg_pool2d(F.conv2d(a, w), 2, 2).permute(0, 2, 3, 1)
    assert (y.cpu() - ref2).abs().max().item() < 3e-2


def test_conv_stem():
    from tennis_b200 import _lib, ops
    g = torch.Generator().manual_seed(8)
    frames = _bf(torch.randn(2, 3, 37, 45, generator=g))
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    es = (torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1)
    x4 = ops.frames_to_nhwc4(frames.cuda())
    assert x4.shape == (2, 37, 45, 4)
    assert (x4[..., :3].float().cpu() == frames.permute(0, 2, 3, 1)).all() and (x4[..., 3] == 0).all()
    conv = ops.Conv(w, mode=_lib.MODE_STEM, epi_scale=es[0], epi_shift=es[1])
    y = conv(x4, stride=2, pad=3, epi_relu=True, out_fp32=True)
    torch.cuda.synchronize()
    ref = (F.conv2d(frames, _bf(w * es[0].reshape(-1, 1, 1, 1)), stride=2, padding=3).permute(0, 2, 3, 1) + es[1]).relu()
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max().item() < 3e-3


def test_no_cpu_fallback():
    from tennis_b200 import _lib, ops
    with pytest.raises(_lib.TennisB200Error):
        ops.dense(torch.zeros(2, 4), torch.zeros(3, 4))
