"""Primitives of the CNN training path (csrc/tn_cnn_train.cu through models/vision/train_graph.py) one by one against
torch.autograd on random fp32 data: training-mode BatchNorm(+ReLU) on a channel slice, convolutions (1x1 direct, 3x3 padded,
strided, 7x7/2 stem) as im2col + SGEMM, max / average pooling."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class _P(object):
    """Stand-in for gluon.Parameter as the graph uses it."""

    def __init__(self, t, grad_req="write"):
        self._data, self.grad_req, self._grad, self._version = t, grad_req, None, 0

    def data(self):
        return self._data

    def _accumulate_grad(self, g):
        if self.grad_req != "null":
            self._grad = g.clone() if self._grad is None else self._grad + g

    def _bump(self):
        self._version += 1


class _Blk(object):
    arch = "test"

    def __init__(self, params):
        self._reg_params = params


def _graph(params):
    from tennis_b200.models.vision.train_graph import CNNTrainGraph
    return CNNTrainGraph(_Blk(params))


def _act(t_nchw, needs_grad=True):
    from tennis_b200.models.vision.train_graph import _Act
    return _Act(t_nchw.permute(0, 2, 3, 1).contiguous().cuda(), needs_grad=needs_grad)


def _nchw(t_nhwc):
    return t_nhwc.permute(0, 3, 1, 2).cpu()


def _close(a, b, tol=2e-5):
    scale = max(1.0, b.abs().max().item())
    assert (a - b).abs().max().item() < tol * scale, ((a - b).abs().max().item(), scale)


@pytest.mark.parametrize("relu", [True, False])
def test_bn_relu_slice_forward_backward(relu):
    g = torch.Generator().manual_seed(0)
    N, C, H, W, c0, Cs = 3, 40, 5, 7, 8, 24   # BN over channels [8, 32) of a 40-channel buffer
    x = torch.randn(N, C, H, W, generator=g) * 3 + 1
    gamma, beta = torch.rand(Cs, generator=g) + 0.5, torch.randn(Cs, generator=g)
    rm, rv = torch.randn(Cs, generator=g), torch.rand(Cs, generator=g) + 0.5
    dy = torch.randn(N, Cs, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = F.batch_norm(xr[:, c0:c0 + Cs], None, None, gr, br, training=True, eps=1e-5)
    if relu:
        y_ref = y_ref.relu()
    (y_ref * dy).sum().backward()
    P = {"bn.gamma": _P(gamma.cuda()), "bn.beta": _P(beta.cuda()), "bn.running_mean": _P(rm.clone().cuda(), "null"),
         "bn.running_var": _P(rv.clone().cuda(), "null")}
    G = _graph(P)
    src = _act(x)
    out = G._bn(src, c0, Cs, "bn", relu=relu)
    _close(_nchw(out.data), y_ref.detach())
    out.grad = dy.permute(0, 2, 3, 1).contiguous().cuda()
    for fn in reversed(G.tape):
        fn()
    _close(P["bn.gamma"]._grad.cpu(), gr.grad)
    _close(P["bn.beta"]._grad.cpu(), br.grad)
    _close(_nchw(src.grad), xr.grad)
    xs = x[:, c0:c0 + Cs]
    _close(P["bn.running_mean"].data().cpu(), 0.9 * rm + 0.1 * xs.mean(dim=(0, 2, 3)))
    _close(P["bn.running_var"].data().cpu(), 0.9 * rv + 0.1 * xs.var(dim=(0, 2, 3), unbiased=False))


@pytest.mark.parametrize("Cin,Cout,R,stride,pad,H,W", [(24, 32, 1, 1, 0, 6, 5), (16, 32, 3, 1, 1, 7, 7), (16, 48, 3, 2, 1, 9, 8),
                                                       (32, 64, 1, 2, 0, 8, 8), (3, 64, 7, 2, 3, 20, 17), (512, 64, 3, 1, 1, 7, 7),
                                                       (128, 32, 3, 1, 1, 28, 28), (200, 128, 1, 1, 0, 14, 14)])
@pytest.mark.parametrize("gemm", ["x3", "fp32", "bf16"])
def test_conv_forward_backward(Cin, Cout, R, stride, pad, H, W, gemm, monkeypatch):
    """x3 = split-bf16 on tcgen05 (the default), fp32 = SIMT SGEMM: both at the fp32 bar (1e-4 of the largest entry); bf16 = one
    tensor-core product: the bf16 bar."""
    monkeypatch.setenv("TN_TRAIN_GEMM", gemm)
    tol = 1e-4 if gemm != "bf16" else 2e-2
    g = torch.Generator().manual_seed(1)
    N, Ct, c0 = 3, Cin + 8, 4            # the conv reads channels [4, 4+Cin) of a wider buffer ...
    Cd, d0 = Cout + 16, 8                # ... and writes channels [8, 8+Cout) of a wider destination
    x = torch.randn(N, Ct, H, W, generator=g)
    w = torch.randn(Cout, Cin, R, R, generator=g) * 0.1
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, Cout, Ho, Wo, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = F.conv2d(xr[:, c0:c0 + Cin], wr, stride=stride, padding=pad)
    (y_ref * dy).sum().backward()
    from tennis_b200.models.vision.train_graph import _Act
    P = {"w": _P(w.cuda())}
    G = _graph(P)
    src = _act(x)
    dst = _Act(torch.zeros(N, Ho, Wo, Cd).cuda())
    G._conv(src, c0, Cin, "w", stride, pad, dst=dst, d0=d0)
    _close(_nchw(dst.data)[:, d0:d0 + Cout], y_ref.detach(), tol)
    assert (dst.data[..., :d0] == 0).all() and (dst.data[..., d0 + Cout:] == 0).all()
    dst.grad = torch.zeros_like(dst.data)
    dst.grad[..., d0:d0 + Cout] = dy.permute(0, 2, 3, 1).cuda()
    for fn in reversed(G.tape):
        fn()
    _close(P["w"]._grad.cpu(), wr.grad, tol)
    _close(_nchw(src.grad), xr.grad, tol)


@pytest.mark.parametrize("ta,tb", [(False, True), (True, False), (False, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(300, 32, 64), (129, 100, 200), (64, 128, 5000), (1000, 257, 72)])
def test_tensor_core_gemm_matches_fp64(ta, tb, M, N, K):
    """tn_split_bf16 + tn_gemm_tc on strided fp32 views, every transpose combination, ragged sizes, split-K (K = 5000), alpha/beta:
    split-bf16 within 3e-5 of the fp64 product (fp32 SGEMM grade), plain bf16 within 1e-2."""
    from tennis_b200 import tcgemm
    g = torch.Generator().manual_seed(5)
    A = torch.randn((K, M + 3) if ta else (M, K + 5), generator=g)
    B = torch.randn((N, K + 2) if tb else (K, N + 7), generator=g)
    C0 = torch.randn(M, N + 4, generator=g)
    Av = A[:, :M] if ta else A[:, :K]
    Bv = B[:, :K] if tb else B[:, :N]
    opA = Av.t() if ta else Av
    opB = Bv.t() if tb else Bv
    ref = 0.5 * (opA.double() @ opB.double()) + 2.0 * C0[:, :N].double()
    scale = ref.abs().max().item()
    for passes, tol in ((3, 3e-5), (1, 1e-2)):
        Ad, Bd, Cd = A.cuda(), B.cuda(), C0.clone().cuda()
        tcgemm.matmul(Ad[:, :M] if ta else Ad[:, :K], Bd[:, :K] if tb else Bd[:, :N], Cd[:, :N], ta=ta, tb=tb, alpha=0.5, beta=2.0,
                      passes=passes)
        err = (Cd[:, :N].cpu().double() - ref).abs().max().item()
        assert err < tol * scale, (passes, err, scale)
        assert torch.equal(Cd[:, N:].cpu(), C0[:, N:])  # columns outside the view untouched


def test_pooling_forward_backward():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 16, 11, 9, generator=g).relu()  # zeros create ties
    G = _graph({})
    for kind in ("max", "avg2", "avg_global"):
        xr = x.clone().requires_grad_(True)
        if kind == "max":
            y_ref = F.max_pool2d(xr, 3, 2, 1)
        elif kind == "avg2":
            y_ref = F.avg_pool2d(xr, 2, 2)
        else:
            y_ref = F.avg_pool2d(xr, (11, 9))
        dy = torch.randn(y_ref.shape, generator=g)
        (y_ref * dy).sum().backward()
        src = _act(x)
        out = G._maxpool(src, 3, 2, 1) if kind == "max" else G._avgpool(src, 2, 2) if kind == "avg2" else G._avgpool(src, 11, 9)
        _close(_nchw(out.data), y_ref.detach())
        out.grad = dy.permute(0, 2, 3, 1).contiguous().cuda()
        G.tape[-1]()
        _close(_nchw(src.grad), xr.grad)
