"""Training graph of the per-frame CNN: forward in training mode (BatchNorm on batch statistics, running statistics updated)
with saved activations, and the backward pass, for the two backbones of the scripts (reference train.py:204 -> gluoncv
DenseNet-121 / ResNet-18 v2 `features`).

fp32 NHWC activations (a dense block is one buffer; layers read channel prefixes and write channel slices through row strides),
SIMT BatchNorm / pooling kernels (csrc/tn_cnn_train.cu) and the convolutions' forward / data-gradient / weight-gradient
contractions on the tensor cores (csrc/tn_gemm_tc.cu through tennis_b200/tcgemm.py): split-bf16 operand planes, three tcgen05
products per contraction ($TN_TRAIN_GEMM = x3, the default; bf16 = one product; fp32 = the SIMT SGEMM of round 1, the parity
anchor).  3x3 / stride-1 convolutions never form an im2col matrix: a tap is a row shift in zero-padded row space.  Python only
sequences C-ABI calls; torch tensors are containers (allocation, views, layout permutes).  The inference path (bf16 tcgen05
kernels) is untouched.  Gradients are checked against torch.autograd of the fp64 oracle in tests/test_gpu_cnn_train.py.
"""
import torch

from ..._lib import check, dptr, lib, stream_ptr

BN_EPS = 1e-5
BN_MOMENTUM = 0.9  # running = 0.9 * running + 0.1 * batch (Gluon BatchNorm default)
_DENSE_CFG = (6, 12, 24, 16)
_RES_CH = (64, 128, 256, 512)


class _Act(object):
    """NHWC fp32 activation buffer (N,H,W,C) with a lazily allocated gradient of the same shape."""

    def __init__(self, data, needs_grad=True):
        self.data = data
        self.grad = None
        self.needs_grad = needs_grad  # False: nothing upstream is trainable (the input frames)

    @property
    def shape(self):
        return tuple(self.data.shape)

    def g(self):
        if self.grad is None:
            self.grad = torch.zeros_like(self.data)
        return self.grad


def _sl(t, c0):
    """Pointer to channel c0 of an NHWC buffer (row stride stays the buffer's channel count)."""
    return dptr(t.reshape(-1, t.shape[-1])[:, c0:])


class CNNTrainGraph(object):
    def __init__(self, features_block):
        self.blk = features_block
        self.P = features_block._reg_params
        self.tape = []
        self.touched_stats = []
        self.named = {}  # block outputs by name (kept for inspection by the tests)
        from ... import tcgemm
        self.gemm_mode = tcgemm.mode()  # 'x3' / 'bf16': tcgen05 (csrc/tn_gemm_tc.cu); 'fp32': the SIMT SGEMM

    # ------------------------------------------------------------------------------------------------ ops
    def _bn(self, src, c0, C, prefix, relu=True):
        N, H, W, Ct = src.shape
        M = N * H * W
        P = self.P
        gamma, beta = P[prefix + ".gamma"], P[prefix + ".beta"]
        rm, rv = P[prefix + ".running_mean"], P[prefix + ".running_var"]
        dev = src.data.device
        out = _Act(torch.empty(N, H, W, C, device=dev))
        mean, var = torch.empty(C, device=dev), torch.empty(C, device=dev)
        check(lib().tn_bn_train_forward(_sl(src.data, c0), Ct, M, C, dptr(gamma.data()), dptr(beta.data()), BN_EPS, BN_MOMENTUM,
                                        dptr(rm.data()), dptr(rv.data()), int(relu), dptr(mean), dptr(var), dptr(out.data), C,
                                        stream_ptr()))
        self.touched_stats += [rm, rv]

        def bwd():
            if out.grad is None:
                return
            dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
            # dx is accumulated into the source's gradient (a dense block's buffer collects it from every later layer);
            # a source without trainable ancestors still gets a scratch dx (the kernel computes both in one pass)
            dx = src.g() if src.needs_grad else torch.zeros_like(src.data)
            check(lib().tn_bn_train_backward(_sl(src.data, c0), Ct, dptr(out.data), C, dptr(out.grad), C, M, C, dptr(mean), dptr(var),
                                             dptr(gamma.data()), BN_EPS, int(relu), dptr(dg), dptr(db), _sl(dx, c0), Ct, 1,
                                             stream_ptr()))
            gamma._accumulate_grad(dg)
            beta._accumulate_grad(db)
            out.grad = None
        self.tape.append(bwd)
        return out

    def _conv(self, src, c0, Cin, wname, stride, pad, dst=None, d0=0):
        from ..captioning.train_graph import sgemm
        from ... import tcgemm
        N, H, W, Ct = src.shape
        Wp = self.P[wname]
        w = Wp.data()
        Cout, _, R, S = w.shape
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        dev = src.data.device
        if dst is None:
            dst = _Act(torch.empty(N, Ho, Wo, Cout, device=dev))
        Cd = dst.shape[3]
        K = R * S * Cin
        wk = w.permute(0, 2, 3, 1).reshape(Cout, K).contiguous()  # (Cout, (r,s,c)): the im2col column order
        direct = R == 1 and S == 1 and stride == 1 and pad == 0
        x2 = src.data.reshape(-1, Ct)[:, c0:c0 + Cin]
        y2 = dst.data.reshape(-1, Cd)[:, d0:d0 + Cout]
        gm = self.gemm_mode
        passes = 3 if gm == "x3" else 1
        lo = passes == 3
        # 3x3 / stride 1 / pad 1 on the tensor cores: a tap is a row shift in the zero-padded row space, no im2col matrix
        shifted = gm != "fp32" and R == 3 and S == 3 and stride == 1 and pad == 1 and Cin % 8 == 0 and Cout % 8 == 0
        if shifted:
            Mp = tcgemm.padded_rows(N, H, W)

        def columns():
            if direct:
                return x2
            col = torch.empty(N * Ho * Wo, K, device=dev)
            check(lib().tn_im2col_nhwc(_sl(src.data, c0), Ct, N, H, W, Cin, R, S, stride, pad, dptr(col), stream_ptr()))
            return col

        if shifted:
            xp = tcgemm.planes(x2, pad_hw=(H, W), lo=lo)
            wp = tcgemm.planes(wk, lo=lo)
            tcgemm.gemm(xp, wp, Mp, Cout, Cin, y2, Cd, taps=tcgemm.taps_conv3x3_forward(W, Cin), unpad_hw=(H, W), passes=passes)
            del xp, wp
        elif gm != "fp32":
            tcgemm.matmul(columns(), wk, y2, tb=True, passes=passes)
        else:
            sgemm(columns(), wk, y2, tb=True)

        def bwd():
            if dst.grad is None:
                return
            dy2 = dst.grad.reshape(-1, Cd)[:, d0:d0 + Cout]
            want_dw = Wp.grad_req != 'null'
            if shifted:
                if want_dw:
                    # dW[n, (dy,dx), c] = sum_m dY[m, n] X[m + dy*P + dx, c]: contraction over the padded pixels (row pitch P a
                    # multiple of 8), one GEMM per tap: the dx shift is baked into three copies of dY^T, dy*P is the TMA offset
                    P8 = tcgemm.pitch8(W)
                    Mp8 = tcgemm.padded_rows(N, H, W, P8)
                    xT = tcgemm.planes(x2, transpose=True, pad_hw=(H, W), lo=lo, pitch=P8)                       # (Cin, Mp8)
                    dyT = tcgemm.empty_planes(3 * Cout, Mp8, dev, lo=lo)                                         # (dx, Cout) x Mp8
                    for dx in (-1, 0, 1):
                        tcgemm.planes(dy2, transpose=True, pad_hw=(H, W), lo=lo, pitch=P8, shift=dx, into=dyT, row0=(dx + 1) * Cout)
                    dwk = torch.empty(Cout, K, device=dev)
                    # tap t = (dy, dx): B rows (dx+1)*Cout.., contraction offset -dy*P8; D_t[c, n] -> dwk[n, t*Cin + c]
                    tcgemm.gemm(xT, dyT, Cin, Cout, Mp8, dwk, 1, K, taps=tcgemm.taps_conv3x3_wgrad(W, Cout), passes=passes,
                                tile_taps=True, c_tap_stride=Cin)
                    del xT, dyT
                    Wp._accumulate_grad(dwk.reshape(Cout, R, S, Cin).permute(0, 3, 1, 2).contiguous())
                if src.needs_grad:
                    # dX[m, c] += sum_t sum_n dY[m - off_t, n] W[n, t, c]
                    dyp = tcgemm.planes(dy2, pad_hw=(H, W), lo=lo)                      # (Mp, Cout)
                    wT = tcgemm.planes(w.permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).contiguous(), lo=lo)  # (Cin, (t, n))
                    tcgemm.gemm(dyp, wT, Mp, Cin, Cout, src.g().reshape(-1, Ct)[:, c0:c0 + Cin], Ct,
                                taps=tcgemm.taps_conv3x3_dgrad(W, Cout), beta=1.0, unpad_hw=(H, W), passes=passes)
                return
            mm = sgemm if gm == "fp32" else (lambda A, B, C, **kw: tcgemm.matmul(A, B, C, passes=passes, **kw))
            col = columns()  # recomputed: keeping every im2col matrix would dominate the memory
            if want_dw:
                dwk = mm(dy2, col, torch.empty(Cout, K, device=dev), ta=True)
                Wp._accumulate_grad(dwk.reshape(Cout, R, S, Cin).permute(0, 3, 1, 2).contiguous())
            if src.needs_grad:
                if direct:
                    mm(dy2, wk, src.g().reshape(-1, Ct)[:, c0:c0 + Cin], beta=1.0)
                else:
                    dcol = mm(dy2, wk, col)  # the im2col buffer is free again: reuse it for its gradient
                    check(lib().tn_col2im_nhwc(dptr(dcol), N, H, W, Cin, R, S, stride, pad, _sl(src.g(), c0), Ct, stream_ptr()))
        self.tape.append(bwd)
        return dst

    def _maxpool(self, src, k, stride, pad, dst=None):
        N, H, W, C = src.shape
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        dev = src.data.device
        if dst is None:
            dst = _Act(torch.empty(N, Ho, Wo, C, device=dev))
        Cd = dst.shape[3]
        idx = torch.empty(N * Ho * Wo * C, dtype=torch.int32, device=dev)
        check(lib().tn_maxpool_nhwc_forward(dptr(src.data), C, N, H, W, C, k, stride, pad, dptr(dst.data), Cd, dptr(idx), stream_ptr()))

        def bwd():
            if dst.grad is None:
                return
            check(lib().tn_maxpool_nhwc_backward(dptr(dst.grad), Cd, dptr(idx), N * Ho * Wo, C, dptr(src.g()), C, stream_ptr()))
        self.tape.append(bwd)
        return dst

    def _avgpool(self, src, kh, kw, dst=None):
        N, H, W, C = src.shape
        Ho, Wo = H // kh, W // kw
        dev = src.data.device
        if dst is None:
            dst = _Act(torch.empty(N, Ho, Wo, C, device=dev))
        Cd = dst.shape[3]
        check(lib().tn_avgpool_nhwc_forward(dptr(src.data), C, N, H, W, C, kh, kw, dptr(dst.data), Cd, stream_ptr()))

        def bwd():
            if dst.grad is None:
                return
            check(lib().tn_avgpool_nhwc_backward(dptr(dst.grad), Cd, N, H, W, C, kh, kw, dptr(src.g()), C, 1, stream_ptr()))
        self.tape.append(bwd)
        return dst

    def _add_into(self, dst, res):
        """dst += res (ResNet shortcut); the gradient of dst flows to both."""
        check(lib().tn_axpy(dptr(dst.data), dptr(res.data), 1.0, dst.data.numel(), stream_ptr()))

        def bwd():
            if dst.grad is None:
                return
            check(lib().tn_axpy(dptr(res.g()), dptr(dst.grad), 1.0, dst.grad.numel(), stream_ptr()))
        self.tape.append(bwd)

    # ------------------------------------------------------------------------------------------------ networks
    def forward(self, x):
        """x: (N,3,H,W) fp32 normalised frames -> (N,D) features (training mode)."""
        x0 = _Act(x.float().permute(0, 2, 3, 1).contiguous(), needs_grad=False)  # frames are leaves
        out = self._densenet(x0) if self.blk.arch == "densenet121" else self._resnet(x0)
        for p in self.touched_stats:
            p._bump()  # the inference engine must re-pack the updated running statistics
        self.out = out
        return out.data

    def _densenet(self, x0):
        dev = x0.data.device
        N = x0.shape[0]
        a = self._conv(x0, 0, 3, "conv0.weight", 2, 3)
        a = self._bn(a, 0, 64, "bn0")
        c = 64
        blk = None
        for b, nl in enumerate(_DENSE_CFG):
            ctot = c + 32 * nl
            if b == 0:
                Hs, Ws = a.shape[1], a.shape[2]
                Hb, Wb = (Hs + 2 - 3) // 2 + 1, (Ws + 2 - 3) // 2 + 1
                blk = _Act(torch.empty(N, Hb, Wb, ctot, device=dev))
                self._maxpool_into(a, blk)
            for l in range(nl):
                pre = "block%d.layer%d" % (b + 1, l + 1)
                t = self._bn(blk, 0, c, pre + ".bn1")
                t = self._conv(t, 0, c, pre + ".conv1.weight", 1, 0)
                t = self._bn(t, 0, 128, pre + ".bn2")
                self._conv(t, 0, 128, pre + ".conv2.weight", 1, 1, dst=blk, d0=c)
                c += 32
            if b < 3:
                pre = "trans%d" % (b + 1)
                t = self._bn(blk, 0, c, pre + ".bn")
                t = self._conv(t, 0, c, pre + ".conv.weight", 1, 0)
                c //= 2
                nxt = _Act(torch.empty(N, t.shape[1] // 2, t.shape[2] // 2, c + 32 * _DENSE_CFG[b + 1], device=dev))
                self._avgpool_into(t, 2, 2, nxt)
                blk = nxt
        t = self._bn(blk, 0, c, "bn5")
        t = self._avgpool(t, 7, 7)
        return self._flatten_channel_major(t)

    def _resnet(self, x0):
        a = self._bn(x0, 0, 3, "bn_data", relu=False)
        a.needs_grad = False  # bn_data has no trainable parameters (scale=False, center=False)
        a = self._conv(a, 0, 3, "conv0.weight", 2, 3)
        a = self._bn(a, 0, 64, "bn0")
        x = self._maxpool(a, 3, 2, 1)
        cin = 64
        for s, c in enumerate(_RES_CH):
            for b in range(2):
                pre = "stage%d.block%d" % (s + 1, b + 1)
                stride = 2 if (b == 0 and s > 0) else 1
                y = self._bn(x, 0, cin, pre + ".bn1")
                residual = x
                if (pre + ".downsample.weight") in self.P:
                    residual = self._conv(y, 0, cin, pre + ".downsample.weight", stride, 0)
                t = self._conv(y, 0, cin, pre + ".conv1.weight", stride, 1)
                t = self._bn(t, 0, c, pre + ".bn2")
                t = self._conv(t, 0, c, pre + ".conv2.weight", 1, 1)
                self._add_into(t, residual)
                self.named[pre] = t
                x = t
                cin = c
        t = self._bn(x, 0, 512, "bn_final")
        t = self._avgpool(t, t.shape[1], t.shape[2])
        return self._flatten_channel_major(t)

    def _maxpool_into(self, src, blk):
        """max-pool 3/2/1 of the stem output written into channels [0, C) of the first dense block's buffer."""
        N, H, W, C = src.shape
        Ho, Wo, Cd = blk.shape[1], blk.shape[2], blk.shape[3]
        idx = torch.empty(N * Ho * Wo * C, dtype=torch.int32, device=src.data.device)
        check(lib().tn_maxpool_nhwc_forward(dptr(src.data), C, N, H, W, C, 3, 2, 1, dptr(blk.data), Cd, dptr(idx), stream_ptr()))

        def bwd():
            if blk.grad is None:
                return
            check(lib().tn_maxpool_nhwc_backward(dptr(blk.grad), Cd, dptr(idx), N * Ho * Wo, C, dptr(src.g()), C, stream_ptr()))
        self.tape.append(bwd)

    def _avgpool_into(self, src, kh, kw, blk):
        N, H, W, C = src.shape
        Cd = blk.shape[3]
        check(lib().tn_avgpool_nhwc_forward(dptr(src.data), C, N, H, W, C, kh, kw, dptr(blk.data), Cd, stream_ptr()))

        def bwd():
            if blk.grad is None:
                return
            check(lib().tn_avgpool_nhwc_backward(dptr(blk.grad), Cd, N, H, W, C, kh, kw, dptr(src.g()), C, 1, stream_ptr()))
        self.tape.append(bwd)

    def _flatten_channel_major(self, t):
        """Gluon Flatten on NCHW: (N,ph,pw,C) NHWC -> (N, C*ph*pw) with channel-major order (pure layout)."""
        N, ph, pw, C = t.shape
        out = _Act(t.data.permute(0, 3, 1, 2).reshape(N, C * ph * pw).contiguous())

        def bwd():
            if out.grad is None:
                return
            g = out.grad.reshape(N, C, ph, pw).permute(0, 2, 3, 1).contiguous()
            check(lib().tn_axpy(dptr(t.g()), dptr(g), 1.0, g.numel(), stream_ptr()))
        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, dfeats):
        self.out.grad = dfeats.contiguous().float()
        for fn in reversed(self.tape):
            fn()
        self.tape = []
        return None
