"""Thin, torch-tensor-facing wrappers over the C ABI.  Tensors are containers only: every op below hands raw
device pointers + the current CUDA stream to libtennis_b200.so.  No op has a PyTorch/CPU fallback."""
import ctypes
import os
from ctypes import c_void_p

import torch

from . import _lib
from ._lib import check, dptr, lib, stream_ptr


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.TennisB200Error("tennis_b200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)


def _host_f32(t):
    return None if t is None else t.detach().to("cpu", torch.float32).contiguous()


def _workspace(nbytes, device):
    # torch's caching allocator returns >=512-byte aligned blocks; over-allocate and align to 1024 ourselves
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf, buf.data_ptr() + off


class Conv:
    """One fused convolution (tn_conv_*): [BN+ReLU prologue] -> conv -> [scale/shift(+ReLU), residual] epilogue."""

    def __init__(self, weight, mode=_lib.MODE_CONV, pro_scale=None, pro_shift=None, epi_scale=None, epi_shift=None,
                 device=0):
        w = _host_f32(weight)
        self.Cout, self.Cin, self.R, self.S = w.shape
        self.mode = mode
        keep = [w] + [_host_f32(t) for t in (pro_scale, pro_shift, epi_scale, epi_shift)]
        self._h = c_void_p()
        check(lib().tn_conv_create(ctypes.byref(self._h), device, dptr(keep[0]), self.Cout, self.Cin, self.R, self.S,
                                   mode, dptr(keep[1]), dptr(keep[2]), dptr(keep[3]), dptr(keep[4])))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().tn_conv_destroy(self._h)
                self._h = c_void_p()
        except Exception:  # interpreter shutdown
            pass

    def out_hw(self, H, W, stride, pad):
        if self.mode == _lib.MODE_POOL2:
            return H // 2, W // 2
        return (H + 2 * pad - self.R) // stride + 1, (W + 2 * pad - self.S) // stride + 1

    def __call__(self, x, stride=1, pad=0, pro_relu=True, epi_relu=False, out=None, out_coff=0, out_fp32=False,
                 residual=None):
        """x: (n,H,W,C) bf16 NHWC cuda tensor; returns / fills `out` (n,Ho,Wo,Cs)."""
        _require_cuda(x, out, residual)
        n, H, W, cs = x.shape
        Ho, Wo = self.out_hw(H, W, stride, pad)
        if out is None:
            out = torch.empty((n, Ho, Wo, self.Cout), dtype=torch.float32 if out_fp32 else torch.bfloat16,
                              device=x.device)
        check(lib().tn_conv_forward(self._h, dptr(x), cs, n, H, W, stride, pad, int(pro_relu), int(epi_relu), dptr(out),
                                    out.shape[3], out_coff, int(out.dtype == torch.float32), dptr(residual),
                                    0 if residual is None else residual.shape[3], stream_ptr()))
        return out


def frames_to_nhwc4(frames):
    _require_cuda(frames)
    n, c, h, w = frames.shape
    assert c == 3 and frames.dtype == torch.float32
    out = torch.empty((n, h, w, 4), dtype=torch.bfloat16, device=frames.device)
    check(lib().tn_frames_to_nhwc4(dptr(frames.contiguous()), dptr(out), n, h, w, stream_ptr()))
    return out


class Backbone:
    """DenseNet-121 / ResNet-18 v2 `.features` (tn_backbone_*)."""

    ARCH = {"densenet121": _lib.ARCH_DENSENET121, "resnet18_v2": _lib.ARCH_RESNET18_V2}

    PRECISION = {"bf16": 0, "split_bf16": 1}  # TN_PRECISION_* of include/tennis_b200.h

    def __init__(self, arch, flat_params, device=0, precision=None):
        """precision: 'bf16' (speed path; logits within ~2e-2 of the fp32 reference) or 'split_bf16' (fp32-grade, three tensor-core
        products per contraction on (hi, lo) bf16 planes; DenseNet-121).  Default: $TN_PRECISION or 'bf16'."""
        self.arch_name = arch.lower()
        self.arch = self.ARCH[self.arch_name]
        self.device = device
        flat = _host_f32(flat_params)
        self._h = c_void_p()
        check(lib().tn_backbone_create(ctypes.byref(self._h), self.arch, device, dptr(flat), flat.numel()))
        self._ws = None
        self._ws_key = None
        self.precision = "bf16"
        self.set_precision(precision or os.environ.get("TN_PRECISION", "bf16"))

    def set_precision(self, precision):
        if precision not in self.PRECISION:
            raise ValueError("precision must be one of %s" % sorted(self.PRECISION))
        if precision != self.precision:
            check(lib().tn_backbone_set_precision(self._h, self.PRECISION[precision]))
            self.precision = precision
            self._ws, self._ws_key = None, None  # the split path needs a larger workspace

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().tn_backbone_destroy(self._h)
                self._h = c_void_p()
        except Exception:  # interpreter shutdown
            pass

    def feature_dim(self, h, w):
        return lib().tn_backbone_feature_dim(self.arch, h, w)

    def workspace_bytes(self, n, h, w):
        return lib().tn_backbone_workspace_bytes(self._h, n, h, w)

    def _get_ws(self, n, h, w, device):
        need = self.workspace_bytes(n, h, w)
        if need == 0:
            raise _lib.TennisB200Error("unsupported input %dx%d for %s" % (h, w, self.arch_name))
        if self._ws is None or self._ws_key[0] < need or self._ws_key[1] != device:
            self._ws = None
            self._ws = _workspace(need, device)
            self._ws_key = (need, device)
        return self._ws[1], self._ws_key[0]

    def __call__(self, frames, want_bf16=False):
        """frames: (n,3,H,W) fp32 normalised NCHW, or (n,H,W,3) uint8 NHWC -> (n,D) fp32 [, (n,D) bf16]."""
        _require_cuda(frames)
        frames = frames.contiguous()
        if frames.dtype == torch.uint8:
            n, h, w, c = frames.shape
            dt = _lib.FRAMES_U8_NHWC
        else:
            n, c, h, w = frames.shape
            dt = _lib.FRAMES_F32_NCHW
            assert frames.dtype == torch.float32
        assert c == 3
        D = self.feature_dim(h, w)
        feats = torch.empty((n, D), dtype=torch.float32, device=frames.device)
        fb = torch.empty((n, D), dtype=torch.bfloat16, device=frames.device) if want_bf16 else None
        if n:
            ws_ptr, ws_bytes = self._get_ws(n, h, w, frames.device)
            check(lib().tn_backbone_forward(self._h, dptr(frames), dt, n, h, w, dptr(feats), dptr(fb), c_void_p(ws_ptr),
                                            ws_bytes, stream_ptr()))
        return (feats, fb) if want_bf16 else feats


def dense(x, weight, bias=None):
    _require_cuda(x, weight, bias)
    x2 = x.reshape(x.shape[0], -1).contiguous().float()
    out_dim, in_dim = weight.shape
    assert x2.shape[1] == in_dim
    y = torch.empty((x2.shape[0], out_dim), dtype=torch.float32, device=x.device)
    check(lib().tn_dense_forward(dptr(x2), dptr(weight.contiguous()), dptr(None if bias is None else bias.contiguous()),
                                 dptr(y), x2.shape[0], in_dim, out_dim, stream_ptr()))
    return y


def temporal_pool(x, pool="max"):
    _require_cuda(x)
    B, T, D = x.shape
    x = x.contiguous().float()
    y = torch.empty((B, D), dtype=torch.float32, device=x.device)
    check(lib().tn_temporal_pool(dptr(x), dptr(y), B, T, D, _lib.POOL_MEAN if pool == "mean" else _lib.POOL_MAX,
                                 stream_ptr()))
    return y


class BiRNN:
    """Fused (bi)directional GRU/LSTM layer (tn_birnn_*).  params: dict with Gluon names l0_*/r0_*."""

    def __init__(self, cell, D, H, params, bidirectional=True, device=0, precise=False):
        self.cell, self.D, self.H = cell, D, H
        self.ndir = 2 if bidirectional else 1
        dirs = ["l0", "r0"][: self.ndir]
        keep = {}
        arrs = []
        for suffix in ("_i2h_weight", "_h2h_weight", "_i2h_bias", "_h2h_bias"):
            arr = (c_void_p * 2)()
            for i, d in enumerate(dirs):
                t = _host_f32(params[d + suffix])
                keep[d + suffix] = t
                arr[i] = t.data_ptr()
            arrs.append(arr)
        self._h = c_void_p()
        check(lib().tn_birnn_create(ctypes.byref(self._h), device, _lib.CELL_GRU if cell == "gru" else _lib.CELL_LSTM, D,
                                    H, self.ndir, arrs[0], arrs[1], arrs[2], arrs[3]))
        self._precise = False
        self.set_precise(precise)
        self._ws = None
        self._ws_bytes = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().tn_birnn_destroy(self._h)
                self._h = c_void_p()
        except Exception:  # interpreter shutdown
            pass

    def set_precise(self, on):
        if bool(on) != getattr(self, "_precise", False):
            check(lib().tn_birnn_set_precise(self._h, int(bool(on))))
            self._precise = bool(on)

    def update_weights(self, params):
        """Refresh the packed weights from fp32 CUDA tensors (same keys as the constructor) without leaving the device."""
        dirs = ["l0", "r0"][: self.ndir]
        arrs, keep = [], []
        for suffix in ("_i2h_weight", "_h2h_weight", "_i2h_bias", "_h2h_bias"):
            arr = (c_void_p * 2)()
            for i, d in enumerate(dirs):
                t = params[d + suffix]
                _require_cuda(t)
                t = t.contiguous().float()
                keep.append(t)
                arr[i] = t.data_ptr()
            arrs.append(arr)
        check(lib().tn_birnn_update_weights(self._h, arrs[0], arrs[1], arrs[2], arrs[3], stream_ptr()))

    def __call__(self, x, valid_len=None, want_y=True, want_max=False, want_state=False):
        """x: (B,T,D) fp32 or bf16 cuda.  Returns dict(y, ymax, h, c) with the requested entries."""
        _require_cuda(x, valid_len)
        B, T, D = x.shape
        assert D == self.D
        x = x.contiguous()
        dev = x.device
        y = torch.empty((B, T, self.ndir * self.H), dtype=torch.float32, device=dev) if want_y else None
        ymax = torch.empty((B, self.ndir * self.H), dtype=torch.float32, device=dev) if want_max else None
        h = torch.empty((self.ndir, B, self.H), dtype=torch.float32, device=dev) if want_state else None
        c = torch.empty((self.ndir, B, self.H), dtype=torch.float32, device=dev) if (want_state and self.cell == "lstm") else None
        vl = None if valid_len is None else valid_len.to(torch.int32).contiguous()
        if B and T:
            need = lib().tn_birnn_workspace_bytes(self._h, B, T)
            if self._ws is None or self._ws_bytes < need:
                self._ws = None
                self._ws = _workspace(need, dev)
                self._ws_bytes = need
            check(lib().tn_birnn_forward(self._h, dptr(x), int(x.dtype == torch.bfloat16), dptr(vl), B, T, dptr(y),
                                         dptr(ymax), dptr(h), dptr(c), c_void_p(self._ws[1]), self._ws_bytes,
                                         stream_ptr()))
        return {"y": y, "ymax": ymax, "h": h, "c": c}

    def forward_train(self, x):
        """Forward keeping the activations the backward needs.  -> dict(y, ymax, gx, cseq, x)."""
        _require_cuda(x)
        B, T, D = x.shape
        x = x.contiguous()
        dev = x.device
        G = 3 if self.cell == "gru" else 4
        y = torch.empty((B, T, self.ndir * self.H), dtype=torch.float32, device=dev)
        ymax = torch.empty((B, self.ndir * self.H), dtype=torch.float32, device=dev)
        gx = torch.empty((B * T, self.ndir * G * self.H), dtype=torch.float32, device=dev)
        cseq = torch.empty_like(y) if self.cell == "lstm" else None
        need = lib().tn_birnn_workspace_bytes(self._h, B, T)
        if self._ws is None or self._ws_bytes < need:
            self._ws = None
            self._ws = _workspace(need, dev)
            self._ws_bytes = need
        check(lib().tn_birnn_forward_train(self._h, dptr(x), int(x.dtype == torch.bfloat16), B, T, dptr(y), dptr(ymax), dptr(gx),
                                           dptr(cseq), c_void_p(self._ws[1]), self._ws_bytes, stream_ptr()))
        return {"y": y, "ymax": ymax, "gx": gx, "cseq": cseq, "x": x}

    def backward(self, saved, d_ymax=None, dy=None, want_dx=False, weights=None):
        """-> dict of gradients keyed like the Gluon parameters (l0_i2h_weight, ..., r0_h2h_bias); with want_dx also
        "dx" (B,T,D) = sum_dir d(gates_x) W_i2h, from the gate gradients tn_birnn_backward leaves at the head of its workspace
        (`weights`: the fp32 CUDA parameters)."""
        x = saved["x"]
        B, T, D = x.shape
        G = 3 if self.cell == "gru" else 4
        GH = G * self.H
        dev = x.device
        dWih = torch.empty((self.ndir * GH, D), dtype=torch.float32, device=dev)
        dWhh = torch.empty((self.ndir * GH, self.H), dtype=torch.float32, device=dev)
        dbih = torch.empty((self.ndir * GH,), dtype=torch.float32, device=dev)
        dbhh = torch.empty((self.ndir * GH,), dtype=torch.float32, device=dev)
        need = lib().tn_birnn_backward_workspace_bytes(self._h, B, T)
        buf, ptr = _workspace(need, dev)
        check(lib().tn_birnn_backward(self._h, dptr(x), int(x.dtype == torch.bfloat16), B, T, dptr(saved["gx"]), dptr(saved["y"]),
                                      dptr(saved["cseq"]), dptr(saved["ymax"]),
                                      dptr(None if d_ymax is None else d_ymax.contiguous().float()),
                                      dptr(None if dy is None else dy.contiguous().float()), dptr(dWih), dptr(dWhh), dptr(dbih),
                                      dptr(dbhh), c_void_p(ptr), need, stream_ptr()))
        out = {}
        if want_dx:
            from .models.captioning.train_graph import sgemm
            off = ptr - buf.data_ptr()
            dgx = buf[off: off + B * T * self.ndir * GH * 4].view(torch.float32).reshape(B * T, self.ndir * GH)
            dx = torch.empty((B * T, D), dtype=torch.float32, device=dev)
            for i, d in enumerate(["l0", "r0"][: self.ndir]):
                sgemm(dgx[:, i * GH:(i + 1) * GH], weights[d + "_i2h_weight"].float(), dx, beta=0.0 if i == 0 else 1.0)
            out["dx"] = dx.reshape(B, T, D)
        for i, d in enumerate(["l0", "r0"][: self.ndir]):
            out[d + "_i2h_weight"] = dWih[i * GH:(i + 1) * GH]
            out[d + "_h2h_weight"] = dWhh[i * GH:(i + 1) * GH]
            out[d + "_i2h_bias"] = dbih[i * GH:(i + 1) * GH]
            out[d + "_h2h_bias"] = dbhh[i * GH:(i + 1) * GH]
        return out


class GNMTDecoderEngine:
    """GNMT decoder + target embedding/projection + beam search on the device (tn_gnmt_*)."""

    def __init__(self, cell, H, E, V, layer_params, query_weight, embed_weight, proj_weight, proj_bias, use_residual=False,
                 device=0):
        """layer_params: list (per decoder layer) of dicts with i2h_weight/h2h_weight/i2h_bias/h2h_bias."""
        self.cell, self.H, self.E, self.V, self.L = cell, H, E, V, len(layer_params)
        keep = []
        arrs = []
        for name in ("i2h_weight", "h2h_weight", "i2h_bias", "h2h_bias"):
            arr = (c_void_p * self.L)()
            for i, lp in enumerate(layer_params):
                t = _host_f32(lp[name])
                keep.append(t)
                arr[i] = t.data_ptr()
            arrs.append(arr)
        others = [_host_f32(t) for t in (query_weight, embed_weight, proj_weight, proj_bias)]
        self._h = c_void_p()
        check(lib().tn_gnmt_create(ctypes.byref(self._h), device, _lib.CELL_GRU if cell == "gru" else _lib.CELL_LSTM, H, E,
                                   V, self.L, int(use_residual), arrs[0], arrs[1], arrs[2], arrs[3], dptr(others[0]),
                                   dptr(others[1]), dptr(others[2]), dptr(others[3])))
        self._ws = None
        self._ws_bytes = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().tn_gnmt_destroy(self._h)
                self._h = c_void_p()
        except Exception:
            pass

    def _workspace(self, rows, max_len, device):
        need = lib().tn_gnmt_workspace_bytes(self._h, rows, max_len)
        if self._ws is None or self._ws_bytes < need:
            self._ws = None
            self._ws = _workspace(need, device)
            self._ws_bytes = need
        return c_void_p(self._ws[1]), self._ws_bytes

    @staticmethod
    def _i32(t):
        return None if t is None else t.to(torch.int32).contiguous()

    def decode_step(self, step_ids, h, c, att, mem, src_len, rows_per_mem=1):
        """step_ids (R,) float; h/c (L,R,H); att (R,H); mem (R/rows_per_mem,T,H) -> logits (R,V), h', c', att'."""
        _require_cuda(step_ids, h, c, att, mem, src_len)
        R, T = step_ids.shape[0], mem.shape[1]
        dev = mem.device
        logits = torch.empty((R, self.V), dtype=torch.float32, device=dev)
        h2, att2 = torch.empty_like(h), torch.empty_like(att)
        c2 = torch.empty_like(c) if c is not None else None
        ws, wsb = self._workspace(R, 1, dev)
        sl = self._i32(src_len)
        check(lib().tn_gnmt_decode_step(self._h, dptr(step_ids.float().contiguous()), dptr(h.contiguous()),
                                        dptr(None if c is None else c.contiguous()), dptr(att.contiguous()),
                                        dptr(mem.contiguous()), dptr(sl), rows_per_mem, R, T, dptr(logits), dptr(h2), dptr(c2),
                                        dptr(att2), ws, wsb, stream_ptr()))
        return logits, h2, c2, att2

    def decoder_step(self, step_emb, h, c, att, mem, src_len, rows_per_mem=1):
        """The decoder BLOCK's step (gnmt.py:306-404): step_emb (R,E) embedded inputs -> rnn_out (R,H), h', c', att'."""
        _require_cuda(step_emb, h, c, att, mem, src_len)
        R, T = step_emb.shape[0], mem.shape[1]
        out = torch.empty((R, self.H), dtype=torch.float32, device=mem.device)
        h2, att2 = torch.empty_like(h), torch.empty_like(att)
        c2 = torch.empty_like(c) if c is not None else None
        ws, wsb = self._workspace(R, 1, mem.device)
        sl = self._i32(src_len)
        check(lib().tn_gnmt_decoder_step(self._h, dptr(step_emb.float().contiguous()), dptr(h.contiguous()),
                                         dptr(None if c is None else c.contiguous()), dptr(att.contiguous()),
                                         dptr(mem.contiguous()), dptr(sl), rows_per_mem, R, T, dptr(out), dptr(h2), dptr(c2),
                                         dptr(att2), ws, wsb, stream_ptr()))
        return out, h2, c2, att2

    def decode_seq(self, tgt_ids, tgt_valid_len, h0, c0, mem, src_len):
        """tgt_ids (B,T_tgt) float; h0/c0 (L,B,H) -> logits (B,T_tgt,V)."""
        _require_cuda(tgt_ids, h0, c0, mem, src_len, tgt_valid_len)
        B, Tt = tgt_ids.shape
        T = mem.shape[1]
        logits = torch.empty((B, Tt, self.V), dtype=torch.float32, device=mem.device)
        ws, wsb = self._workspace(B, Tt, mem.device)
        tv, sl = self._i32(tgt_valid_len), self._i32(src_len)
        check(lib().tn_gnmt_decode_seq(self._h, dptr(tgt_ids.float().contiguous()), dptr(tv), dptr(h0.contiguous()),
                                       dptr(None if c0 is None else c0.contiguous()), dptr(mem.contiguous()), dptr(sl), B, T, Tt,
                                       dptr(logits), ws, wsb, stream_ptr()))
        return logits

    def beam_search(self, mem, src_len, h0, c0, beam, max_len, alpha, K, bos, eos):
        """-> samples (B,beam,L) int32, scores (B,beam), valid_length (B,beam) int32 (BeamSearchSampler's result)."""
        _require_cuda(mem, src_len, h0, c0)
        B, T = mem.shape[0], mem.shape[1]
        dev = mem.device
        samples = torch.empty((B, beam, max_len + 2), dtype=torch.int32, device=dev)
        scores = torch.empty((B, beam), dtype=torch.float32, device=dev)
        vlen = torch.empty((B, beam), dtype=torch.int32, device=dev)
        out_len = ctypes.c_int(0)
        ws, wsb = self._workspace(B * beam, max_len, dev)
        sl = self._i32(src_len)
        check(lib().tn_gnmt_beam_search(self._h, dptr(mem.contiguous()), dptr(sl), dptr(h0.contiguous()),
                                        dptr(None if c0 is None else c0.contiguous()), B, T, beam, max_len, float(alpha),
                                        float(K), bos, eos, dptr(samples), dptr(scores), dptr(vlen), ctypes.byref(out_len), ws,
                                        wsb, stream_ptr()))
        return samples[:, :, : out_len.value].contiguous(), scores, vlen


def softmax_ce(logits, labels, want_grad=False):
    """-> (loss (B,), dlogits (B,C) or None)."""
    _require_cuda(logits, labels)
    B, C = logits.shape
    logits = logits.contiguous().float()
    lab = labels.to(torch.int32).contiguous()
    loss = torch.empty((B,), dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    check(lib().tn_softmax_ce(dptr(logits), dptr(lab), dptr(loss), dptr(dl), B, C, stream_ptr()))
    return loss, dl


def dense_backward(x, weight, dy):
    """-> (dx (R,in), dW (out,in), db (out))."""
    _require_cuda(x, weight, dy)
    R, in_dim = x.shape
    out_dim = weight.shape[0]
    dy = dy.contiguous().float()
    dx = torch.empty((R, in_dim), dtype=torch.float32, device=x.device)
    dw = torch.empty((out_dim, in_dim), dtype=torch.float32, device=x.device)
    db = torch.empty((out_dim,), dtype=torch.float32, device=x.device)
    check(lib().tn_dense_backward(dptr(x), dptr(weight.contiguous()), dptr(dy), dptr(dx), dptr(dw), dptr(db), R, in_dim, out_dim,
                                  stream_ptr()))
    return dx, dw, db


def sgd_mom_update(w, g, mom, lr, momentum, wd, rescale):
    _require_cuda(w, g, mom)
    assert w.is_contiguous() and mom.is_contiguous()
    check(lib().tn_sgd_mom_update(dptr(w), dptr(g.contiguous()), dptr(mom), w.numel(), lr, momentum, wd, rescale, stream_ptr()))


def adam_update(w, g, mean, var, lr, b1, b2, eps, wd, rescale, t):
    _require_cuda(w, g, mean, var)
    assert w.is_contiguous()
    check(lib().tn_adam_update(dptr(w), dptr(g.contiguous()), dptr(mean), dptr(var), w.numel(), lr, b1, b2, eps, wd, rescale, t,
                               stream_ptr()))


def masked_softmax_ce(pred, label, valid_len):
    """gluonnlp MaskedSoftmaxCELoss forward: pred (B,T,V), label (B,T), valid_len (B) -> loss (B)."""
    _require_cuda(pred, label, valid_len)
    B, T, V = pred.shape
    loss = torch.empty((B,), dtype=torch.float32, device=pred.device)
    check(lib().tn_masked_softmax_ce(dptr(pred.contiguous().float()), dptr(label.contiguous().float()),
                                     dptr(valid_len.contiguous().float()), dptr(loss), B, T, V, stream_ptr()))
    return loss
