"""Training graph of the captioner: forward with saved activations + backward through
NMTModel.forward = encode -> init_state_from_encoder -> decode_seq -> tgt_proj  (reference train_gnmt.py:330-334,
models/captioning/gnmt.py:136-160,224-304,345-404; gluonnlp NMTModel / DotProductAttentionCell).

Python only sequences entry points of libtennis_b200.so (csrc/tn_seq_train.cu: tn_sgemm, tn_rnn_unroll_forward/backward for
whole-sequence cells, tn_rnn_cell_* + tn_attention_* for the attention-coupled decoder layer, tn_embedding_backward,
tn_mul_mask, tn_dropout_mask); torch tensors are the containers (allocation, views, gathers by index).  fp32 throughout, so the logits agree with the inference engines to ~1e-5 and the gradients with autograd
of the CPU oracle to ~1e-4 (tests/test_gpu_gnmt_train.py).

Structure used for the backward pass: only decoder layer 0 is coupled to the attention (its input at step t carries the
attention vector of step t-1), so it is the one cell that is stepped together with the attention; every other cell
(encoder l/r/uni cells, decoder layers >= 1) is a plain unroll whose input projection is one GEMM over all steps.
"""
import ctypes

import torch

from ... import _lib
from ..._lib import check, dptr, lib, stream_ptr

_GATES = {"gru": 3, "lstm": 4}


def _ld(t):
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "need a row-major 2-D view"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def sgemm(A, B, C, ta=False, tb=False, alpha=1.0, beta=0.0):
    """C = alpha * op(A) op(B) + beta * C on 2-D row-major views (row stride = stride(0))."""
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    N = B.shape[0] if tb else B.shape[1]
    assert (B.shape[1] if tb else B.shape[0]) == K and tuple(C.shape) == (M, N), (A.shape, B.shape, C.shape, ta, tb)
    check(lib().tn_sgemm(int(ta), int(tb), M, N, K, alpha, dptr(A), _ld(A), dptr(B), _ld(B), beta, dptr(C), _ld(C),
                         stream_ptr()))
    return C


def _s(t):
    return 0 if t is None else (t.stride(0) if t.dim() == 2 else 0)


def cell_forward(cell, gi, gh, bi, bh, h_prev, c_prev, h_out, h_out2, c_out, save):
    B, H = h_out.shape
    check(lib().tn_rnn_cell_forward(_lib.CELL_GRU if cell == "gru" else _lib.CELL_LSTM, B, H, dptr(gi), _s(gi), dptr(gh), _s(gh),
                                    dptr(bi), dptr(bh), dptr(h_prev), _s(h_prev), dptr(c_prev), _s(c_prev), dptr(h_out),
                                    _s(h_out), dptr(h_out2), _s(h_out2), dptr(c_out), _s(c_out), dptr(save), _s(save),
                                    stream_ptr()))


def cell_backward(cell, t, lens, save, h_prev, c_prev, c_cur, dy, dy2, dh_last, dc_last, dh, dc, dgi, dgh):
    B, H = dh.shape
    check(lib().tn_rnn_cell_backward(_lib.CELL_GRU if cell == "gru" else _lib.CELL_LSTM, B, H, t, dptr(lens), dptr(save),
                                     _s(save), dptr(h_prev), _s(h_prev), dptr(c_prev), _s(c_prev), dptr(c_cur), _s(c_cur),
                                     dptr(dy), _s(dy), dptr(dy2), _s(dy2), dptr(dh_last), dptr(dc_last), dptr(dh), dptr(dc),
                                     dptr(dgi), _s(dgi), dptr(dgh), _s(dgh), stream_ptr()))


def mul_mask(x, mask, lens):
    """y = x * mask, rows (b, t >= lens[b]) zeroed; x (B,T,C) contiguous."""
    B, T, C = x.shape
    y = torch.empty_like(x)
    check(lib().tn_mul_mask(dptr(x), dptr(mask), dptr(lens), dptr(y), B, T, C, stream_ptr()))
    return y


def dropout_mask(shape, p, seed, device):
    m = torch.empty(shape, device=device, dtype=torch.float32)
    check(lib().tn_dropout_mask(dptr(m), m.numel(), float(p), ctypes.c_ulonglong(seed), stream_ptr()))
    return m


def masked_softmax_ce(pred, label, valid_length, head_grad=None):
    """-> loss (B,), dpred (B,T,V) or None (gluonnlp MaskedSoftmaxCELoss and its gradient)."""
    B, T, V = pred.shape
    pred = pred.contiguous()
    label = label.contiguous().float()
    vl = valid_length.contiguous().float()
    loss = torch.empty(B, device=pred.device)
    ws = torch.empty(B * T, device=pred.device)
    dpred = torch.empty_like(pred) if head_grad is not None else None
    hg = None if head_grad is None else head_grad.contiguous().float()
    check(lib().tn_masked_softmax_ce_grad(dptr(pred), dptr(label), dptr(vl), dptr(hg), dptr(loss), dptr(dpred), dptr(ws), B, T, V,
                                          stream_ptr()))
    return loss, dpred


def _reverse_index(lens, T):
    """SequenceReverse(use_sequence_length=True) as a gather index: position s reads len-1-s for s < len, else s."""
    ar = torch.arange(T, device=lens.device).reshape(1, T)
    ln = lens.reshape(-1, 1).long()
    return torch.where(ar < ln, ln - 1 - ar, ar)


def _seq_reverse(x, idx):
    return x.gather(1, idx.unsqueeze(-1).expand(-1, -1, x.shape[2])).contiguous()


class _Unroll(object):
    """One cell unrolled over (B,T,in) with saved activations: cell.unroll(length, inputs, begin_state, valid_length) (A.4)."""

    def __init__(self, cell_type, cell_block, X, h0, c0, lens):
        self.cell, self.blk = cell_type, cell_block
        G = _GATES[cell_type]
        w = cell_block.weights()
        self.Wi, self.Wh, self.bi, self.bh = w["i2h_weight"], w["h2h_weight"], w["i2h_bias"], w["h2h_bias"]
        B, T, D = X.shape
        H = self.Wh.shape[1]
        dev = X.device
        self.X, self.lens, self.B, self.T, self.H, self.G = X, lens, B, T, H, G
        self.GI = torch.empty(B, T, G * H, device=dev)
        sgemm(X.reshape(B * T, D), self.Wi, self.GI.reshape(B * T, G * H), tb=True)
        self.Hb = torch.zeros(B, T + 1, H, device=dev)
        self.Cb = torch.zeros(B, T + 1, H, device=dev) if cell_type == "lstm" else None
        if h0 is not None:
            self.Hb[:, 0] = h0
        if c0 is not None and self.Cb is not None:
            self.Cb[:, 0] = c0
        self.S = torch.empty(B, T, 4 * H, device=dev)
        gh = torch.empty(B, G * H, device=dev)
        check(lib().tn_rnn_unroll_forward(_lib.CELL_GRU if cell_type == "gru" else _lib.CELL_LSTM, B, T, H, dptr(self.GI),
                                          dptr(self.Wh), dptr(self.bi), dptr(self.bh), dptr(self.Hb), dptr(self.Cb), dptr(self.S),
                                          dptr(gh), stream_ptr()))

    def outputs(self):
        """(B,T,H), zero past valid_length (SequenceMask inside unroll)."""
        return mul_mask(self.Hb[:, 1:].contiguous(), None, self.lens)

    def last_states(self):
        """States after step len-1 (SequenceLast)."""
        ar = torch.arange(self.B, device=self.Hb.device)
        ln = self.lens.long()
        h = self.Hb[ar, ln].contiguous()
        return [h] if self.Cb is None else [h, self.Cb[ar, ln].contiguous()]

    def backward(self, dY, dh_last=None, dc_last=None, need_dx=True):
        """dY (B,T,H) w.r.t. outputs(); dh_last/dc_last w.r.t. last_states().  Accumulates parameter gradients; returns
        (dX or None, dh0, dc0)."""
        B, T, H, G = self.B, self.T, self.H, self.G
        dev = dY.device
        dY = dY.contiguous()
        dh = torch.zeros(B, H, device=dev)
        dc = torch.zeros(B, H, device=dev)
        DGI = torch.empty(B, T, G * H, device=dev)
        DGH = torch.empty(B, T, G * H, device=dev) if self.cell == "gru" else DGI
        dWh = torch.zeros_like(self.Wh)
        check(lib().tn_rnn_unroll_backward(_lib.CELL_GRU if self.cell == "gru" else _lib.CELL_LSTM, B, T, H, dptr(self.lens),
                                           dptr(self.S), dptr(self.Hb), dptr(self.Cb), dptr(self.Wh), dptr(dY), dptr(dh_last),
                                           dptr(dc_last), dptr(dh), dptr(dc), dptr(DGI), dptr(DGH), dptr(dWh), stream_ptr()))
        D = self.X.shape[2]
        X2, DGI2, DGH2 = self.X.reshape(B * T, D), DGI.reshape(B * T, G * H), DGH.reshape(B * T, G * H)
        ones = torch.ones(B * T, 1, device=dev)
        dWi = sgemm(DGI2, X2, torch.empty_like(self.Wi), ta=True)
        dbi = sgemm(DGI2, ones, torch.empty(G * H, 1, device=dev), ta=True).reshape(-1)
        dbh = dbi if self.cell == "lstm" else sgemm(DGH2, ones, torch.empty(G * H, 1, device=dev), ta=True).reshape(-1)
        p = self.blk
        p.i2h_weight._accumulate_grad(dWi)
        p.h2h_weight._accumulate_grad(dWh)
        p.i2h_bias._accumulate_grad(dbi)
        p.h2h_bias._accumulate_grad(dbh)
        dX = None
        if need_dx:
            dX = sgemm(DGI2, self.Wi, torch.empty(B * T, D, device=dev)).reshape(B, T, D)
        return dX, dh, (dc if self.cell == "lstm" else None)


class GNMTTrainGraph(object):
    """One recorded forward of NMTModel (features or frozen-CNN sources) and its backward."""

    def __init__(self, model, seed=0):
        self.m = model
        self.seed = seed

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, src, tgt_ids, src_vl, tgt_vl):
        m = self.m
        enc, dec = m.encoder, m.decoder
        cell = enc._cell_type
        if enc._use_residual or dec._use_residual:
            raise NotImplementedError("use_residual=True is never selected by the scripts (gnmt.py:408 default False)")
        dev = src.device
        # a source that carries a backward (frames through a trainable TimeDistributed(CNN), train_gnmt.py:150-170) wants d(loss)/d(src)
        self.need_dsrc = getattr(src, "_tn_node", None) is not None
        src = src.contiguous().float()
        B, Ts, _ = src.shape
        H = enc._hidden_size
        self.B, self.Ts, self.H, self.cell = B, Ts, H, cell
        self.slens = (torch.full((B,), Ts, device=dev, dtype=torch.int32) if src_vl is None
                      else src_vl.to(torch.int32).clamp(max=Ts).contiguous())
        p_enc, p_dec = float(enc._dropout), float(dec._dropout)
        self.ridx = _reverse_index(self.slens, Ts)

        # ---- encoder (gnmt.py:136-160)
        self.enc_layers = []
        x = src
        enc_states = []
        for i, blk in enumerate(enc.rnn_cells):
            layer = {"in": x}
            if i < enc._num_bi_layers:
                for c in (blk.l_cell, blk.r_cell):
                    c.ensure(x.shape[2], dev)
                fw = _Unroll(cell, blk.l_cell, x, None, None, self.slens)
                bw = _Unroll(cell, blk.r_cell, _seq_reverse(x, self.ridx), None, None, self.slens)
                out = torch.cat([fw.outputs(), _seq_reverse(bw.outputs(), self.ridx)], dim=2)
                enc_states.append(bw.last_states())  # "we use the states of the backward RNN" (gnmt.py:146-148)
                layer.update(kind="bi", fw=fw, bw=bw)
            else:
                blk.ensure(x.shape[2], dev)
                un = _Unroll(cell, blk, x, None, None, self.slens)
                out = un.outputs()
                enc_states.append(un.last_states())
                layer.update(kind="uni", un=un)
            layer["mask"] = None
            if p_enc > 0:
                layer["mask"] = dropout_mask(out.shape, p_enc, self.seed * 1000 + i, dev)
                out = mul_mask(out, layer["mask"], None)
            self.enc_layers.append(layer)
            x = out
        mem = mul_mask(x, None, self.slens)  # final SequenceMask (gnmt.py:157-159)
        self.mem = mem

        # ---- decoder layer 0 + attention, step by step (gnmt.py:345-385)
        ids = tgt_ids.contiguous().float()
        self.ids = ids
        Tt = ids.shape[1]
        self.Tt = Tt
        E = m.tgt_embed.weight.shape[1]
        self.E = E
        cells = list(dec.rnn_cells)
        cells[0].ensure(E + H, dev)
        for c in cells[1:]:
            c.ensure(2 * H, dev)
        G = _GATES[cell]
        w0 = cells[0].weights()
        Wq = dec.attention_cell.proj_query.weight.data()
        X0 = torch.zeros(B, Tt, E + H, device=dev)
        X0[:, :, :E] = m.tgt_embed.weight.data()[ids.long()]
        X1 = torch.empty(B, Tt, 2 * H, device=dev)
        H0 = torch.zeros(B, Tt + 1, H, device=dev)
        C0 = torch.zeros(B, Tt + 1, H, device=dev) if cell == "lstm" else None
        H0[:, 0] = enc_states[0][0]
        if C0 is not None:
            C0[:, 0] = enc_states[0][1]
        S0 = torch.empty(B, Tt, 4 * H, device=dev)
        Q = torch.empty(B, Tt, H, device=dev)
        AW = torch.empty(Tt, B, Ts, device=dev)
        gi = torch.empty(B, G * H, device=dev)
        gh = torch.empty(B, G * H, device=dev)
        L = lib()
        for t in range(Tt):
            sgemm(X0[:, t], w0["i2h_weight"], gi, tb=True)
            sgemm(H0[:, t], w0["h2h_weight"], gh, tb=True)
            cell_forward(cell, gi, gh, w0["i2h_bias"], w0["h2h_bias"], H0[:, t], None if C0 is None else C0[:, t], H0[:, t + 1],
                         X1[:, t, :H], None if C0 is None else C0[:, t + 1], S0[:, t])
            sgemm(H0[:, t + 1], Wq, Q[:, t], tb=True)
            ctx2 = X0[:, t + 1, E:] if t + 1 < Tt else None
            check(L.tn_attention_forward(dptr(Q[:, t]), Tt * H, dptr(mem), dptr(self.slens), B, Ts, H, dptr(AW[t]),
                                         dptr(X1[:, t, H:]), Tt * 2 * H, dptr(ctx2), 0 if ctx2 is None else Tt * (E + H),
                                         stream_ptr()))
        self.X0, self.X1, self.H0, self.C0, self.S0, self.Q, self.AW = X0, X1, H0, C0, S0, Q, AW

        # ---- decoder layers >= 1: plain unrolls over [previous output, attention vector] (gnmt.py:387-396)
        self.tlens = (torch.full((B,), Tt, device=dev, dtype=torch.int32) if tgt_vl is None
                      else tgt_vl.to(torch.int32).clamp(max=Tt).contiguous())
        full = torch.full((B,), Tt, device=dev, dtype=torch.int32)
        self.dec_layers = []
        xin = X1
        out = None
        for i in range(1, len(cells)):
            st = enc_states[i]
            un = _Unroll(cell, cells[i], xin, st[0], st[1] if cell == "lstm" else None, full)
            out = un.Hb[:, 1:].contiguous()
            mask = None
            if p_dec > 0:
                mask = dropout_mask(out.shape, p_dec, self.seed * 1000 + 500 + i, dev)
                out = mul_mask(out, mask, None)
            self.dec_layers.append({"un": un, "mask": mask})
            if i + 1 < len(cells):
                xin = torch.cat([out, X1[:, :, H:]], dim=2)
        if out is None:  # single-layer decoder: the output is layer 0's
            out = H0[:, 1:].contiguous()
        out = mul_mask(out, None, self.tlens)  # decode_seq SequenceMask (gnmt.py:298-301)
        self.out = out
        V = m.tgt_proj.weight.shape[0]
        from ... import ops
        logits = ops.dense(out.reshape(B * Tt, H), m.tgt_proj.weight.data(), m.tgt_proj.bias.data()).reshape(B, Tt, V)
        return logits

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, dlogits):
        from ... import ops
        m = self.m
        enc, dec = m.encoder, m.decoder
        cell, B, Ts, Tt, H, E = self.cell, self.B, self.Ts, self.Tt, self.H, self.E
        dev = dlogits.device
        G = _GATES[cell]
        V = dlogits.shape[2]
        dout, dWp, dbp = ops.dense_backward(self.out.reshape(B * Tt, H), m.tgt_proj.weight.data(),
                                            dlogits.reshape(B * Tt, V).contiguous())
        m.tgt_proj.weight._accumulate_grad(dWp)
        m.tgt_proj.bias._accumulate_grad(dbp)
        d = mul_mask(dout.reshape(B, Tt, H).contiguous(), None, self.tlens)
        cells = list(dec.rnn_cells)
        dX1 = torch.zeros(B, Tt, 2 * H, device=dev)
        dinit = [None] * len(cells)  # gradients w.r.t. the decoder's initial states = the encoder's last states
        dH0_from_top = None
        if self.dec_layers:
            for li in range(len(self.dec_layers) - 1, -1, -1):
                layer = self.dec_layers[li]
                if layer["mask"] is not None:
                    d = mul_mask(d, layer["mask"], None)
                dX, dh0, dc0 = layer["un"].backward(d)
                dinit[li + 1] = (dh0, dc0)
                if li == 0:
                    dX1 += dX
                else:
                    dX1[:, :, H:] += dX[:, :, H:]
                    d = dX[:, :, :H].contiguous()
        else:
            dH0_from_top = d

        # ---- decoder layer 0 + attention, reverse time
        w0 = cells[0].weights()
        Wq = dec.attention_cell.proj_query.weight.data()
        X0, X1, H0, C0, S0, Q, AW = self.X0, self.X1, self.H0, self.C0, self.S0, self.Q, self.AW
        dX0 = torch.zeros(B, Tt, E + H, device=dev)
        dmem = torch.zeros(B, Ts, H, device=dev)
        dh = torch.zeros(B, H, device=dev)
        dc = torch.zeros(B, H, device=dev)
        dq = torch.empty(B, H, device=dev)
        dhq = torch.empty(B, H, device=dev)
        DGI = torch.empty(B, Tt, G * H, device=dev)
        DGH = torch.empty(B, Tt, G * H, device=dev) if cell == "gru" else DGI
        dWh = torch.zeros_like(w0["h2h_weight"])
        dWq = torch.zeros_like(Wq)
        L = lib()
        for t in range(Tt - 1, -1, -1):
            d2 = dX0[:, t + 1, E:] if t + 1 < Tt else None
            check(L.tn_attention_backward(dptr(Q[:, t]), Tt * H, dptr(self.mem), dptr(self.slens), B, Ts, H, dptr(AW[t]),
                                          dptr(dX1[:, t, H:]), Tt * 2 * H, dptr(d2), 0 if d2 is None else Tt * (E + H), dptr(dq), H,
                                          dptr(dmem), stream_ptr()))
            sgemm(dq, Wq, dhq)                                   # q = h Wq^T  ->  dh += dq Wq
            sgemm(dq, H0[:, t + 1], dWq, ta=True, beta=1.0)      # dWq += dq^T h
            if dH0_from_top is not None:
                check(L.tn_axpy(dptr(dhq), dptr(dH0_from_top[:, t].contiguous()), 1.0, B * H, stream_ptr()))
            cell_backward(cell, t, None, S0[:, t], H0[:, t], None if C0 is None else C0[:, t], None if C0 is None else C0[:, t + 1],
                          dX1[:, t, :H], dhq, None, None, dh, dc, DGI[:, t], DGH[:, t] if cell == "gru" else None)
            sgemm(DGH[:, t], w0["h2h_weight"], dh, beta=1.0)
            sgemm(DGH[:, t], H0[:, t], dWh, ta=True, beta=1.0)
            sgemm(DGI[:, t], w0["i2h_weight"], dX0[:, t])
        ones = torch.ones(B * Tt, 1, device=dev)
        DGI2, DGH2 = DGI.reshape(B * Tt, G * H), DGH.reshape(B * Tt, G * H)
        dWi = sgemm(DGI2, X0.reshape(B * Tt, E + H), torch.empty_like(w0["i2h_weight"]), ta=True)
        dbi = sgemm(DGI2, ones, torch.empty(G * H, 1, device=dev), ta=True).reshape(-1)
        dbh = dbi if cell == "lstm" else sgemm(DGH2, ones, torch.empty(G * H, 1, device=dev), ta=True).reshape(-1)
        cells[0].i2h_weight._accumulate_grad(dWi)
        cells[0].h2h_weight._accumulate_grad(dWh)
        cells[0].i2h_bias._accumulate_grad(dbi)
        cells[0].h2h_bias._accumulate_grad(dbh)
        dec.attention_cell.proj_query.weight._accumulate_grad(dWq)
        dinit[0] = (dh, dc if cell == "lstm" else None)
        if m.tgt_embed.weight.grad_req != 'null':
            dE = torch.zeros_like(m.tgt_embed.weight.data())
            check(L.tn_embedding_backward(dptr(self.ids), dptr(dX0), E + H, dptr(dE), B * Tt, E, dE.shape[0], stream_ptr()))
            m.tgt_embed.weight._accumulate_grad(dE)

        # ---- encoder, top layer first
        d = dmem  # rows past valid_length are zero already (the attention never reads them)
        for i in range(len(self.enc_layers) - 1, -1, -1):
            layer = self.enc_layers[i]
            if layer["mask"] is not None:
                d = mul_mask(d, layer["mask"], None)
            dh_last, dc_last = dinit[i] if i < len(dinit) and dinit[i] is not None else (None, None)
            need_dx = i > 0 or self.need_dsrc
            if layer["kind"] == "uni":
                d, _, _ = layer["un"].backward(d, dh_last, dc_last, need_dx=need_dx)
            else:
                dl = d[:, :, :H].contiguous()
                dr = _seq_reverse(d[:, :, H:].contiguous(), self.ridx)
                dxl, _, _ = layer["fw"].backward(dl, None, None, need_dx=need_dx)
                dxr, _, _ = layer["bw"].backward(dr, dh_last, dc_last, need_dx=need_dx)  # decoder starts from the backward cell
                if need_dx:
                    d = dxl
                    check(L.tn_axpy(dptr(d), dptr(_seq_reverse(dxr, self.ridx)), 1.0, d.numel(), stream_ptr()))
        return d if self.need_dsrc else None  # (B, T_src, D): gradient of the source features, zero past valid_length
