"""CPU oracle for the captioning half of the hot path (TEST INFRASTRUCTURE — never imported by the product).

PARITY UNPINNED (see oracle/vision.py): GluonNLP/MXNet are not installable here and the reference has no golden
vectors; this restates, in plain PyTorch fp32 on the CPU,

  * models/captioning/gnmt.py:136-160   GNMTEncoder.forward       (cell.unroll with valid_length, A.4)
  * models/captioning/gnmt.py:224-252   GNMTDecoder.init_state_from_encoder
  * models/captioning/gnmt.py:345-404   GNMTDecoder.hybrid_forward (one step, scaled-Luong attention A.5)
  * models/captioning/gnmt.py:254-304   GNMTDecoder.decode_seq
  * [UPSTREAM] gluonnlp NMTModel glue (A.6), BeamSearchScorer / BeamSearchSampler (A.7), MaskedSoftmaxCELoss (A.8)
  * utils/translation.py:51-82          BeamSearchTranslator
"""
import math

import torch

from tennis_b200.synthetic import (cell_param_shapes, gnmt_param_shapes, synthetic_gnmt_params,  # noqa: F401
                                   synthetic_sources)

from .vision import gru_cell, lstm_cell

NEG = -1e18


# ----------------------------------------------------------------------------------------------------------------------
def _cell_step(cell, p, prefix, x, state):
    w = [p[prefix + n] for n in ("i2h_weight", "h2h_weight", "i2h_bias", "h2h_bias")]
    if cell == "gru":
        h = gru_cell(x, state[0], *w)
        return h, [h]
    h, c = lstm_cell(x, state[0], state[1], *w)
    return h, [h, c]


def _zero_state(cell, B, H, ref):
    return [ref.new_zeros(B, H)] if cell == "gru" else [ref.new_zeros(B, H), ref.new_zeros(B, H)]


def _sequence_reverse(x, vl):
    """mx.nd.SequenceReverse(use_sequence_length=True) on (B,T,C): reverse the first vl[b] steps, padding stays."""
    out = x.clone()
    for b in range(x.shape[0]):
        n = int(vl[b])
        out[b, :n] = x[b, :n].flip(0)
    return out


def _unroll(cell, p, prefix, x, vl, H):
    """RecurrentCell.unroll(valid_length=...) (A.4): NO masking of the recurrence; states = SequenceLast; outputs masked."""
    B, T, _ = x.shape
    state = _zero_state(cell, B, H, x)
    outs, all_states = [], []
    for t in range(T):
        o, state = _cell_step(cell, p, prefix, x[:, t], state)
        outs.append(o)
        all_states.append(state)
    out = torch.stack(outs, 1)
    if vl is not None:
        idx = (vl.long() - 1).clamp(min=0)
        last = [torch.stack([all_states[int(idx[b])][k][b] for b in range(B)]) for k in range(len(state))]
        mask = (torch.arange(T).reshape(1, T) < vl.reshape(B, 1)).to(out.dtype).unsqueeze(-1)
        return out * mask, last
    return out, state


def encoder_forward(p, x, vl, cell="lstm", H=128, num_layers=2, num_bi_layers=1, use_residual=False):
    """gnmt.py:136-160.  Returns (outputs (B,T,H), states: list per layer of [h] or [h,c])."""
    inputs = x
    new_states = []
    outputs = x
    for i in range(num_layers):
        pre = "encoder.rnn_cells.%d." % i
        if i < num_bi_layers:
            l_out, _ = _unroll(cell, p, pre + "l_cell.", inputs, vl, H)
            rev_in = inputs.flip(1) if vl is None else _sequence_reverse(inputs, vl)
            r_out, r_states = _unroll(cell, p, pre + "r_cell.", rev_in, vl, H)
            r_out = r_out.flip(1) if vl is None else _sequence_reverse(r_out, vl)
            outputs = torch.cat([l_out, r_out], dim=2)
            new_states.append(r_states)  # "we use the states of the backward RNN" (gnmt.py:146-148)
        else:
            outputs, st = _unroll(cell, p, pre, inputs, vl, H)
            new_states.append(st)
        if use_residual and i > num_bi_layers:  # strict '>' (Appendix C #13)
            outputs = outputs + inputs
        inputs = outputs
    if vl is not None:
        T = outputs.shape[1]
        outputs = outputs * (torch.arange(T).reshape(1, T) < vl.reshape(-1, 1)).to(outputs.dtype).unsqueeze(-1)
    return outputs, new_states


def init_state_from_encoder(mem_value, rnn_states, vl):
    """gnmt.py:224-252."""
    B, T, Hm = mem_value.shape
    attention_vec = mem_value.new_zeros(B, Hm)
    masks = None if vl is None else (torch.arange(T).reshape(1, T) < vl.reshape(B, 1)).to(mem_value.dtype)
    return [rnn_states, attention_vec, mem_value, masks]


def attention(p, query, mem_value, masks, H):
    """DotProductAttentionCell(units=H, scaled=True, use_bias=False, luong_style=True) (A.5)."""
    q = query @ p["decoder.attention_cell.proj_query.weight"].t()
    q = q / math.sqrt(H)
    score = torch.einsum("bh,bth->bt", q, mem_value)
    if masks is not None:
        score = torch.where(masks > 0, score, torch.full_like(score, NEG))
        w = torch.softmax(score, dim=-1) * masks
    else:
        w = torch.softmax(score, dim=-1)
    return torch.einsum("bt,bth->bh", w, mem_value), w


def decoder_step(p, step_input, states, cell="lstm", H=128, num_layers=2, use_residual=False):
    """gnmt.py:345-404: step_input (rows,E) already embedded.  Returns rnn_out (rows,H), new states."""
    rnn_states, attention_output, mem_value, masks = states
    new_rnn_states = []
    rnn_out, st = _cell_step(cell, p, "decoder.rnn_cells.0.", torch.cat([step_input, attention_output], -1), rnn_states[0])
    new_rnn_states.append(st)
    attention_vec, _ = attention(p, rnn_out, mem_value, masks, H)
    for i in range(1, num_layers):
        curr = rnn_out
        rnn_out, st = _cell_step(cell, p, "decoder.rnn_cells.%d." % i, torch.cat([curr, attention_vec], -1), rnn_states[i])
        if use_residual:
            rnn_out = rnn_out + curr
        new_rnn_states.append(st)
    return rnn_out, [new_rnn_states, attention_vec, mem_value, masks]


def decode_step_logits(p, token_ids, states, **kw):
    """NMTModel.decode_step: tgt_proj(decoder(tgt_embed(ids), states)) (A.6)."""
    emb = p["tgt_embed.weight"][token_ids.long()]
    out, new_states = decoder_step(p, emb, states, **kw)
    return out @ p["tgt_proj.weight"].t() + p["tgt_proj.bias"], new_states


def decode_seq_logits(p, tgt_ids, states, tgt_vl=None, **kw):
    """NMTModel.decode_seq (teacher forcing, gnmt.py:254-304): (B,T_tgt) ids -> (B,T_tgt,V) logits (masked)."""
    B, T = tgt_ids.shape
    outs = []
    for t in range(T):
        emb = p["tgt_embed.weight"][tgt_ids[:, t].long()]
        o, states = decoder_step(p, emb, states, **kw)
        outs.append(o)
    out = torch.stack(outs, 1)
    if tgt_vl is not None:
        out = out * (torch.arange(T).reshape(1, T) < tgt_vl.reshape(B, 1)).to(out.dtype).unsqueeze(-1)
    return out @ p["tgt_proj.weight"].t() + p["tgt_proj.bias"]


def nmt_forward(p, src, tgt_ids, src_vl, tgt_vl, cell="lstm", H=128, num_layers=2, num_bi_layers=1):
    """NMTModel.forward (A.6): encode -> init_state_from_encoder -> decode_seq -> logits."""
    mem, st = encoder_forward(p, src, src_vl, cell, H, num_layers, num_bi_layers)
    states = init_state_from_encoder(mem, st, src_vl)
    return decode_seq_logits(p, tgt_ids, states, tgt_vl, cell=cell, H=H, num_layers=num_layers)


def masked_softmax_ce(pred, label, vl):
    """MaskedSoftmaxCELoss (A.8): per-token CE x SequenceMask weights, MEAN over T (padded) -> (B,)."""
    B, T, V = pred.shape
    logp = torch.log_softmax(pred, dim=-1)
    ce = -logp.gather(-1, label.long().unsqueeze(-1)).squeeze(-1)
    w = (torch.arange(T).reshape(1, T) < vl.reshape(B, 1)).to(pred.dtype)
    return (ce * w).mean(dim=1)


# ----------------------------------------------------------------------------------------------------------------------
def beam_search_scorer(log_probs, scores, step, alpha=1.0, K=5):
    """gluonnlp BeamSearchScorer (A.7)."""
    prev_lp = (K + step - 1) ** alpha / (K + 1) ** alpha if step != 1 else 1.0
    lp = (K + step) ** alpha / (K + 1) ** alpha
    return (log_probs + (scores * prev_lp).unsqueeze(-1)) / lp


def _gather_states(states, gidx):
    rnn_states, att, mem, masks = states
    return [[[t[gidx] for t in layer] for layer in rnn_states], att[gidx], mem[gidx], None if masks is None else masks[gidx]]


def beam_search(p, mem, rnn_states, src_vl, beam=5, max_length=150, bos=2, eos=3, alpha=1.0, K=5, **kw):
    """gluonnlp BeamSearchSampler (A.7) driven by log_softmax(decode_step) (utils/translation.py:51-53).
    Returns samples (B,beam,L) int32, scores (B,beam) descending, valid_length (B,beam) int32."""
    B = mem.shape[0]
    V = p["tgt_proj.weight"].shape[0]
    rep = torch.arange(B).repeat_interleave(beam)
    states = _gather_states(init_state_from_encoder(mem, rnn_states, src_vl), rep)
    step_input = torch.full((B * beam,), float(bos))
    alive = torch.ones(B, beam)
    vlen = torch.ones(B, beam, dtype=torch.int32)
    scores = torch.zeros(B, beam)
    scores[:, 1:] = NEG
    samples = step_input.reshape(B, beam, 1).clone()
    batch_shift = (torch.arange(B) * beam).reshape(B, 1)
    for i in range(max_length):
        logits, new_states = decode_step_logits(p, step_input, states, **kw)
        log_probs = torch.log_softmax(logits, dim=-1).reshape(B, beam, V)
        cand = beam_search_scorer(log_probs, scores, i + 1, alpha, K)
        cand = alive.unsqueeze(-1) * cand + (1 - alive.unsqueeze(-1)) * NEG
        fin = torch.where(alive > 0, torch.full_like(scores, NEG), scores)
        allc = torch.cat([cand.reshape(B, beam * V), fin], dim=1)
        # topk, ties -> lowest index first: stable descending sort
        order = torch.sort(allc, dim=1, descending=True, stable=True).indices[:, :beam]
        new_scores = allc.gather(1, order)
        use_prev = order >= beam * V
        word = torch.where(use_prev, torch.full_like(order, -1), order % V)
        src_beam = torch.where(use_prev, order - beam * V, order // V)
        gidx = (src_beam + batch_shift).reshape(-1)
        samples = torch.cat([samples.reshape(B * beam, -1)[gidx].reshape(B, beam, -1), word.unsqueeze(-1).float()], dim=-1)
        vlen = vlen.reshape(-1)[gidx].reshape(B, beam) + 1 - use_prev.int()
        states = _gather_states(new_states, gidx)
        alive = alive.reshape(-1)[gidx].reshape(B, beam) * (word != eos).float()
        scores = new_scores
        step_input = word.clamp(min=0).reshape(-1).float()
        if alive.sum() == 0:
            return samples.round().int(), scores, vlen
    final = torch.where(alive > 0, torch.full_like(alive, float(eos)), torch.full_like(alive, -1.0))
    samples = torch.cat([samples, final.unsqueeze(-1)], dim=-1)
    vlen = vlen + alive.int()
    return samples.round().int(), scores, vlen


def translate(p, src, src_vl, cell="lstm", H=128, num_layers=2, num_bi_layers=1, beam=5, max_length=150, bos=2, eos=3,
              alpha=1.0, K=5):
    """BeamSearchTranslator.translate (utils/translation.py:55-82)."""
    mem, st = encoder_forward(p, src, src_vl, cell, H, num_layers, num_bi_layers)
    return beam_search(p, mem, st, src_vl, beam=beam, max_length=max_length, bos=bos, eos=eos, alpha=alpha, K=K,
                       cell=cell, H=H, num_layers=num_layers)


def best_tokens(samples, vlen):
    """train_gnmt.py:289-294: best beam, BOS/EOS stripped."""
    out = []
    for i in range(samples.shape[0]):
        out.append([int(t) for t in samples[i, 0, 1:int(vlen[i, 0]) - 1]])
    return out
