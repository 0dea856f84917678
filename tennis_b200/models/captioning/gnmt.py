"""GNMT encoder / decoder (mirror of the reference's models/captioning/gnmt.py: same classes, constructor
signatures and call conventions), NMTModel glue and the scorer that gluonnlp provided.  All arithmetic runs in
libtennis_b200.so: the encoder through the fused RNN scan with valid_length, the decoder through tn_gnmt_*."""
import torch

from ... import ops
from ...gluon import Block, Dense, Dropout, Embedding, HybridSequential, Parameter


def _gates(cell_type):
    if cell_type not in ("gru", "lstm"):
        raise ValueError("cell_type must be 'gru' or 'lstm' (train_gnmt.py:63)")
    return 3 if cell_type == "gru" else 4


class _Cell(Block):
    """rnn.GRUCell / rnn.LSTMCell parameter holder (i2h_weight (G*H,in), h2h_weight (G*H,H), biases)."""

    def __init__(self, cell_type, hidden_size, input_size=0, i2h_bias_initializer="zeros", **kw):
        super(_Cell, self).__init__(**kw)
        G = _gates(cell_type)
        self.cell_type, self.hidden_size = cell_type, hidden_size
        self.i2h_weight = Parameter("i2h_weight", (G * hidden_size, input_size))
        self.h2h_weight = Parameter("h2h_weight", (G * hidden_size, hidden_size))
        self.i2h_bias = Parameter("i2h_bias", (G * hidden_size,), init=i2h_bias_initializer)
        self.h2h_bias = Parameter("h2h_bias", (G * hidden_size,), init="zeros")

    def state_info(self):
        return [None] if self.cell_type == "gru" else [None, None]

    def ensure(self, in_dim, device):
        if self.i2h_weight._data is None:
            self.i2h_weight._finish_deferred((self.i2h_weight.shape[0], in_dim))
        for p in self._reg_params.values():
            if p._data is not None and p._data.device != device:
                p.reset_ctx(device)

    def weights(self):
        return {n: p.data() for n, p in self._reg_params.items()}

    def version(self):
        return tuple(p._version for p in self._reg_params.values())


class BidirectionalCell(Block):
    def __init__(self, l_cell, r_cell, **kw):
        super(BidirectionalCell, self).__init__(**kw)
        self.l_cell, self.r_cell = l_cell, r_cell

    def state_info(self):
        return self.l_cell.state_info() + self.r_cell.state_info()


class GNMTEncoder(Block):
    """Bidirectional layers followed by unidirectional ones (reference gnmt.py:30-160)."""

    def __init__(self, cell_type='lstm', num_layers=2, num_bi_layers=1, hidden_size=128, dropout=0.0, use_residual=True,
                 i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer='zeros',
                 h2h_bias_initializer='zeros', prefix=None, params=None):
        super(GNMTEncoder, self).__init__(prefix=prefix, params=params)
        assert num_bi_layers <= num_layers, \
            'Number of bidirectional layers must be smaller than the total number of layers, ' \
            'num_bi_layers={}, num_layers={}'.format(num_bi_layers, num_layers)
        self._cell_type, self._num_bi_layers, self._num_layers = cell_type, num_bi_layers, num_layers
        self._hidden_size, self._dropout, self._use_residual = hidden_size, dropout, use_residual
        with self.name_scope():
            self.dropout_layer = Dropout(dropout)
            self.rnn_cells = HybridSequential()
            for i in range(num_layers):
                if i < num_bi_layers:
                    self.rnn_cells.add(BidirectionalCell(
                        l_cell=_Cell(cell_type, hidden_size, i2h_bias_initializer=i2h_bias_initializer),
                        r_cell=_Cell(cell_type, hidden_size, i2h_bias_initializer=i2h_bias_initializer)))
                else:
                    self.rnn_cells.add(_Cell(cell_type, hidden_size, i2h_bias_initializer=i2h_bias_initializer))
        self._engines = {}

    def _engine(self, i, cell, in_dim, device):
        cells = [cell.l_cell, cell.r_cell] if isinstance(cell, BidirectionalCell) else [cell]
        for c in cells:
            c.ensure(in_dim, device)
        key = (device.index or 0, in_dim) + tuple(c.version() for c in cells)
        hit = self._engines.get(i)
        if hit is None or hit[0] != key:
            params = {}
            for d, c in zip(("l0", "r0"), cells):
                for n, t in c.weights().items():
                    params[d + "_" + n] = t
            eng = ops.BiRNN(self._cell_type, in_dim, self._hidden_size, params, bidirectional=len(cells) == 2,
                            device=device.index or 0, precise=True)
            self._engines[i] = (key, eng)
        return self._engines[i][1]

    def __call__(self, inputs, states=None, valid_length=None):
        return self.forward(inputs, states, valid_length)

    def forward(self, inputs, states=None, valid_length=None):
        """inputs (B,T,C) -> [outputs (B,T,H), new_states], []   (reference gnmt.py:136-160).
        With valid_length: per-row scan lengths, reverse direction starting at the last valid step, outputs zeroed
        past valid_length and states taken at the last valid step (MXNet unroll(valid_length=...) semantics)."""
        if states is not None:
            raise NotImplementedError("the scripts never pass initial encoder states (train_gnmt.py / translation.py)")
        ops._require_cuda(inputs, valid_length)
        if self._dropout and getattr(self, "_training", False):
            raise NotImplementedError("dropout > 0 is a training-time feature")
        new_states = []
        x = inputs
        outputs = inputs
        vl = None if valid_length is None else valid_length.to(torch.int32)
        for i, cell in enumerate(self.rnn_cells):
            eng = self._engine(i, cell, x.shape[2], x.device)
            res = eng(x.float(), valid_len=vl, want_y=True, want_state=True)
            outputs = res["y"]
            if i < self._num_bi_layers:
                # "For bidirectional RNN, we use the states of the backward RNN" (gnmt.py:146-148)
                st = [res["h"][1]] if self._cell_type == "gru" else [res["h"][1], res["c"][1]]
            else:
                st = [res["h"][0]] if self._cell_type == "gru" else [res["h"][0], res["c"][0]]
            new_states.append(st)
            outputs = self.dropout_layer(outputs)
            if self._use_residual and i > self._num_bi_layers:  # strict '>' as in the reference (gnmt.py:153-155)
                outputs = outputs + x
            x = outputs
        # outputs past valid_length are already zero (the scan never writes them): SequenceMask of gnmt.py:157-159
        return [outputs, new_states], []


class _AttentionCell(Block):
    """gluonnlp DotProductAttentionCell(units=H, scaled=True, use_bias=False, luong_style=True): only the query is
    projected."""

    def __init__(self, units, **kw):
        super(_AttentionCell, self).__init__(**kw)
        self.proj_query = Dense(units, in_units=units, flatten=False, use_bias=False)


class GNMTDecoder(Block):
    """gnmt_v2-style decoder with scaled-Luong attention (reference gnmt.py:163-404)."""

    def __init__(self, cell_type='lstm', attention_cell='scaled_luong', num_layers=2, hidden_size=128, dropout=0.0,
                 use_residual=True, output_attention=False, i2h_weight_initializer=None, h2h_weight_initializer=None,
                 i2h_bias_initializer='zeros', h2h_bias_initializer='zeros', prefix=None, params=None):
        super(GNMTDecoder, self).__init__(prefix=prefix, params=params)
        if attention_cell != 'scaled_luong':
            raise NotImplementedError("the scripts only use attention_cell='scaled_luong' (gnmt.py:205)")
        if output_attention:
            raise NotImplementedError("output_attention is never enabled by the scripts")
        self._cell_type, self._num_layers, self._hidden_size = cell_type, num_layers, hidden_size
        self._dropout, self._use_residual, self._output_attention = dropout, use_residual, output_attention
        with self.name_scope():
            self.attention_cell = _AttentionCell(hidden_size)
            self.dropout_layer = Dropout(dropout)
            self.rnn_cells = HybridSequential()
            for i in range(num_layers):
                self.rnn_cells.add(_Cell(cell_type, hidden_size, i2h_bias_initializer=i2h_bias_initializer))

    def init_state_from_encoder(self, encoder_outputs, encoder_valid_length=None):
        """[rnn_states, attention_vec (zeros), mem_value, mem_masks]  (reference gnmt.py:224-252)."""
        mem_value, rnn_states = encoder_outputs
        batch_size, mem_length, mem_size = mem_value.shape
        attention_vec = torch.zeros((batch_size, mem_size), device=mem_value.device)
        decoder_states = [rnn_states, attention_vec, mem_value]
        if encoder_valid_length is not None:
            mem_masks = (torch.arange(mem_length, device=mem_value.device).reshape(1, -1)
                         < encoder_valid_length.reshape(-1, 1)).float()
            decoder_states.append(mem_masks)
        return decoder_states

    # ---- the decoder block's own step API (reference gnmt.py:254-343); the scripts reach it through NMTModel, which fuses the
    # target embedding and projection around the same device step
    def _block_engine(self, E, device):
        H = self._hidden_size
        for i, c in enumerate(self.rnn_cells):
            c.ensure(E + H if i == 0 else 2 * H, device)
        plist = list(self.collect_params().values())
        key = (device.index or 0, E) + tuple(p._version for p in plist)
        if getattr(self, "_blk_engine", None) is None or self._blk_key != key:
            z = torch.zeros  # embedding / projection are outside the block: one-row placeholders, never read by decoder_step
            self._blk_engine = ops.GNMTDecoderEngine(
                self._cell_type, H, E, 1, [c.weights() for c in self.rnn_cells], self.attention_cell.proj_query.weight.data(),
                z(1, E), z(1, H), z(1), use_residual=self._use_residual, device=device.index or 0)
            self._blk_key = key
        return self._blk_engine

    def __call__(self, step_input, states):
        """One-step-ahead decoding (gnmt.py:306-404): step_input (B, C_in) embedded target -> (rnn_out (B,H),
        [rnn_states, attention_vec, mem_value(, mem_masks)], [])."""
        ops._require_cuda(step_input, states[2])
        if self._dropout and getattr(self, "_training", False):
            raise NotImplementedError("dropout > 0 is a training-time feature (training runs through GNMTTrainGraph)")
        eng = self._block_engine(step_input.shape[1], states[2].device)
        h, c = NMTModel._pack_states(states[0])
        rows_per_mem = step_input.shape[0] // states[2].shape[0]
        out, h2, c2, att2 = eng.decoder_step(step_input, h, c, states[1], states[2], NMTModel._mask_to_len(states), rows_per_mem)
        return out, [NMTModel._unpack_states(h2, c2), att2] + list(states[2:]), []

    forward = __call__

    def decode_seq(self, inputs, states, valid_length=None):
        """Decode the (embedded) decoder inputs step by step (gnmt.py:254-304): inputs (B, T, C_in) -> output (B, T, H) masked past
        valid_length, states taken at each row's last valid step (_nested_sequence_last), []."""
        B, T = inputs.shape[:2]
        fixed = list(states[2:])
        outs, rnn_l, att_l = [], [], []
        for i in range(T):
            o, states, _ = self(inputs[:, i].contiguous(), states)
            outs.append(o)
            rnn_l.append(states[0])
            att_l.append(states[1])
        output = torch.stack(outs, dim=1)
        if valid_length is not None:
            last = (valid_length.to(torch.int64) - 1).clamp(min=0)
            rows = torch.arange(B, device=output.device)

            def seq_last(per_step):
                return torch.stack(per_step, dim=0)[last, rows]
            rnn_states = [[seq_last([st[l][k] for st in rnn_l]) for k in range(len(rnn_l[0][l]))] for l in range(len(rnn_l[0]))]
            states = [rnn_states, seq_last(att_l)] + fixed
            keep = torch.arange(T, device=output.device).reshape(1, T, 1) < valid_length.reshape(B, 1, 1)
            output = torch.where(keep, output, torch.zeros_like(output))  # SequenceMask
        return output, states, []


def get_gnmt_encoder_decoder(cell_type='lstm', attention_cell='scaled_luong', num_layers=2, num_bi_layers=1,
                             hidden_size=128, dropout=0.0, use_residual=False, i2h_weight_initializer=None,
                             h2h_weight_initializer=None, i2h_bias_initializer='lstmbias', h2h_bias_initializer='zeros',
                             prefix='gnmt_', params=None):
    """Build a pair of GNMT encoder/decoder (reference gnmt.py:407-455; default i2h bias = LSTMBias(forget_bias=1.0))."""
    encoder = GNMTEncoder(cell_type=cell_type, num_layers=num_layers, num_bi_layers=num_bi_layers, hidden_size=hidden_size,
                          dropout=dropout, use_residual=use_residual, i2h_bias_initializer=i2h_bias_initializer,
                          h2h_bias_initializer=h2h_bias_initializer, prefix=prefix + 'enc_', params=params)
    decoder = GNMTDecoder(cell_type=cell_type, attention_cell=attention_cell, num_layers=num_layers, hidden_size=hidden_size,
                          dropout=dropout, use_residual=use_residual, i2h_bias_initializer=i2h_bias_initializer,
                          h2h_bias_initializer=h2h_bias_initializer, prefix=prefix + 'dec_', params=params)
    return encoder, decoder


class NMTModel(Block):
    """gluonnlp.model.translation.NMTModel as the scripts use it (train_gnmt.py:228-229)."""

    def __init__(self, src_vocab, tgt_vocab, encoder, decoder, embed_size=None, prefix=None, src_embed=None, tgt_embed=None,
                 **kw):
        super(NMTModel, self).__init__(prefix=prefix)
        self.src_vocab, self.tgt_vocab = src_vocab, tgt_vocab
        self.encoder, self.decoder = encoder, decoder
        if src_embed is None:
            raise NotImplementedError("the scripts always pass src_embed (features or TimeDistributed CNN)")
        self.src_embed = src_embed
        self.tgt_embed = tgt_embed if tgt_embed is not None else Embedding(len(tgt_vocab), embed_size)
        self.tgt_proj = Dense(len(tgt_vocab), in_units=decoder._hidden_size, flatten=False)
        self._engine = None
        self._engine_key = None

    # ---- engine over decoder + tgt_embed + tgt_proj parameters
    def _get_engine(self, device):
        dec = self.decoder
        H = dec._hidden_size
        E = self.tgt_embed.weight.shape[1]
        for i, c in enumerate(dec.rnn_cells):
            c.ensure(E + H if i == 0 else 2 * H, device)
        plist = list(dec.collect_params().values()) + [self.tgt_embed.weight, self.tgt_proj.weight, self.tgt_proj.bias]
        for p in plist:
            if p._data is not None and p._data.device != device:
                p.reset_ctx(device)
        key = (device.index or 0,) + tuple(p._version for p in plist)
        if self._engine is None or self._engine_key != key:
            self._engine = ops.GNMTDecoderEngine(
                dec._cell_type, H, E, self.tgt_proj.weight.shape[0], [c.weights() for c in dec.rnn_cells],
                dec.attention_cell.proj_query.weight.data(), self.tgt_embed.weight.data(), self.tgt_proj.weight.data(),
                self.tgt_proj.bias.data(), use_residual=dec._use_residual, device=device.index or 0)
            self._engine_key = key
        return self._engine

    @staticmethod
    def _pack_states(rnn_states):
        h = torch.stack([s[0] for s in rnn_states])
        c = torch.stack([s[1] for s in rnn_states]) if len(rnn_states[0]) > 1 else None
        return h, c

    @staticmethod
    def _unpack_states(h, c):
        return [[h[i]] if c is None else [h[i], c[i]] for i in range(h.shape[0])]

    @staticmethod
    def _mask_to_len(states):
        return states[3].sum(dim=1).to(torch.int32) if len(states) == 4 else None

    # ---- NMTModel API
    def encode(self, inputs, states=None, valid_length=None):
        return self.encoder(self.src_embed(inputs), states, valid_length)

    def decode_step(self, step_input, states):
        """(rows,) token ids -> logits (rows,V), new states, []  (model.decode_step, translation.py:52)."""
        eng = self._get_engine(states[2].device)
        h, c = self._pack_states(states[0])
        rows_per_mem = step_input.shape[0] // states[2].shape[0]
        logits, h2, c2, att2 = eng.decode_step(step_input, h, c, states[1], states[2], self._mask_to_len(states), rows_per_mem)
        return logits, [self._unpack_states(h2, c2), att2] + list(states[2:]), []

    def decode_seq(self, inputs, states, valid_length=None):
        eng = self._get_engine(states[2].device)
        h, c = self._pack_states(states[0])
        logits = eng.decode_seq(inputs, valid_length, h, c, states[2], self._mask_to_len(states))
        return logits, states, []

    def forward(self, src_seq, tgt_seq, src_valid_length=None, tgt_valid_length=None):
        """encode -> init_state_from_encoder -> decode_seq -> (logits (B,T_tgt,V), additional outputs).
        Under autograd.record() (train_gnmt.py:330-334) the same computation runs through the training graph, which keeps
        the activations and tags the logits with its backward."""
        from ... import autograd
        if autograd.is_recording():
            from .train_graph import GNMTTrainGraph
            self._train_calls = getattr(self, "_train_calls", 0) + 1
            graph = GNMTTrainGraph(self, seed=self._train_calls)
            src = self.src_embed(src_seq)
            logits = graph.forward(src, tgt_seq, src_valid_length, tgt_valid_length)
            # a source produced by a trainable TimeDistributed(CNN) continues the walk with d(loss)/d(source features)
            autograd.tag(logits, graph.backward, src if getattr(src, "_tn_node", None) is not None else None)
            return logits, [[], []]
        encoder_outputs, enc_add = self.encode(src_seq, valid_length=src_valid_length)
        decoder_states = self.decoder.init_state_from_encoder(encoder_outputs, encoder_valid_length=src_valid_length)
        outputs, _, dec_add = self.decode_seq(tgt_seq, decoder_states, tgt_valid_length)
        return outputs, [enc_add, dec_add]

    def beam_search(self, decoder_states, beam_size, max_length, alpha, K, bos, eos):
        eng = self._get_engine(decoder_states[2].device)
        h, c = self._pack_states(decoder_states[0])
        return eng.beam_search(decoder_states[2], self._mask_to_len(decoder_states), h, c, beam_size, max_length, alpha, K,
                               bos, eos)


class BeamSearchScorer(object):
    """gluonnlp.model.BeamSearchScorer(alpha, K): length-normalised cumulative log-probability."""

    def __init__(self, alpha=1.0, K=5.0):
        self._alpha, self._K = float(alpha), float(K)
