"""mx.gluon.rnn.GRU / LSTM layer stand-ins (layout 'NTC'), running tn_birnn_forward."""
from . import ops
from .gluon import Block, Parameter


class _RNNLayer(Block):
    _cell = None
    _gates = 0

    def __init__(self, hidden_size, num_layers=1, layout='NTC', bidirectional=False, input_size=0, **kw):
        super(_RNNLayer, self).__init__(**kw)
        if num_layers != 1 or layout != 'NTC':
            raise NotImplementedError("hot path uses 1 layer, layout='NTC' (definitions.py:94-96)")
        self._hidden_size, self._input_size, self._bidirectional = hidden_size, input_size, bidirectional
        G, H = self._gates, hidden_size
        for d in (['l0', 'r0'] if bidirectional else ['l0']):
            self._reg_params[d + '_i2h_weight'] = Parameter(d + '_i2h_weight', (G * H, input_size))
            self._reg_params[d + '_h2h_weight'] = Parameter(d + '_h2h_weight', (G * H, H))
            self._reg_params[d + '_i2h_bias'] = Parameter(d + '_i2h_bias', (G * H,), init='zeros')
            self._reg_params[d + '_h2h_bias'] = Parameter(d + '_h2h_bias', (G * H,), init='zeros')
        self._engine = None
        self._engine_key = None

    def _get_engine(self, x, precise=False):
        D = x.shape[2]
        for n, p in self._reg_params.items():
            if p._data is None and n.endswith('_i2h_weight'):
                p._finish_deferred((self._gates * self._hidden_size, D))
                p.reset_ctx(x.device)
        static = (x.device.index or 0, D)
        versions = tuple(p._version for p in self._reg_params.values())
        params = {n: p.data() for n, p in self._reg_params.items()}
        if self._engine is None or self._engine_key[0] != static:
            self._engine = ops.BiRNN(self._cell, D, self._hidden_size, params, self._bidirectional, device=x.device.index or 0)
            self._engine_key = (static, versions)
        elif self._engine_key[1] != versions:
            # parameters changed (Trainer.step, set_data, load_parameters): re-pack on the device instead of rebuilding
            if all(t.is_cuda for t in params.values()):
                self._engine.update_weights(params)
            else:
                self._engine = ops.BiRNN(self._cell, D, self._hidden_size, params, self._bidirectional, device=x.device.index or 0)
            self._engine_key = (static, versions)
        self._engine.set_precise(precise)
        return self._engine

    @staticmethod
    def _pick_input(x):
        twin = getattr(x, "_tn_bf16", None)
        return twin if twin is not None else x

    def _train_forward(self, x, pooled):
        from . import autograd
        # training runs the input projection in split-bf16 on the fp32 features: the max-over-time arg-max (which routes
        # the gradient) must not flip on bf16 rounding noise
        eng = self._get_engine(x, precise=True)
        saved = eng.forward_train(x.float())
        out = saved["ymax"] if pooled else saved["y"]

        upstream = x if getattr(x, "_tn_node", None) is not None else None  # features of a trainable CNN (end-to-end training)

        def bwd(g, eng=eng, saved=saved, self=self):
            grads = eng.backward(saved, d_ymax=g if pooled else None, dy=None if pooled else g, want_dx=upstream is not None,
                                 weights={n: p.data() for n, p in self._reg_params.items()})
            dx = grads.pop("dx", None)
            for name, gr in grads.items():
                self._reg_params[name]._accumulate_grad(gr)
            return dx  # None when the per-frame features are leaves (frozen / pre-extracted backbone)
        return autograd.tag(out, bwd, upstream)

    def forward(self, x):
        """(B,T,D) -> (B,T,ndir*H)."""
        from . import autograd
        ops._require_cuda(x)
        if autograd.is_recording():
            return self._train_forward(x, pooled=False)
        if getattr(x, "_tn_precise", False):
            return self._get_engine(x, precise=True)(x.float(), want_y=True)["y"]
        return self._get_engine(x)(self._pick_input(x), want_y=True)["y"]

    def forward_max(self, x):
        """max over time of the layer output, fused into the scan: (B,T,D) -> (B,ndir*H) (definitions.py:106-107)."""
        from . import autograd
        ops._require_cuda(x)
        if autograd.is_recording():
            return self._train_forward(x, pooled=True)
        if getattr(x, "_tn_precise", False):
            return self._get_engine(x, precise=True)(x.float(), want_y=False, want_max=True)["ymax"]
        return self._get_engine(x)(self._pick_input(x), want_y=False, want_max=True)["ymax"]


class GRU(_RNNLayer):
    _cell, _gates = 'gru', 3


class LSTM(_RNNLayer):
    _cell, _gates = 'lstm', 4
