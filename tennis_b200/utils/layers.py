"""Custom layers (mirror of the reference's utils/layers.py)."""
from ..gluon import Block


class TimeDistributed(Block):
    """Apply `model` to every timestep (reference utils/layers.py:8-48).

    'reshape' folds (B,T,...) -> (B*T,...) (a zero-copy view, layers.py:39), runs the model once on all frames and
    unfolds the result to (B,T,...) (layers.py:46).  'for' (layers.py:27-36) yields the same values; here it is
    routed to the same folded launch because per-timestep launches only shrink the GEMM M dimension.
    Tuple/list outputs are unfolded element-wise (layers.py:41-44)."""

    def __init__(self, model, style='reshape', **kwargs):
        super(TimeDistributed, self).__init__(**kwargs)
        assert style in ['reshape', 'for']
        self._style = style
        with self.name_scope():
            self.model = model

    @staticmethod
    def _unfold(y, B, T):
        out = y.reshape((B, T) + tuple(y.shape[1:]))
        twin = getattr(y, "_tn_bf16", None)
        if twin is not None:
            out._tn_bf16 = twin.reshape((B, T) + tuple(twin.shape[1:]))
        if getattr(y, "_tn_precise", False):  # fp32-grade features: downstream layers keep the split-bf16 arithmetic
            out._tn_precise = True
        if getattr(y, "_tn_node", None) is not None:  # recorded (trainable model): the unfold is a view, pass the gradient through
            from .. import autograd
            autograd.tag(out, lambda g, shp=tuple(y.shape): g.reshape(shp), y)
        return out

    def forward(self, x):
        B, T = x.shape[0], x.shape[1]
        y = self.model(x.reshape((B * T,) + tuple(x.shape[2:])))
        if isinstance(y, tuple):
            return tuple(self._unfold(yi, B, T) for yi in y)
        if isinstance(y, list):
            return [self._unfold(yi, B, T) for yi in y]
        return self._unfold(y, B, T)
