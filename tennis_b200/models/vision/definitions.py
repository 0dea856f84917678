"""Event-detector model shells: the three classes the reference's scripts assemble (reference models/vision/definitions.py:
FrameModel :10-33, TemporalPooling :36-72, CNNRNN :75-110) with the same constructor arguments, public attributes (`backbone`,
`td`, `rnn`, `classes`, `pool`, `feats`, `swap`) and structural parameter names, so checkpoints and `train.py` / `evaluate.py`
model assembly carry over.  The forwards dispatch to libtennis_b200.so; the temporal reduction of CNNRNN is fused into the
recurrent scan instead of being a separate op."""
from ... import ops, rnn
from ...gluon import Block, Dense
from ...utils.layers import TimeDistributed


def _classifier(num_classes, inherited=None):
    """The `classes` attribute for a given `num_classes`: a new Dense when positive, the wrapped model's own classifier
    when 0 (only where the reference allows it), nothing when negative."""
    if num_classes > 0:
        return Dense(num_classes, flatten=True)
    return inherited if num_classes == 0 else None


class FrameModel(Block):
    """Backbone CNN followed by one Dense layer over the classes; `swap` exchanges axes 1 and 2 of the input first."""

    def __init__(self, backbone, num_classes=-1, swap=False, **kwargs):
        super(FrameModel, self).__init__(**kwargs)
        self.swap = bool(swap)
        self.backbone = backbone
        self.classes = _classifier(num_classes)

    def forward(self, x):
        feats = self.backbone(x.transpose(1, 2) if self.swap else x)
        return self.classes(feats) if self.classes else feats


class TemporalPooling(Block):
    """Per-frame model applied over time, outputs reduced by mean or max over the time axis, optional classifier.

    num_classes  < 0: the wrapped model's own outputs are pooled as they are;
                 = 0: the wrapped model's *backbone* features are pooled and its classifier runs on the pooled vector;
                 > 0: a new classifier on the pooled outputs.
    model=None (pre-extracted features, `feats=True`): only the classifier exists."""

    def __init__(self, model, num_classes=-1, pool='max', feats=False, **kwargs):
        super(TemporalPooling, self).__init__(**kwargs)
        self.pool, self.feats = pool, feats
        self.classes = None
        if model is None:
            self.classes = Dense(num_classes, flatten=True)
            return
        reuse_head = num_classes == 0
        self.td = TimeDistributed(model.backbone if reuse_head else model)
        self.classes = _classifier(num_classes, inherited=model.classes if reuse_head else None)

    def forward(self, x):
        seq = x if self.feats else self.td(x)
        pooled = ops.temporal_pool(seq, 'mean' if self.pool == 'mean' else 'max')
        return self.classes(pooled) if self.classes else pooled


class CNNRNN(Block):
    """Backbone over time -> one bidirectional GRU/LSTM layer -> max over time -> classifier.

    model=None means the inputs already are per-frame features (the published feature-based configuration);
    num_classes = 0 borrows the wrapped model's classifier, > 0 creates one, < 0 returns the pooled RNN output."""

    def __init__(self, model, num_classes=-1, hidden_size=128, type='gru', **kwargs):
        super(CNNRNN, self).__init__(**kwargs)
        self.feats = model is None
        if not self.feats:
            self.td = TimeDistributed(model.backbone)
        layer = rnn.LSTM if type == 'lstm' else rnn.GRU
        self.rnn = layer(hidden_size, layout="NTC", bidirectional=True)
        self.classes = _classifier(num_classes, inherited=None if self.feats else model.classes)

    def forward(self, x):
        seq = x if self.feats else self.td(x)
        pooled = self.rnn.forward_max(seq)  # rnn(x) and the max over axis 1 in one scan kernel
        return self.classes(pooled) if self.classes else pooled
