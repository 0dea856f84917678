"""Model specifications (mirror of the reference's models/vision/definitions.py; same class names, constructor
signatures and attributes, so train.py / evaluate.py assemble them unchanged)."""
from ... import rnn
from ...gluon import Block, Dense
from ...utils.layers import TimeDistributed
from ... import ops


class FrameModel(Block):
    def __init__(self, backbone, num_classes=-1, swap=False, **kwargs):
        """A framewise model: backbone CNN + one Dense layer to the classes (reference definitions.py:10-33)."""
        super(FrameModel, self).__init__(**kwargs)
        self.swap = swap
        with self.name_scope():
            self.backbone = backbone
            self.classes = None
            if num_classes > 0:
                self.classes = Dense(num_classes, flatten=True)

    def forward(self, x):
        if self.swap:
            x = x.transpose(1, 2)
        x = self.backbone(x)
        if self.classes:
            x = self.classes(x)
        return x


class TemporalPooling(Block):
    def __init__(self, model, num_classes=-1, pool='max', feats=False, **kwargs):
        """Temporal pooling over per-frame outputs (reference definitions.py:36-72).
        num_classes: -1 -> model output is pooled as is; 0 -> pool the backbone features, then model.classes."""
        super(TemporalPooling, self).__init__(**kwargs)
        self.pool = pool
        self.feats = feats
        with self.name_scope():
            self.classes = None
            if model is not None:
                if num_classes == 0:
                    self.td = TimeDistributed(model.backbone)
                    self.classes = model.classes
                else:
                    self.td = TimeDistributed(model)
                    if num_classes > 0:
                        self.classes = Dense(num_classes, flatten=True)
            else:
                self.classes = Dense(num_classes, flatten=True)

    def forward(self, x):
        if not self.feats:
            x = self.td(x)
        x = ops.temporal_pool(x, 'mean' if self.pool == 'mean' else 'max')
        if self.classes:
            x = self.classes(x)
        return x


class CNNRNN(Block):
    def __init__(self, model, num_classes=-1, hidden_size=128, type='gru', **kwargs):
        """CNN + bidirectional GRU/LSTM + max over time + Dense (reference definitions.py:75-110).
        model=None -> inputs are pre-extracted features (the published 0042 configuration)."""
        super(CNNRNN, self).__init__(**kwargs)
        self.feats = model is None
        with self.name_scope():
            if model is not None:
                self.td = TimeDistributed(model.backbone)
            if type == 'lstm':
                self.rnn = rnn.LSTM(hidden_size, layout="NTC", bidirectional=True)
            else:
                self.rnn = rnn.GRU(hidden_size, layout="NTC", bidirectional=True)
            self.classes = None
            if num_classes == 0:
                self.classes = model.classes
            elif num_classes > 0:
                self.classes = Dense(num_classes, flatten=True)

    def forward(self, x):
        if not self.feats:
            x = self.td(x)
        x = self.rnn.forward_max(x)  # rnn(x) followed by F.max(axis=1), fused in the scan kernel
        if self.classes:
            x = self.classes(x)
        return x
