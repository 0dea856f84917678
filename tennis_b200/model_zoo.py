"""`get_model(name).features` for the two backbones the reference's scripts use (train.py:204, flag default
`resnet18_v2`, published models `DenseNet121`).  Stands in for gluoncv.model_zoo.get_model: same call shape,
parameter inventory in Gluon's collect_params() order; the forward runs the sm_100a layer plan in
libtennis_b200.so (tn_backbone_forward)."""
import os

import torch

from . import ops
from .gluon import Block, Parameter

_DENSE_CFG = (6, 12, 24, 16)
_RES_CH = (64, 128, 256, 512)


def _bn(prefix):
    return [(prefix + ".gamma", "ones"), (prefix + ".beta", "zeros"), (prefix + ".running_mean", "zeros"),
            (prefix + ".running_var", "ones")]


def _densenet_inventory():
    inv = [("conv0.weight", (64, 3, 7, 7), "uniform")] + [(n, (64,), i) for n, i in _bn("bn0")]
    c = 64
    for b, nl in enumerate(_DENSE_CFG):
        for l in range(nl):
            p = "block%d.layer%d" % (b + 1, l + 1)
            inv += [(n, (c,), i) for n, i in _bn(p + ".bn1")]
            inv += [(p + ".conv1.weight", (128, c, 1, 1), "uniform")]
            inv += [(n, (128,), i) for n, i in _bn(p + ".bn2")]
            inv += [(p + ".conv2.weight", (32, 128, 3, 3), "uniform")]
            c += 32
        if b < 3:
            p = "trans%d" % (b + 1)
            inv += [(n, (c,), i) for n, i in _bn(p + ".bn")]
            inv += [(p + ".conv.weight", (c // 2, c, 1, 1), "uniform")]
            c //= 2
    inv += [(n, (c,), i) for n, i in _bn("bn5")]
    return inv


def _resnet18_inventory():
    inv = [(n, (3,), i) for n, i in _bn("bn_data")]
    inv += [("conv0.weight", (64, 3, 7, 7), "uniform")] + [(n, (64,), i) for n, i in _bn("bn0")]
    cin = 64
    for s, c in enumerate(_RES_CH):
        for b in range(2):
            p = "stage%d.block%d" % (s + 1, b + 1)
            inv += [(n, (cin,), i) for n, i in _bn(p + ".bn1")]
            inv += [(p + ".conv1.weight", (c, cin, 3, 3), "uniform")]
            inv += [(n, (c,), i) for n, i in _bn(p + ".bn2")]
            inv += [(p + ".conv2.weight", (c, c, 3, 3), "uniform")]
            if b == 0 and cin != c:
                inv += [(p + ".downsample.weight", (c, cin, 1, 1), "uniform")]
            cin = c
    inv += [(n, (512,), i) for n, i in _bn("bn_final")]
    return inv


_INVENTORY = {"densenet121": _densenet_inventory, "resnet18_v2": _resnet18_inventory}


class Features(Block):
    """`.features` of a zoo model: (N,3,H,W) fp32 normalised [or (N,H,W,3) uint8] -> (N,D) fp32."""

    def __init__(self, arch, **kw):
        super(Features, self).__init__(**kw)
        self.arch = arch
        self._names = []
        for name, shape, init in _INVENTORY[arch]():
            p = Parameter(name, shape, init=init)
            if name.startswith("bn_data.") and (name.endswith("gamma") or name.endswith("beta")):
                p.grad_req = "null"  # BatchNorm(scale=False, center=False)
            if name.endswith("running_mean") or name.endswith("running_var"):
                p.grad_req = "null"
            self._reg_params[name] = p
            self._names.append(name)
        self._engine = None
        self._engine_key = None
        # arithmetic of the inference engine: 'bf16' (speed) or 'split_bf16' (fp32-grade: logits within 1e-3 of the fp32 reference,
        # DenseNet-121); default from $TN_PRECISION.  Features.precision = '...' switches an existing model.
        self.precision = os.environ.get("TN_PRECISION", "bf16")

    def flat_params(self):
        return torch.cat([self._reg_params[n].data().reshape(-1).float().cpu() for n in self._names])

    def _get_engine(self, device):
        key = (device.index or 0,) + tuple(self._reg_params[n]._version for n in self._names)
        if self._engine is None or self._engine_key != key:
            self._engine = ops.Backbone(self.arch, self.flat_params(), device=device.index or 0)
            self._engine_key = key
        self._engine.set_precision(self.precision)
        return self._engine

    def feature_dim(self, h, w):
        from ._lib import lib
        return lib().tn_backbone_feature_dim(ops.Backbone.ARCH[self.arch], h, w)

    def forward(self, x):
        ops._require_cuda(x)
        from . import autograd
        if autograd.is_recording() and any(p.grad_req != 'null' for n, p in self._reg_params.items()
                                           if not n.endswith(("running_mean", "running_var"))):
            # trainable backbone under ag.record() (train.py:415-421): training-mode forward (batch statistics) with saved
            # activations on the fp32 training path; the returned features carry the backward of the whole CNN
            if x.dim() != 4 or x.dtype == torch.uint8:
                raise ValueError("training expects normalised fp32 frames (N,3,H,W)")
            from .models.vision.train_graph import CNNTrainGraph
            for p in self._reg_params.values():
                if p._data is not None and p._data.device != x.device:
                    p.reset_ctx(x.device)
            graph = CNNTrainGraph(self)
            feats = graph.forward(x)
            return autograd.tag(feats, graph.backward, None)
        eng = self._get_engine(x.device)
        if eng.precision == "split_bf16":
            feats = eng(x)
            feats._tn_precise = True  # the temporal head then runs its input projection in split-bf16 on these fp32 features
            return feats
        feats, fb = eng(x, want_bf16=True)
        feats._tn_bf16 = fb  # bf16 twin written by the same kernel; lets the RNN skip a cast pass
        return feats


class _ZooModel(Block):
    def __init__(self, arch):
        super(_ZooModel, self).__init__()
        self.features = Features(arch)


def get_model(name, pretrained=False, ctx=None, **kwargs):
    arch = name.lower()
    if arch not in _INVENTORY:
        raise ValueError("model '%s' is outside the hot path (supported: %s)" % (name, sorted(_INVENTORY)))
    model = _ZooModel(arch)
    if pretrained:
        # ImageNet weights live on a model-zoo server; this image has no network.  load_parameters() of a converted
        # GluonCV .params file fills the same inventory.
        raise RuntimeError("pretrained=True: no network access for ImageNet weights; pass pretrained=False and "
                           "load_parameters() from a local .params file")
    return model
