"""CPU oracle for the event-detector half of the hot path (TEST INFRASTRUCTURE — never imported by the product;
the seeded weight / input generators it shares with the benchmark live in tennis_b200/synthetic.py).

PARITY UNPINNED: the reference (HaydenFaulkner/Tennis) ships no golden vectors and executes all arithmetic
inside MXNet / GluonCV, which are neither vendored in /root/reference nor installable here.
This file restates, in plain PyTorch fp32/fp64 on the CPU, the published semantics of

  * gluoncv.model_zoo densenet121 / resnet18_v2 `.features`      (call sites train.py:204, evaluate.py:125)
  * utils/layers.py:26-48                TimeDistributed ('reshape' style)
  * models/vision/definitions.py:10-33   FrameModel
  * models/vision/definitions.py:36-72   TemporalPooling
  * models/vision/definitions.py:75-110  CNNRNN (mx.gluon.rnn.GRU/LSTM, layout NTC, bidirectional)

following SURVEY.md Appendix A.2/A.3.  tests/ pin it against two independent implementations available in
this image (torchvision's densenet121 graph and torch.nn.GRU/LSTM) and against committed fixtures.
"""
import math

import torch
import torch.nn.functional as F

from tennis_b200.synthetic import (BOTTLENECK, DENSE_CFG, GROWTH, PARAM_SHAPES, RESNET_CH,  # noqa: F401  (shared seeded
                                   densenet121_param_shapes, flatten_params, normalize_u8,     # generators: the product-side
                                   resnet18_v2_param_shapes, rnn_param_shapes, synthetic_frames,  # bench uses the same tensors
                                   synthetic_params, synthetic_rnn_params)                       # without importing the oracle)

BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------------------------------
_BN_TRAINING = [False]  # set by the `training=` argument of the feature functions below


def _bn(x, p, prefix):
    if _BN_TRAINING[0]:
        # training mode (under autograd.record()): normalise with the batch statistics (biased variance); the running estimates
        # are updated as running = 0.9 * running + 0.1 * batch with the BIASED batch variance (MXNet)
        if p.get("_update_running", False):
            with torch.no_grad():
                mu = x.mean(dim=(0, 2, 3))
                var = x.var(dim=(0, 2, 3), unbiased=False)
                p[prefix + ".running_mean"].mul_(0.9).add_(0.1 * mu)
                p[prefix + ".running_var"].mul_(0.9).add_(0.1 * var)
        return F.batch_norm(x, None, None, p[prefix + ".gamma"], p[prefix + ".beta"], training=True, eps=BN_EPS)
    # inference BatchNorm with running statistics: (x - mean) / sqrt(var + eps) * gamma + beta
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".gamma"],
                        p[prefix + ".beta"], training=False, eps=BN_EPS)


def _with_mode(fn):
    def wrapped(x, p, training=False):
        prev = _BN_TRAINING[0]
        _BN_TRAINING[0] = bool(training)
        try:
            return fn(x, p)
        finally:
            _BN_TRAINING[0] = prev
    wrapped.__doc__ = fn.__doc__
    return wrapped


def densenet121_features(x, p):
    """x: (N,3,H,W) -> (N, 1024*ph*pw).  [UPSTREAM gluoncv densenet] as summarised in SURVEY.md §8a V1 / A.2:
    conv7x7/2 -> BN -> relu -> maxpool3/2/1 -> dense blocks [6,12,24,16] (BN-relu-conv1x1(128)-BN-relu-conv3x3(32),
    concat [input, new]) with transitions BN-relu-conv1x1(C/2)-avgpool2 -> BN -> relu -> AvgPool2D(7) (stride 7,
    'valid', i.e. NOT global) -> Flatten (channel-major)."""
    x = F.conv2d(x, p["conv0.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, p, "bn0"))
    x = F.max_pool2d(x, 3, 2, 1)
    for b, nl in enumerate(DENSE_CFG):
        for l in range(nl):
            pre = "block%d.layer%d" % (b + 1, l + 1)
            y = F.conv2d(F.relu(_bn(x, p, pre + ".bn1")), p[pre + ".conv1.weight"])
            y = F.conv2d(F.relu(_bn(y, p, pre + ".bn2")), p[pre + ".conv2.weight"], padding=1)
            x = torch.cat([x, y], dim=1)
        if b < 3:
            pre = "trans%d" % (b + 1)
            x = F.conv2d(F.relu(_bn(x, p, pre + ".bn")), p[pre + ".conv.weight"])
            x = F.avg_pool2d(x, 2, 2)
    x = F.relu(_bn(x, p, "bn5"))
    x = F.avg_pool2d(x, 7)  # pool_size=7 => stride 7, floor ('valid'): 7x7 -> 1x1, 16x16 -> 2x2
    return x.flatten(1)


def resnet18_v2_features(x, p):
    """[UPSTREAM gluoncv resnetv2] SURVEY.md §8a V2 / A.2."""
    x = _bn(x, p, "bn_data")
    x = F.conv2d(x, p["conv0.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, p, "bn0"))
    x = F.max_pool2d(x, 3, 2, 1)
    cin = 64
    for s, c in enumerate(RESNET_CH):
        for b in range(2):
            pre = "stage%d.block%d" % (s + 1, b + 1)
            stride = 2 if (b == 0 and s > 0) else 1
            residual = x
            y = F.relu(_bn(x, p, pre + ".bn1"))
            if (pre + ".downsample.weight") in p:
                residual = F.conv2d(y, p[pre + ".downsample.weight"], stride=stride)
            y = F.conv2d(y, p[pre + ".conv1.weight"], stride=stride, padding=1)
            y = F.relu(_bn(y, p, pre + ".bn2"))
            y = F.conv2d(y, p[pre + ".conv2.weight"], padding=1)
            x = y + residual
            cin = c
    x = F.relu(_bn(x, p, "bn_final"))
    return x.mean(dim=(2, 3))  # GlobalAvgPool2D + Flatten


densenet121_features = _with_mode(densenet121_features)
resnet18_v2_features = _with_mode(resnet18_v2_features)
FEATURES = {"densenet121": densenet121_features, "resnet18_v2": resnet18_v2_features}


# ----------------------------------------------------------------------------------------------------------------------
def dense(x, weight, bias=None):
    """gluon nn.Dense(flatten=True): y = x W^T + b with W (out,in)."""
    y = x.reshape(x.shape[0], -1) @ weight.t()
    return y if bias is None else y + bias


def time_distributed(fn, x):
    """utils/layers.py:38-46 ('reshape' style): fold (B,T,...) -> (B*T,...), apply, unfold to (B,T,...)."""
    B, T = x.shape[:2]
    y = fn(x.reshape((B * T,) + tuple(x.shape[2:])))
    return y.reshape((B, T) + tuple(y.shape[1:]))


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """A.3: gates [r,z,n]; n = tanh(i2h_n + r*h2h_n); h' = (1-z)*n + z*h."""
    H = h.shape[-1]
    i2h = x @ w_ih.t() + b_ih
    h2h = h @ w_hh.t() + b_hh
    r = torch.sigmoid(i2h[:, :H] + h2h[:, :H])
    z = torch.sigmoid(i2h[:, H:2 * H] + h2h[:, H:2 * H])
    n = torch.tanh(i2h[:, 2 * H:] + r * h2h[:, 2 * H:])
    return (1 - z) * n + z * h


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """A.3: gates [i,f,g,o]."""
    H = h.shape[-1]
    g = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i, f = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H])
    gg, o = torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
    c2 = f * c + i * gg
    return o * torch.tanh(c2), c2


def birnn_layer(x, p, cell="gru", H=128):
    """mx.gluon.rnn.GRU/LSTM(H, layout='NTC', bidirectional=True), zero initial state (A.3):
    out[:, t] = concat(h_fwd[t], h_bwd[t])."""
    B, T, _ = x.shape
    outs = []
    for d, order in (("l0", range(T)), ("r0", range(T - 1, -1, -1))):
        if (d + "_i2h_weight") not in p:
            continue
        w_ih, w_hh, b_ih, b_hh = (p[d + s] for s in ("_i2h_weight", "_h2h_weight", "_i2h_bias", "_h2h_bias"))
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        ys = [None] * T
        for t in order:
            if cell == "gru":
                h = gru_cell(x[:, t], h, w_ih, w_hh, b_ih, b_hh)
            else:
                h, c = lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            ys[t] = h
        outs.append(torch.stack(ys, dim=1))
    return torch.cat(outs, dim=2)


def frame_model(x, arch, p, classes_w=None, classes_b=None):
    """definitions.py:27-33: backbone -> optional Dense."""
    f = FEATURES[arch](x, p)
    return f if classes_w is None else dense(f, classes_w, classes_b)


def temporal_pooling(x, feature_fn, pool, classes_w=None, classes_b=None, feats=False):
    """definitions.py:63-72."""
    if not feats:
        x = time_distributed(feature_fn, x)
    x = x.mean(dim=1) if pool == "mean" else x.max(dim=1).values
    return x if classes_w is None else dense(x, classes_w, classes_b)


def cnnrnn(x, feature_fn, rnn_p, cell, H, classes_w=None, classes_b=None, feats=False):
    """definitions.py:103-110: td -> bi-RNN -> max over time -> Dense."""
    if not feats:
        x = time_distributed(feature_fn, x)
    y = birnn_layer(x, rnn_p, cell, H)
    y = y.max(dim=1).values
    return y if classes_w is None else dense(y, classes_w, classes_b)
