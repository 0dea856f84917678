"""A minimal Gluon-shaped `Block` runtime: just the surface the reference scripts call on their models
(SURVEY.md §8b): `model(x)`, `.initialize`, `.collect_params()` (+ `.values()`, `.grad_req`, `.reset_ctx`),
`.hybridize`, `.summary`, `.save_parameters` / `.load_parameters`.  Tensors are torch CUDA tensors used as
containers; all arithmetic happens in libtennis_b200.so (see ops.py) — a Block called with CPU tensors raises.
"""
import collections
import math
import os

import numpy as np
import torch

from . import _lib


def gpu(i=0):
    return torch.device("cuda", i)


def cpu():
    return torch.device("cpu")


class Parameter(object):
    def __init__(self, name, shape=None, init="uniform", grad_req="write"):
        self.name = name
        self.shape = tuple(shape) if shape is not None else None
        self.init = init
        self.grad_req = grad_req
        self.lr_mult = 1.0
        self.wd_mult = 1.0
        self._data = None
        self._grad = None
        self._version = 0

    # -- Gluon-like accessors
    def data(self, ctx=None):
        if self._data is None:
            raise RuntimeError("Parameter '%s' has not been initialized" % self.name)
        return self._data

    def grad(self, ctx=None):
        if self._grad is None:
            raise RuntimeError("Parameter '%s' has no gradient (grad_req=%s, backward not run?)" % (self.name, self.grad_req))
        return self._grad

    def zero_grad(self):
        self._grad = None

    def _accumulate_grad(self, g):
        if self.grad_req == 'null':
            return
        self._grad = g.clone() if self._grad is None else self._grad + g

    def _bump(self):
        self._version += 1

    def list_ctx(self):
        return [] if self._data is None else [self._data.device]

    def set_data(self, t):
        t = torch.as_tensor(t, dtype=torch.float32)
        if self.shape is not None and all(s > 0 for s in self.shape) and tuple(t.shape) != self.shape:
            raise ValueError("Parameter '%s' shape %s does not match %s" % (self.name, self.shape, tuple(t.shape)))
        dev = self._data.device if self._data is not None else t.device
        self.shape = tuple(t.shape)
        self._data = t.detach().to(dev).contiguous().clone()
        self._version += 1

    def reset_ctx(self, ctx):
        if self._data is not None:
            self._data = self._data.to(_one_ctx(ctx))
            self._version += 1

    def initialize(self, init=None, ctx=None, force_reinit=False, generator=None):
        if self._data is not None and not force_reinit:
            return
        if self.shape is None or any(s <= 0 for s in self.shape):
            self._deferred = (init, ctx, generator)  # shape known at first forward
            return
        self._data = _init_tensor(self, init, generator).to(_one_ctx(ctx))
        self._version += 1

    def _finish_deferred(self, shape):
        self.shape = tuple(shape)
        init, ctx, gen = getattr(self, "_deferred", (None, None, None))
        self._data = _init_tensor(self, init, gen).to(_one_ctx(ctx))
        self._version += 1


def _one_ctx(ctx):
    if ctx is None:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if isinstance(ctx, (list, tuple)):
        ctx = ctx[0]
    return torch.device(ctx)


class Uniform(object):
    """mx.init.Uniform(scale): weights ~ U(-scale, scale); biases 0; BN gamma 1, beta 0, mean 0, var 1."""

    def __init__(self, scale=0.07):
        self.scale = scale


def _init_tensor(param, init, generator):
    kind = param.init
    shape = param.shape
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "lstmbias":  # mx.init.LSTMBias(forget_bias=1.0): zeros, then arr[n/4 : n/2] = 1  (Appendix C #14)
        t = torch.zeros(shape)
        n = shape[0]
        t[n // 4: n // 2] = 1.0
        return t
    scale = init.scale if isinstance(init, Uniform) else 0.07
    return (torch.rand(shape, generator=generator) * 2 - 1) * scale


class ParameterDict(collections.OrderedDict):
    def reset_ctx(self, ctx):
        for p in self.values():
            p.reset_ctx(ctx)

    def initialize(self, init=None, ctx=None, force_reinit=False):
        for p in self.values():
            p.initialize(init, ctx, force_reinit)

    def setattr(self, name, value):
        for p in self.values():
            setattr(p, name, value)


class Block(object):
    """Base class: registers child Blocks / Parameters assigned as attributes, in assignment order."""

    def __init__(self, prefix=None, params=None):
        object.__setattr__(self, "_children", collections.OrderedDict())
        object.__setattr__(self, "_reg_params", collections.OrderedDict())
        self._prefix = prefix or ""
        self._hybridized = False

    def __setattr__(self, name, value):
        if isinstance(value, Block):
            self._children[name] = value
        elif isinstance(value, Parameter):
            self._reg_params[name] = value
        elif name in getattr(self, "_children", {}):  # e.g. self.classes = None after being a Block
            del self._children[name]
        object.__setattr__(self, name, value)

    def name_scope(self):
        return _NullCtx()

    # -- structural parameter names ("td.model.conv0.weight"), like Gluon's save_parameters
    def _collect(self, prefix, out):
        for n, p in self._reg_params.items():
            out[prefix + n] = p
        for n, c in self._children.items():
            c._collect(prefix + n + ".", out)

    def collect_params(self, select=None):
        out = ParameterDict()
        self._collect("", out)
        if select:
            import re
            pat = re.compile(select)
            out = ParameterDict((k, v) for k, v in out.items() if pat.match(k))
        return out

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        gen = torch.Generator().manual_seed(int(os.environ.get("TN_INIT_SEED", "0")))
        for p in self.collect_params().values():
            p.initialize(init, ctx, force_reinit, generator=gen)

    def hybridize(self, active=True, **kwargs):
        self._hybridized = bool(active)
        for c in self._children.values():
            c.hybridize(active, **kwargs)

    def cast(self, dtype):
        return self

    def __call__(self, *args):
        return self.forward(*args)

    def forward(self, *args):
        raise NotImplementedError

    def summary(self, *inputs):
        """Print a per-child table of output shapes and parameter counts (train.py:252-261)."""
        rows = []

        def hook(name, blk):
            orig = blk.forward

            def wrapped(*a):
                out = orig(*a)
                o = out[0] if isinstance(out, (tuple, list)) else out
                n = sum(int(np.prod(p.shape)) for p in blk._reg_params.values() if p.shape)
                rows.append((name or type(blk).__name__, type(blk).__name__, tuple(o.shape) if hasattr(o, "shape") else "-", n))
                return out
            object.__setattr__(blk, "forward", wrapped)
            return orig

        saved = []

        def walk(prefix, blk):
            saved.append((blk, hook(prefix, blk)))
            for n, c in blk._children.items():
                walk(prefix + ("." if prefix else "") + n, c)
        walk("", self)
        try:
            self(*inputs)
        finally:
            for blk, orig in saved:
                object.__delattr__(blk, "forward")
        print("-" * 100)
        print("%-48s %-22s %-20s %s" % ("Layer", "Type", "Output shape", "Params"))
        print("=" * 100)
        for r in reversed(rows):
            print("%-48s %-22s %-20s %d" % r)
        total = sum(int(np.prod(p.shape)) for p in self.collect_params().values() if p.shape)
        print("=" * 100)
        print("Total params: %d" % total)
        print("-" * 100)

    # -- checkpoints: numpy .npz container keyed by structural names (MXNet .params codec: see params_io.py)
    def save_parameters(self, filename):
        from . import params_io
        params_io.save(filename, {k: p.data().detach().cpu().numpy() for k, p in self.collect_params().items()
                                  if p._data is not None})

    def load_parameters(self, filename, ctx=None, allow_missing=False, ignore_extra=False):
        from . import params_io
        loaded = params_io.load(filename)
        params = self.collect_params()
        from .gluoncv_names import translate_checkpoint_keys
        loaded = translate_checkpoint_keys(loaded, params)  # checkpoints written by the reference use GluonCV's child indices
        if not allow_missing:
            missing = [k for k in params if k not in loaded]
            if missing:
                raise KeyError("Parameters missing in '%s': %s" % (filename, missing[:5]))
        if not ignore_extra:
            extra = [k for k in loaded if k not in params]
            if extra:
                raise KeyError("Parameters in '%s' not present in the Block: %s" % (filename, extra[:5]))
        dev = _one_ctx(ctx)
        for k, arr in loaded.items():
            if k in params:
                p = params[k]
                if p._data is None:
                    p.shape = tuple(arr.shape)
                    p._data = torch.zeros(arr.shape, device=dev)
                p.set_data(torch.from_numpy(np.ascontiguousarray(arr)))
                if ctx is not None:
                    p.reset_ctx(dev)


HybridBlock = Block


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class HybridSequential(Block):
    def __init__(self, prefix=None, params=None):
        super(HybridSequential, self).__init__(prefix, params)
        self._n = 0

    def add(self, *blocks):
        for b in blocks:
            setattr(self, str(self._n), b)
            self._n += 1

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._children[str(i if i >= 0 else self._n + i)]

    def __iter__(self):
        return iter(self._children.values())

    def forward(self, x):
        for b in self._children.values():
            x = b(x)
        return x


class Dropout(Block):
    """Inference / rate-0 dropout is the identity (train_gnmt.py:190-192 uses Dropout(0.0) as src_embed)."""

    def __init__(self, rate=0.0, **kw):
        super(Dropout, self).__init__(**kw)
        self.rate = rate

    def forward(self, x):
        from . import autograd
        if self.rate > 0 and autograd.is_recording():
            # the captioner's training graph applies its dropout masks itself (models/captioning/train_graph.py); a generic
            # recorded use of this block would silently train WITHOUT dropout, so refuse instead
            raise NotImplementedError("Dropout(rate=%g) under autograd.record() is only implemented inside GNMTTrainGraph" % self.rate)
        return x


class Dense(Block):
    """gluon nn.Dense(units, flatten=True|False): y = x W^T + b, weight (units, in_units); in_units deferred."""

    def __init__(self, units, in_units=0, flatten=True, use_bias=True, **kw):
        super(Dense, self).__init__(**kw)
        self._units, self._flatten = units, flatten
        self.weight = Parameter("weight", (units, in_units))
        self.bias = Parameter("bias", (units,), init="zeros") if use_bias else None

    def forward(self, x):
        from . import ops
        x_orig = x  # the tape links tensors by object identity; reshape() below makes a new object
        lead = None
        if not self._flatten and x.dim() > 2:
            lead = x.shape[:-1]
            x = x.reshape(-1, x.shape[-1])
        else:
            x = x.reshape(x.shape[0], -1)
        if self.weight._data is None:
            self.weight._finish_deferred((self._units, x.shape[1]))
            self.weight.reset_ctx(x.device)
        y = ops.dense(x, self.weight.data(), None if self.bias is None else self.bias.data())
        from . import autograd
        out = y if lead is None else y.reshape(tuple(lead) + (self._units,))
        if autograd.is_recording():
            xin = x.contiguous().float()

            def bwd(dy, xin=xin, self=self):
                # flatten=False folds the leading axes into rows: the same dense backward on the folded view
                dx, dw, db = ops.dense_backward(xin, self.weight.data(), dy.reshape(xin.shape[0], -1).contiguous())
                self.weight._accumulate_grad(dw)
                if self.bias is not None:
                    self.bias._accumulate_grad(db)
                return dx.reshape(x_orig.shape)
            autograd.tag(out, bwd, x_orig)
        return out


class Embedding(Block):
    def __init__(self, input_dim, output_dim, **kw):
        super(Embedding, self).__init__(**kw)
        self.weight = Parameter("weight", (input_dim, output_dim))

    def forward(self, ids):
        # pure gather (indexing, no arithmetic): rows of the table selected by (float or int) token ids
        return self.weight.data()[ids.long()]


class SoftmaxCrossEntropyLoss(object):
    """gluon.loss.SoftmaxCrossEntropyLoss(): -log_softmax(pred)[label] per sample -> (B,)  (train.py:324, A.8)."""

    def __call__(self, pred, label):
        from . import autograd, ops
        loss, dlogits = ops.softmax_ce(pred, label, want_grad=autograd.is_recording())
        if autograd.is_recording():
            autograd.tag(loss, lambda g, d=dlogits: d * g.reshape(-1, 1), pred)
        return loss


class MaskedSoftmaxCELoss(object):
    """gluonnlp.loss.MaskedSoftmaxCELoss forward (train_gnmt.py:256): (B,T,V) logits, (B,T) labels, (B,) valid lengths -> (B,)."""

    def __call__(self, pred, label, valid_length):
        from . import autograd, ops
        if autograd.is_recording():
            from .models.captioning.train_graph import masked_softmax_ce
            loss, _ = masked_softmax_ce(pred, label, valid_length)
            autograd.tag(loss, lambda g, a=(pred, label, valid_length): masked_softmax_ce(a[0], a[1], a[2], head_grad=g)[1], pred)
            return loss
        return ops.masked_softmax_ce(pred, label, valid_length)


class Trainer(object):
    """gluon.Trainer(params, 'sgd'|'adam', {...}) as the scripts use it (train.py:298-299,424; train_gnmt.py:310,337):
    step(n) rescales the summed gradients by 1/n, applies wd to EVERY parameter (Gluon wd_mult = 1) and updates in place.
    With torch.distributed initialised, gradients are summed across ranks first (the reference's KVStore('device') sum)."""

    def __init__(self, params, optimizer, optimizer_params=None):
        self._params = [p for p in (params.values() if hasattr(params, "values") else params)]
        self._opt = optimizer.lower()
        if self._opt not in ("sgd", "adam"):
            raise ValueError("optimizer must be 'sgd' or 'adam'")
        op = dict(optimizer_params or {})
        self._lr = float(op.get("learning_rate", 0.01))
        self._momentum = float(op.get("momentum", 0.0))
        self._wd = float(op.get("wd", 0.0))
        self._b1, self._b2, self._eps = float(op.get("beta1", 0.9)), float(op.get("beta2", 0.999)), float(op.get("epsilon", 1e-8))
        self._state = {}
        self._t = 0

    @property
    def learning_rate(self):
        return self._lr

    def set_learning_rate(self, lr):
        self._lr = float(lr)

    def step(self, batch_size, ignore_stale_grad=False):
        from . import ops
        from .parallel import sum_gradients_across_ranks
        self._t += 1
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # every rank must enter the same all-reduce: a rank whose shard of the batch was empty (or that did not touch a
            # parameter) contributes zeros, like an idle context in the reference's split_and_load(even_split=False) loop
            live = [p for p in self._params if p.grad_req != 'null' and p._data is not None]
            for p in live:
                if p._grad is None:
                    p._grad = torch.zeros_like(p.data())
            # a rank that has not materialised a deferred-shape parameter yet would enter a smaller all-reduce and hang every
            # rank: compare the bucket sizes first (one 2-element all-reduce) and fail loudly instead
            n = sum(p._grad.numel() for p in live)
            chk = torch.tensor([n, -n], dtype=torch.int64, device=live[0]._grad.device if live else "cuda")
            dist.all_reduce(chk, op=dist.ReduceOp.MAX)
            if int(chk[0]) != -int(chk[1]):
                raise RuntimeError("Trainer.step: the ranks hold different parameter sets (%d gradient elements here, %d..%d over "
                                   "the ranks): run one forward on every rank before the first step (deferred-shape parameters)"
                                   % (n, -int(chk[1]), int(chk[0])))
        else:
            live = [p for p in self._params if p.grad_req != 'null' and p._grad is not None]
        for p in live:
            p._grad = p._grad.contiguous()
        sum_gradients_across_ranks([p._grad for p in live])  # no-op in a single process
        for p in live:
            g = p._grad
            w = p.data()
            st = self._state.get(id(p))
            if self._opt == "sgd":
                if st is None:
                    st = self._state[id(p)] = (torch.zeros_like(w),)
                ops.sgd_mom_update(w, g, st[0], self._lr, self._momentum, self._wd, 1.0 / batch_size)
            else:
                if st is None:
                    st = self._state[id(p)] = (torch.zeros_like(w), torch.zeros_like(w))
                ops.adam_update(w, g, st[0], st[1], self._lr, self._b1, self._b2, self._eps, self._wd, 1.0 / batch_size, self._t)
            p._bump()
            p._grad = None
