"""A minimal tape for the recorded training steps (mxnet.autograd surface used at train.py:415-421 and train_gnmt.py:330-334:
`with ag.record(): out = model(x); loss = loss_fn(out, y)` then `ag.backward(losses)`).

Tensors stay plain torch CUDA tensors; an op executed while recording tags its output with `_tn_node =
(backward_fn, input_tensor)`.  backward() walks that chain from the loss: loss -> classifier Dense -> fused (bi)RNN +
max-over-time -> TimeDistributed unfold -> CNN training graph (models/vision/train_graph.py, which keeps its own internal tape
for the branching inside the network), or loss -> NMTModel training graph (models/captioning/train_graph.py).  A node whose
input is a leaf (pre-extracted features, frozen backbone, input frames) ends the walk."""
import contextlib
import threading

import torch

_state = threading.local()


def is_recording():
    return getattr(_state, "recording", False)


@contextlib.contextmanager
def record(train_mode=True):
    prev = is_recording()
    _state.recording = True
    try:
        yield
    finally:
        _state.recording = prev


def tag(out, backward_fn, inp):
    out._tn_node = (backward_fn, inp)
    return out


def backward(heads, head_grads=None):
    """ag.backward(list_of_losses): seeds every head with ones (A.8) and accumulates into Parameter gradients."""
    if not isinstance(heads, (list, tuple)):
        heads = [heads]
    for i, h in enumerate(heads):
        grad = torch.ones_like(h) if head_grads is None else head_grads[i]
        t = h
        while t is not None and getattr(t, "_tn_node", None) is not None:
            fn, inp = t._tn_node
            grad = fn(grad)
            t = inp
            if grad is None:
                break
