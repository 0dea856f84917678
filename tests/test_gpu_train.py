"""Head-only training step (V7): gradients and optimiser updates against torch autograd on the CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _torch_model(cell, D, H, p, cw, cb):
    mod = (torch.nn.GRU if cell == "gru" else torch.nn.LSTM)(D, H, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for d, suf in (("l0", ""), ("r0", "_reverse")):
            getattr(mod, "weight_ih_l0" + suf).copy_(p[d + "_i2h_weight"])
            getattr(mod, "weight_hh_l0" + suf).copy_(p[d + "_h2h_weight"])
            getattr(mod, "bias_ih_l0" + suf).copy_(p[d + "_i2h_bias"])
            getattr(mod, "bias_hh_l0" + suf).copy_(p[d + "_h2h_bias"])
    lin = torch.nn.Linear(2 * H, cw.shape[0])
    with torch.no_grad():
        lin.weight.copy_(cw)
        lin.bias.copy_(cb)
    return mod, lin


@pytest.mark.parametrize("cell", ["gru", "lstm"])
def test_head_gradients_match_torch_autograd(cell):
    from oracle import vision as O
    from tennis_b200 import autograd
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss
    from tennis_b200.models.vision.definitions import CNNRNN
    B, T, D, H, C = 6, 9, 64, 128, 11
    p = O.synthetic_rnn_params(cell, D, H, seed=4321)
    g = torch.Generator().manual_seed(1)
    cw = (torch.rand(C, 2 * H, generator=g) * 2 - 1) * 0.3
    cb = torch.randn(C, generator=g) * 0.1
    x = torch.randn(B, T, D, generator=g).relu()
    y = torch.randint(0, C, (B,), generator=g)
    # torch reference
    mod, lin = _torch_model(cell, D, H, p, cw, cb)
    out_ref = lin(mod(x)[0].max(dim=1).values)
    loss_ref = torch.nn.functional.cross_entropy(out_ref, y, reduction="none")
    loss_ref.sum().backward()
    # ours
    dev = torch.device("cuda", 0)
    model = CNNRNN(None, C, hidden_size=H, type=cell)
    model.initialize(ctx=dev)
    for k, v in p.items():
        prm = model.rnn._reg_params[k]
        prm.shape, prm._data = tuple(v.shape), v.to(dev)
        prm._version += 1
    model.classes.weight.shape, model.classes.weight._data = tuple(cw.shape), cw.to(dev)
    model.classes.bias._data = cb.to(dev)
    loss_fn = SoftmaxCrossEntropyLoss()
    with autograd.record():
        out = model(x.to(dev))
        loss = loss_fn(out, y.to(dev))
    autograd.backward([loss])
    torch.cuda.synchronize()
    assert (loss.cpu() - loss_ref.detach()).abs().max().item() < 2e-3
    ref = {"classes.weight": lin.weight.grad, "classes.bias": lin.bias.grad}
    for d, suf in (("l0", ""), ("r0", "_reverse")):
        ref["rnn.%s_i2h_weight" % d] = getattr(mod, "weight_ih_l0" + suf).grad
        ref["rnn.%s_h2h_weight" % d] = getattr(mod, "weight_hh_l0" + suf).grad
        ref["rnn.%s_i2h_bias" % d] = getattr(mod, "bias_ih_l0" + suf).grad
        ref["rnn.%s_h2h_bias" % d] = getattr(mod, "bias_hh_l0" + suf).grad
    params = model.collect_params()
    for k, gref in ref.items():
        got = params[k].grad().cpu()
        scale = max(1e-3, gref.abs().max().item())
        err = (got - gref).abs().max().item()
        assert err < 2e-2 * scale, "%s: err %g scale %g" % (k, err, scale)  # forward projection runs in bf16


def test_sgd_and_adam_updates_match_reference_formulas():
    from tennis_b200 import ops
    g = torch.Generator().manual_seed(3)
    w0, gr = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    lr, mu, wd, rs = 0.01, 0.9, 1e-4, 1.0 / 64
    w, m = w0.clone().cuda(), torch.zeros(1000).cuda()
    wr, mr = w0.clone(), torch.zeros(1000)
    for _ in range(3):
        ops.sgd_mom_update(w, gr.cuda(), m, lr, mu, wd, rs)
        gp = rs * gr + wd * wr
        mr = mu * mr - lr * gp
        wr = wr + mr
    assert (w.cpu() - wr).abs().max().item() < 1e-6
    w, m, v = w0.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    wr, mr, vr = w0.clone().double(), torch.zeros(1000).double(), torch.zeros(1000).double()
    b1, b2, eps = 0.9, 0.999, 1e-8
    for t in range(1, 4):
        ops.adam_update(w, gr.cuda(), m, v, 1e-3, b1, b2, eps, 0.0, 1.0, t)
        gp = gr.double()
        mr = b1 * mr + (1 - b1) * gp
        vr = b2 * vr + (1 - b2) * gp * gp
        lr_t = 1e-3 * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
        wr = wr - lr_t * mr / (vr.sqrt() + eps)
    assert (w.cpu().double() - wr).abs().max().item() < 1e-6


def test_training_reduces_loss_on_features():
    """A few SGD steps of the published configuration (features -> BiGRU -> max -> Dense) lower the loss."""
    from tennis_b200 import autograd
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss, Trainer, Uniform
    from tennis_b200.models.vision.definitions import CNNRNN
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    B, T, D = 32, 8, 64
    x = torch.randn(B, T, D, generator=g).relu().to(dev)
    y = (x.mean(dim=(1, 2)) > x.mean()).long() * 3  # two learnable classes
    model = CNNRNN(None, 11, hidden_size=128, type="gru")
    model.initialize(init=Uniform(0.07), ctx=dev)
    trainer, loss_fn = None, SoftmaxCrossEntropyLoss()
    losses = []
    for it in range(30):
        with autograd.record():
            loss = loss_fn(model(x), y)
        if trainer is None:
            trainer = Trainer(model.collect_params(), "sgd", {"learning_rate": 0.5, "momentum": 0.9, "wd": 1e-4})
        autograd.backward([loss])
        trainer.step(B)
        losses.append(loss.mean().item())
    assert losses[-1] < 0.5 * losses[0], losses


def test_dense_flatten_false_backward_and_dropout_guard():
    """Dense(flatten=False) on a (B,T,C) input under autograd.record() carries its backward (it silently dropped the
    gradients); Dropout(rate>0) refuses to run recorded outside the captioner's training graph."""
    import torch
    from tennis_b200 import autograd
    from tennis_b200.gluon import Dense, Dropout
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 5, 8, generator=g)
    w = torch.randn(4, 8, generator=g)
    b = torch.randn(4, generator=g)
    d = Dense(4, in_units=8, flatten=False)
    d.initialize(ctx=torch.device("cuda", 0))
    d.weight.set_data(w)
    d.bias.set_data(b)
    xc = x.cuda()
    with autograd.record():
        y = d(xc)
    head = torch.randn(3, 5, 4, generator=g)
    autograd.backward([y], [head.cuda()])
    xr = x.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    (torch.nn.functional.linear(xr, wr, br) * head).sum().backward()
    assert y.shape == (3, 5, 4)
    assert (d.weight.grad().cpu() - wr.grad).abs().max().item() < 1e-5
    assert (d.bias.grad().cpu() - br.grad).abs().max().item() < 1e-5
    with pytest.raises(NotImplementedError):
        with autograd.record():
            Dropout(0.3)(xc)
    assert Dropout(0.3)(xc) is xc  # inference: identity


def test_device_metrics_match_host_metrics_bit_exact():
    """tn_metrics_update (confusion matrix, top-1 / top-k hits accumulated on the device) against metrics/vision.py on the host,
    including exact ties (lowest class index first, like a stable argsort / numpy argmax)."""
    import torch
    from tennis_b200.metrics.device import DeviceMetrics
    from tennis_b200.metrics.vision import PRF1, Accuracy
    names = ["c%d" % i for i in range(11)]
    g = torch.Generator().manual_seed(4)
    dm = DeviceMetrics(names, top_k=5, device="cuda:0")
    acc, top5, prf = Accuracy(), Accuracy("top5", top_k=5), PRF1(label_names=names)
    for n in (64, 1, 37):
        logits = torch.randn(n, 11, generator=g)
        logits[::3] = (logits[::3] * 2).round() / 2  # many exact ties
        labels = torch.randint(0, 11, (n,), generator=g)
        dm.update(labels.cuda(), logits.cuda())
        for m in (acc, top5, prf):
            m.update([labels], [logits])
    a, t, p = dm.finish()
    assert (a.hit, a.n) == (acc.hit, acc.n) and (t.hit, t.n) == (top5.hit, top5.n)
    assert (p.mat == prf.mat).all() and (p.scores == prf.scores).all()
    assert p.get() == prf.get()
