"""Captioner training step (G9): the recorded forward, MaskedSoftmaxCELoss, backward and Adam step through the C ABI against
torch autograd over the CPU oracle (oracle/captioning.py) with the same seeded weights and inputs.  fp32 kernels: logits 1e-4,
gradients 2e-4 relative to the largest gradient entry of each tensor."""
import pytest
import torch

from test_gpu_gnmt import _build

pytestmark = pytest.mark.gpu


def _oracle_loss_and_grads(p, x, tgt, vl, tvl, cell, H, scale):
    from oracle import captioning as C
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = C.nmt_forward(q, x, tgt[:, :-1], vl, tvl - 1, cell=cell, H=H)
    loss_vec = C.masked_softmax_ce(out, tgt[:, 1:], tvl - 1)
    loss = loss_vec.mean() * scale
    loss.backward()
    return out.detach(), loss_vec.detach(), {k: v.grad for k, v in q.items()}


@pytest.mark.parametrize("cell", ["lstm", "gru"])
@pytest.mark.parametrize("with_vl", [True, False])
def test_gnmt_training_gradients_match_oracle_autograd(cell, with_vl):
    from oracle import captioning as C
    from tennis_b200 import autograd
    from tennis_b200.gluon import MaskedSoftmaxCELoss
    H, D, E, V = 32, 48, 20, 37
    model, p, _ = _build(cell, H, D, E, V, 0.3)
    B, T, Tt = 5, 9, 7
    x, vl = C.synthetic_sources(B, T, D, seed=3)
    g = torch.Generator().manual_seed(5)
    tgt = torch.randint(0, V, (B, Tt), generator=g).float()
    tvl = torch.tensor([7., 6., 4., 2., 7.])
    if not with_vl:
        vl = torch.full((B,), float(T))
        tvl = torch.full((B,), float(Tt))
    scale = float((Tt - 1) / (tvl - 1).mean())  # train_gnmt.py:333
    ref_out, ref_loss, ref_grads = _oracle_loss_and_grads(p, x, tgt, vl, tvl, cell, H, scale)

    loss_fn = MaskedSoftmaxCELoss()
    xs, ts, vls, tvls = x.cuda(), tgt.cuda(), vl.cuda(), tvl.cuda()
    with autograd.record():
        out, _ = model(xs, ts[:, :-1], vls, tvls - 1)
        loss = loss_fn(out, ts[:, 1:], tvls - 1)
    autograd.backward([loss], [torch.full_like(loss, scale / B)])
    torch.cuda.synchronize()
    assert (out.cpu() - ref_out).abs().max().item() < 1e-4
    assert (loss.cpu() - ref_loss).abs().max().item() < 1e-5
    params = model.collect_params()
    worst = ("", 0.0)
    for k, gref in ref_grads.items():
        got = params[k].grad().cpu()
        assert got.shape == gref.shape, k
        rel = (got - gref).abs().max().item() / max(gref.abs().max().item(), 1e-6)
        if rel > worst[1]:
            worst = (k, rel)
        assert rel < 2e-4, "%s: relative gradient error %.3g (max |g| %.3g)" % (k, rel, gref.abs().max().item())
    print("gnmt %s vl=%s: worst relative gradient error %.2e at %s" % (cell, with_vl, worst[1], worst[0]))
    # the inference engines (tensor-core projections) and the fp32 training forward agree within the stated 1e-3
    inf_out, _ = model(xs, ts[:, :-1], vls, tvls - 1)
    assert (inf_out - out).abs().max().item() < 1e-3


def test_gnmt_adam_steps_reduce_loss_with_dropout():
    """train_gnmt.py:310,330-337 loop on one synthetic batch: Adam(lr 1e-3), dropout 0.2 (the flag default); the loss must fall
    and stay finite, and a step without dropout must reproduce the oracle's Adam update."""
    from oracle import captioning as C
    from tennis_b200 import autograd
    from tennis_b200.gluon import MaskedSoftmaxCELoss, Trainer
    cell, H, D, E, V = "lstm", 32, 48, 20, 37
    model, p, _ = _build(cell, H, D, E, V, 0.3)
    B, T, Tt = 6, 8, 6
    x, vl = C.synthetic_sources(B, T, D, seed=9)
    g = torch.Generator().manual_seed(1)
    tgt = torch.randint(4, V, (B, Tt), generator=g).float()
    tvl = torch.tensor([6., 5., 6., 3., 4., 6.])
    scale = float((Tt - 1) / (tvl - 1).mean())
    xs, ts, vls, tvls = x.cuda(), tgt.cuda(), vl.cuda(), tvl.cuda()
    loss_fn = MaskedSoftmaxCELoss()
    trainer = Trainer(model.collect_params(), 'adam', {'learning_rate': 1e-3})
    # step 1 without dropout against the oracle
    _, _, ref_grads = _oracle_loss_and_grads(p, x, tgt, vl, tvl, cell, H, scale)
    with autograd.record():
        out, _ = model(xs, ts[:, :-1], vls, tvls - 1)
        loss = loss_fn(out, ts[:, 1:], tvls - 1)
    autograd.backward([loss], [torch.full_like(loss, scale / B)])
    trainer.step(1)
    k = "decoder.rnn_cells.0.i2h_weight"
    gk = ref_grads[k]
    m, v = 0.1 * gk, 0.001 * gk * gk
    lr_t = 1e-3 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    ref_w = p[k] - lr_t * m / (v.sqrt() + 1e-8)
    got_w = model.collect_params()[k].data().cpu()
    big = gk.abs() > 1e-3 * gk.abs().max()  # Adam's g/|g| amplifies noise on near-zero gradients
    assert (got_w - ref_w)[big].abs().max().item() < 2e-5
    first = float(loss.mean().item() * scale)
    # now with dropout, a few more steps
    model.encoder._dropout = 0.2
    model.decoder._dropout = 0.2
    last = None
    for _ in range(25):
        with autograd.record():
            out, _ = model(xs, ts[:, :-1], vls, tvls - 1)
            loss = loss_fn(out, ts[:, 1:], tvls - 1)
        autograd.backward([loss], [torch.full_like(loss, scale / B)])
        trainer.step(1)
        last = float(loss.mean().item() * scale)
        assert last == last and abs(last) < 1e4
    print("loss %.4f -> %.4f after 26 Adam steps" % (first, last))
    assert last < first


@pytest.mark.parametrize("cell", ["lstm", "gru"])
def test_gnmt_source_feature_gradient_matches_oracle_autograd(cell):
    """train_gnmt.py:150-170 with a trainable CNN as `src_embed`: the training graph hands d(loss)/d(source features) to whatever
    produced the source (here a recorded leaf that captures it); rows past valid_length get exactly zero."""
    from oracle import captioning as C
    from tennis_b200 import autograd
    from tennis_b200.gluon import MaskedSoftmaxCELoss
    H, D, E, V = 32, 48, 20, 37
    model, p, _ = _build(cell, H, D, E, V, 0.3)
    B, T, Tt = 5, 9, 7
    x, vl = C.synthetic_sources(B, T, D, seed=3)
    tgt = torch.randint(0, V, (B, Tt), generator=torch.Generator().manual_seed(5)).float()
    tvl = torch.tensor([7., 6., 4., 2., 7.])
    xr = x.clone().requires_grad_(True)
    out = C.nmt_forward({k: v.clone() for k, v in p.items()}, xr, tgt[:, :-1], vl, tvl - 1, cell=cell, H=H)
    C.masked_softmax_ce(out, tgt[:, 1:], tvl - 1).sum().backward()
    got = []
    xs = x.cuda()
    autograd.tag(xs, lambda g: got.append(g) or None, None)   # stands for TimeDistributed(CNN)'s output
    with autograd.record():
        o, _ = model(xs, tgt[:, :-1].cuda(), vl.cuda(), tvl.cuda() - 1)
        loss = MaskedSoftmaxCELoss()(o, tgt[:, 1:].cuda(), tvl.cuda() - 1)
    autograd.backward([loss])
    assert len(got) == 1 and tuple(got[0].shape) == (B, T, D)
    d = got[0].cpu()
    assert (d - xr.grad).abs().max().item() < 2e-4 * xr.grad.abs().max().item()
    for b in range(B):
        assert (d[b, int(vl[b]):] == 0).all()
