"""CPU-only: pin the oracle (a) against independent implementations shipped in this image (torchvision's DenseNet-121
graph, torch.nn.GRU / LSTM) and (b) against the committed golden fixtures (tests/golden, made by tools/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vision as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vision_oracle.npz")


def _torchvision_densenet_with(p):
    tv = pytest.importorskip("torchvision")
    m = tv.models.densenet121(weights=None).eval()
    sd = {}

    def bn(dst, src):
        sd[dst + ".weight"] = p[src + ".gamma"]
        sd[dst + ".bias"] = p[src + ".beta"]
        sd[dst + ".running_mean"] = p[src + ".running_mean"]
        sd[dst + ".running_var"] = p[src + ".running_var"]

    sd["features.conv0.weight"] = p["conv0.weight"]
    bn("features.norm0", "bn0")
    for b, nl in enumerate(O.DENSE_CFG):
        for l in range(nl):
            src = "block%d.layer%d" % (b + 1, l + 1)
            dst = "features.denseblock%d.denselayer%d" % (b + 1, l + 1)
            bn(dst + ".norm1", src + ".bn1")
            sd[dst + ".conv1.weight"] = p[src + ".conv1.weight"]
            bn(dst + ".norm2", src + ".bn2")
            sd[dst + ".conv2.weight"] = p[src + ".conv2.weight"]
        if b < 3:
            bn("features.transition%d.norm" % (b + 1), "trans%d.bn" % (b + 1))
            sd["features.transition%d.conv.weight" % (b + 1)] = p["trans%d.conv.weight" % (b + 1)]
    bn("features.norm5", "bn5")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert all(k.startswith("classifier") or k.endswith("num_batches_tracked") for k in missing), missing
    assert not unexpected
    return m


def test_densenet_oracle_matches_torchvision_graph():
    p = O.synthetic_params("densenet121", seed=1234)
    tvm = _torchvision_densenet_with(p)
    _, x = O.synthetic_frames(1, 224, seed=5)
    with torch.no_grad():
        mine = O.densenet121_features(x, p)
        ref = torch.nn.functional.avg_pool2d(torch.relu(tvm.features(x)), 7).flatten(1)
    assert mine.shape == (1, 1024)
    assert torch.allclose(mine, ref, rtol=1e-4, atol=1e-4), (mine - ref).abs().max()


def test_feature_width_follows_crop_size():
    """AvgPool2D(7) is stride 7 / valid, not global: 224 -> 1024-d, 512 -> 4096-d (train.py:259,261)."""
    p = O.synthetic_params("densenet121", seed=1)
    with torch.no_grad():
        assert O.densenet121_features(torch.zeros(1, 3, 224, 224), p).shape == (1, 1024)
        assert O.densenet121_features(torch.zeros(1, 3, 512, 512), p).shape == (1, 4096)
        assert O.resnet18_v2_features(torch.zeros(1, 3, 224, 224), O.synthetic_params("resnet18_v2", 1)).shape == (1, 512)


@pytest.mark.parametrize("cell", ["gru", "lstm"])
def test_birnn_oracle_matches_torch_nn(cell):
    D, H, B, T = 24, 16, 3, 6
    p = O.synthetic_rnn_params(cell, D, H, seed=9)
    mod = (torch.nn.GRU if cell == "gru" else torch.nn.LSTM)(D, H, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for d, suf in (("l0", ""), ("r0", "_reverse")):
            getattr(mod, "weight_ih_l0" + suf).copy_(p[d + "_i2h_weight"])
            getattr(mod, "weight_hh_l0" + suf).copy_(p[d + "_h2h_weight"])
            getattr(mod, "bias_ih_l0" + suf).copy_(p[d + "_i2h_bias"])
            getattr(mod, "bias_hh_l0" + suf).copy_(p[d + "_h2h_bias"])
        x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(1))
        ref, _ = mod(x)
        mine = O.birnn_layer(x, p, cell, H)
    assert torch.allclose(mine, ref, atol=1e-5), (mine - ref).abs().max()


def test_time_distributed_shape_kat():
    """The reference's only executable check (definitions.py:156-168): (3,2,3,2,2) -> Conv2D(4,k=2) -> (3,2,4,1,1)."""
    w = torch.randn(4, 3, 2, 2)
    y = O.time_distributed(lambda t: torch.relu(torch.nn.functional.conv2d(t, w)), torch.ones(3, 2, 3, 2, 2))
    assert y.shape == (3, 2, 4, 1, 1)


def test_param_counts_match_survey_inventory():
    n = sum(int(np.prod(s)) for k, s in O.densenet121_param_shapes() if not k.endswith(("running_mean", "running_var")))
    assert n == 6953856  # SURVEY.md Appendix B (torchvision restatement)
    assert O.flatten_params("densenet121", O.synthetic_params("densenet121")).numel() == 7037504


def test_oracle_reproduces_golden_fixtures():
    gold = np.load(GOLD)
    with torch.no_grad():
        _, x = O.synthetic_frames(2, 224, seed=100)
        for arch in ("densenet121", "resnet18_v2"):
            f = O.FEATURES[arch](x, O.synthetic_params(arch, seed=1234)).numpy()
            ref = gold[arch + "_feats"]
            assert np.abs(f - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()), arch
        g = torch.Generator().manual_seed(3)
        feats = torch.randn(3, 5, 64, generator=g).relu()
        for cell in ("gru", "lstm"):
            y = O.birnn_layer(feats, O.synthetic_rnn_params(cell, 64, 128, seed=4321), cell, 128).numpy()
            assert np.abs(y - gold["birnn_%s_y" % cell]).max() < 1e-5
        gg = torch.Generator().manual_seed(77)
        cw = (torch.rand(11, 256, generator=gg) * 2 - 1) * 0.07
        logits = O.cnnrnn(feats, None, O.synthetic_rnn_params("gru", 64, 128, seed=4321), "gru", 128, cw, torch.zeros(11),
                          feats=True).numpy()
        assert np.abs(logits - gold["cnnrnn_feats_logits"]).max() < 1e-5
        assert np.abs(O.temporal_pooling(feats, None, "mean", feats=True).numpy() - gold["temporal_pool_mean"]).max() < 1e-6


def test_training_mode_oracle_matches_torchvision_train_mode():
    """The CNN-backward parity tests differentiate the oracle in TRAINING mode (batch statistics).  Pin that mode against
    torchvision's DenseNet-121 in .train(): same features, same gradient at the first and at a late layer, and the running
    statistics follow the MXNet convention the reference runs on (0.9 / 0.1 blend with the BIASED batch variance)."""
    p = O.synthetic_params("densenet121", seed=1234)
    _, x = O.synthetic_frames(2, 224, seed=3)
    tv = _torchvision_densenet_with(p).train()
    feats_tv = torch.flatten(torch.nn.functional.avg_pool2d(torch.relu(tv.features(x)), 7), 1)
    feats_tv.square().sum().backward()
    q = {k: v.clone().requires_grad_(not k.endswith(("running_mean", "running_var"))) for k, v in p.items()}
    q["_update_running"] = True
    feats = O.FEATURES["densenet121"](x, q, training=True)
    feats.square().sum().backward()
    assert (feats - feats_tv).abs().max().item() < 1e-4 * feats_tv.abs().max().item()
    for ours, theirs in (("conv0.weight", tv.features.conv0.weight), ("block4.layer16.conv2.weight",
                                                                       tv.features.denseblock4.denselayer16.conv2.weight),
                         ("bn5.gamma", tv.features.norm5.weight)):
        a, b = q[ours].grad, theirs.grad
        cos = float((a * b).sum() / (a.norm() * b.norm()))
        assert cos > 0.9999, (ours, cos)
    # running statistics of the stem BatchNorm: blend of the old value and the batch statistics of conv0's output
    with torch.no_grad():
        y = torch.nn.functional.conv2d(x, p["conv0.weight"], stride=2, padding=3)
        mu, var_b = y.mean(dim=(0, 2, 3)), y.var(dim=(0, 2, 3), unbiased=False)
    assert (q["bn0.running_mean"] - (0.9 * p["bn0.running_mean"] + 0.1 * mu)).abs().max().item() < 1e-5
    assert (q["bn0.running_var"] - (0.9 * p["bn0.running_var"] + 0.1 * var_b)).abs().max().item() < 1e-5
    # and inference mode is untouched by the switch
    with torch.no_grad():
        assert torch.equal(O.FEATURES["densenet121"](x, p), O.FEATURES["densenet121"](x, p, training=False))


@pytest.mark.parametrize("cell", ["lstm", "gru"])
def test_captioning_unroll_with_valid_length_matches_packed_torch_rnn(cell):
    """SURVEY.md A.4 (MXNet `unroll(valid_length=...)` + BidirectionalCell with SequenceReverse) restated in
    oracle/captioning.py, against an independent implementation of the same semantics: torch.nn.LSTM/GRU on a packed
    sequence (outputs zero past each length, final states taken at the last valid step, reverse direction starting at
    position len-1)."""
    from oracle import captioning as C
    B, T, D, H = 4, 9, 12, 16
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, D, generator=g)
    vl = torch.tensor([9., 3., 6., 1.])
    G = 4 if cell == "lstm" else 3
    p = {}
    for side in ("l_cell", "r_cell"):
        for n, shape in (("i2h_weight", (G * H, D)), ("h2h_weight", (G * H, H)), ("i2h_bias", (G * H,)), ("h2h_bias", (G * H,))):
            p["enc.%s.%s" % (side, n)] = torch.randn(shape, generator=g) * 0.3
    l_out, l_st = C._unroll(cell, p, "enc.l_cell.", x, vl, H)
    r_out, r_st = C._unroll(cell, p, "enc.r_cell.", C._sequence_reverse(x, vl), vl, H)
    out = torch.cat([l_out, C._sequence_reverse(r_out, vl)], dim=2)
    mod = (torch.nn.LSTM if cell == "lstm" else torch.nn.GRU)(D, H, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for side, suf in (("l_cell", ""), ("r_cell", "_reverse")):
            getattr(mod, "weight_ih_l0" + suf).copy_(p["enc.%s.i2h_weight" % side])
            getattr(mod, "weight_hh_l0" + suf).copy_(p["enc.%s.h2h_weight" % side])
            getattr(mod, "bias_ih_l0" + suf).copy_(p["enc.%s.i2h_bias" % side])
            getattr(mod, "bias_hh_l0" + suf).copy_(p["enc.%s.h2h_bias" % side])
        packed = torch.nn.utils.rnn.pack_padded_sequence(x, vl.long(), batch_first=True, enforce_sorted=False)
        y, hn = mod(packed)
        y, _ = torch.nn.utils.rnn.pad_packed_sequence(y, batch_first=True, total_length=T)
    assert (out - y).abs().max().item() < 1e-5
    h_n = hn[0] if cell == "lstm" else hn
    assert (l_st[0] - h_n[0]).abs().max().item() < 1e-5 and (r_st[0] - h_n[1]).abs().max().item() < 1e-5
    if cell == "lstm":
        assert (l_st[1] - hn[1][0]).abs().max().item() < 1e-5 and (r_st[1] - hn[1][1]).abs().max().item() < 1e-5


def test_captioning_attention_and_masked_ce_match_torch_primitives():
    """SURVEY.md A.5 / A.8 restated in oracle/captioning.py against torch's own primitives: scaled-Luong attention ==
    scaled_dot_product_attention of the projected query over the raw memory with a key mask; MaskedSoftmaxCELoss ==
    per-token cross entropy, masked, averaged over the PADDED length."""
    import torch.nn.functional as F
    from oracle import captioning as C
    g = torch.Generator().manual_seed(1)
    B, T, H = 3, 7, 16
    p = {"decoder.attention_cell.proj_query.weight": torch.randn(H, H, generator=g) * 0.3}
    query, mem = torch.randn(B, H, generator=g), torch.randn(B, T, H, generator=g)
    vl = torch.tensor([7, 2, 5])
    masks = (torch.arange(T).reshape(1, T) < vl.reshape(B, 1)).float()
    ctx, w = C.attention(p, query, mem, masks, H)
    qp = (query @ p["decoder.attention_cell.proj_query.weight"].t()).unsqueeze(1)
    ref = F.scaled_dot_product_attention(qp, mem, mem, attn_mask=masks.bool().unsqueeze(1), scale=1.0 / H ** 0.5).squeeze(1)
    assert (ctx - ref).abs().max().item() < 1e-5
    assert (w.sum(dim=1) - 1).abs().max().item() < 1e-5 and (w * (1 - masks)).abs().max().item() == 0
    V, Tt = 11, 6
    pred = torch.randn(B, Tt, V, generator=g)
    label = torch.randint(0, V, (B, Tt), generator=g).float()
    tvl = torch.tensor([6., 1., 4.])
    ce = F.cross_entropy(pred.reshape(-1, V), label.long().reshape(-1), reduction="none").reshape(B, Tt)
    keep = (torch.arange(Tt).reshape(1, Tt) < tvl.reshape(B, 1)).float()
    assert (C.masked_softmax_ce(pred, label, tvl) - (ce * keep).sum(dim=1) / Tt).abs().max().item() < 1e-6


def test_beam_search_oracle_invariants_and_scorer_values():
    """gluonnlp BeamSearchScorer / BeamSearchSampler as restated in oracle/captioning.py: known values of the
    length penalty, and structural invariants of the sampler output that hold for any model: beam 1 == greedy decoding, scores
    sorted descending, every hypothesis starts with BOS, ends with EOS at valid_length-1 and is padded with -1 afterwards."""
    from oracle import captioning as C
    # scorer: (log_probs + scores * prev_lp) / lp with lp = (K+step)^a / (K+1)^a and prev_lp = 1 at step 1
    lp = torch.tensor([[[-1.0, -2.0]]])
    sc = torch.tensor([[-3.0]])
    assert torch.allclose(C.beam_search_scorer(lp, sc, 1, alpha=1.0, K=5), (lp + sc.unsqueeze(-1)) / 1.0)
    assert torch.allclose(C.beam_search_scorer(lp, sc, 2, alpha=1.0, K=5), (lp + (sc * 1.0).unsqueeze(-1)) / (7.0 / 6.0))
    assert torch.allclose(C.beam_search_scorer(lp, sc, 3, alpha=1.0, K=5), (lp + (sc * (7.0 / 6.0)).unsqueeze(-1)) / (8.0 / 6.0))
    cell, H, D, E, V = "gru", 16, 12, 8, 13
    p = C.synthetic_gnmt_params(seed=3, scale=0.5, cell=cell, H=H, D_src=D, E=E, V=V)
    x, vl = C.synthetic_sources(3, 6, D, seed=4)
    with torch.no_grad():
        s, scores, vlen = C.translate(p, x, vl, cell=cell, H=H, beam=4, max_length=9, bos=2, eos=3)
        s1, _, v1 = C.translate(p, x, vl, cell=cell, H=H, beam=1, max_length=9, bos=2, eos=3)
        # greedy reference: repeatedly take the arg-max token of decode_step
        mem, st = C.encoder_forward(p, x, vl, cell, H)
        states = C.init_state_from_encoder(mem, st, vl)
        tok = torch.full((3,), 2.0)
        greedy = [tok.clone()]
        alive = torch.ones(3, dtype=torch.bool)
        for _ in range(9):
            logits, states = C.decode_step_logits(p, tok, states, cell=cell, H=H)
            nxt = logits.argmax(dim=-1).float()
            greedy.append(torch.where(alive, nxt, torch.full_like(nxt, -1.0)))
            alive = alive & (nxt != 3)
            tok = nxt.clamp(min=0)
            if not alive.any():
                break
    assert (scores[:, :-1] >= scores[:, 1:]).all()
    B, beam, L = s.shape
    for b in range(B):
        for k in range(beam):
            n = int(vlen[b, k])
            assert s[b, k, 0] == 2 and s[b, k, n - 1] == 3 and (s[b, k, n:] == -1).all() and (s[b, k, 1:n - 1] >= 0).all()
    g = torch.stack(greedy, dim=1)
    for b in range(3):
        n = int(v1[b, 0])
        m = min(n, g.shape[1])
        assert torch.equal(s1[b, 0, :m].float(), g[b, :m]), (s1[b, 0], g[b])


def test_captioning_oracle_reproduces_golden_fixtures():
    """tests/golden/captioning_oracle.npz (tools/make_golden_captioning.py): teacher-forced logits, masked-CE loss and the beam
    search (token ids, valid lengths bit-exact; scores 1e-5) of seeded LSTM / GRU captioners."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_captioning", os.path.join(os.path.dirname(__file__), "..", "tools",
                                                                                          "make_golden_captioning.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "captioning_oracle.npz"))
    for cell in ("lstm", "gru"):
        got = mod.compute(cell)
        assert np.abs(got["logits"] - gold[cell + "_logits"]).max() < 1e-5
        assert np.abs(got["loss"] - gold[cell + "_loss"]).max() < 1e-5
        assert np.array_equal(got["samples"], gold[cell + "_samples"])
        assert np.array_equal(got["valid_length"], gold[cell + "_valid_length"])
        assert np.abs(got["scores"] - gold[cell + "_scores"]).max() < 1e-5
