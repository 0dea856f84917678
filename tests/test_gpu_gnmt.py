"""GNMT encoder / decoder / beam search (through the C ABI) against the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(cell, H, D, E, V, scale, seed=10000):
    from oracle import captioning as C
    from tennis_b200.gluon import Dropout, Embedding, HybridSequential
    from tennis_b200.models.captioning.gnmt import NMTModel, get_gnmt_encoder_decoder
    from tennis_b200.vocab import Vocab, count_tokens
    p = C.synthetic_gnmt_params(seed=seed, scale=scale, cell=cell, H=H, D_src=D, E=E, V=V)
    vocab = Vocab(count_tokens(["w%03d" % i for i in range(V - 4)]))
    assert len(vocab) == V
    src_embed = HybridSequential()
    src_embed.add(Dropout(0.0))
    enc, dec = get_gnmt_encoder_decoder(cell_type=cell, hidden_size=H, dropout=0.0, num_layers=2, num_bi_layers=1)
    model = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=enc, decoder=dec, embed_size=E, prefix='gnmt_',
                     src_embed=src_embed, tgt_embed=Embedding(V, E))
    params = model.collect_params()
    assert sorted(params.keys()) == sorted(p.keys()), set(params.keys()) ^ set(p.keys())
    dev = torch.device("cuda", 0)
    for k, v in p.items():
        params[k].shape = tuple(v.shape)
        params[k]._data = v.to(dev)
        params[k]._version += 1
    return model, p, vocab


@pytest.mark.parametrize("cell", ["lstm", "gru"])
def test_encoder_matches_oracle(cell):
    from oracle import captioning as C
    H, D = 128, 64
    model, p, _ = _build(cell, H, D, 20, 30, 0.1)
    x, vl = C.synthetic_sources(4, 13, D, seed=100)
    with torch.no_grad():
        mem_ref, st_ref = C.encoder_forward(p, x, vl, cell, H)
    (mem, st), _ = model.encode(x.cuda(), valid_length=vl.cuda())
    torch.cuda.synchronize()
    assert (mem.cpu() - mem_ref).abs().max().item() < 2e-4
    for a, b in zip(st, st_ref):
        for ta, tb in zip(a, b):
            assert (ta.cpu() - tb).abs().max().item() < 2e-4
    # without valid_length
    with torch.no_grad():
        mem_ref2, _ = C.encoder_forward(p, x, None, cell, H)
    (mem2, _), _ = model.encode(x.cuda(), valid_length=None)
    assert (mem2.cpu() - mem_ref2).abs().max().item() < 2e-4


@pytest.mark.parametrize("cell", ["lstm", "gru"])
def test_decode_step_and_teacher_forcing_match_oracle(cell):
    from oracle import captioning as C
    H, D, E, V = 128, 64, 100, 254
    model, p, _ = _build(cell, H, D, E, V, 0.1)
    B, T, Tt = 5, 11, 7
    x, vl = C.synthetic_sources(B, T, D, seed=3)
    g = torch.Generator().manual_seed(5)
    tgt = torch.randint(0, V, (B, Tt), generator=g).float()
    tvl = torch.tensor([7., 6., 4., 2., 7.])
    with torch.no_grad():
        ref = C.nmt_forward(p, x, tgt, vl, tvl, cell=cell, H=H)
        mem, st = C.encoder_forward(p, x, vl, cell, H)
        states = C.init_state_from_encoder(mem, st, vl)
        ref_step, ref_states = C.decode_step_logits(p, tgt[:, 0], states, cell=cell, H=H)
    out, _ = model(x.cuda(), tgt.cuda(), vl.cuda(), tvl.cuda())
    enc_out, _ = model.encode(x.cuda(), valid_length=vl.cuda())
    dstates = model.decoder.init_state_from_encoder(enc_out, vl.cuda())
    logits, new_states, _ = model.decode_step(tgt[:, 0].cuda(), dstates)
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (B, Tt, V)
    assert (out.cpu() - ref).abs().max().item() < 1e-3  # north-star tolerance for fp32 logits
    assert (logits.cpu() - ref_step).abs().max().item() < 1e-3
    assert (new_states[1].cpu() - ref_states[1]).abs().max().item() < 1e-3
    assert torch.equal(out.cpu().argmax(-1), ref.argmax(-1))


@pytest.mark.parametrize("cell,H,beam", [("lstm", 128, 5), ("gru", 128, 4), ("gru", 256, 4)])
def test_beam_search_token_ids_bit_exact(cell, H, beam):
    from oracle import captioning as C
    from tennis_b200.models.captioning.gnmt import BeamSearchScorer
    from tennis_b200.utils.translation import BeamSearchTranslator
    D, E, V = 64, 100, 254
    model, p, vocab = _build(cell, H, D, E, V, 0.35)  # peaked distributions: no fp-noise-sized near-ties
    B, T = 6, 17
    x, vl = C.synthetic_sources(B, T, D, seed=11)
    with torch.no_grad():
        s_ref, sc_ref, v_ref = C.translate(p, x, vl, cell=cell, H=H, beam=beam, max_length=20, bos=2, eos=3, alpha=1.0, K=5)
    tr = BeamSearchTranslator(model, beam_size=beam, scorer=BeamSearchScorer(alpha=1.0, K=5), max_length=20)
    s, sc, v = tr.translate(x.cuda(), vl.cuda())
    torch.cuda.synchronize()
    assert s.shape == s_ref.shape, (s.shape, s_ref.shape)
    assert torch.equal(v.cpu(), v_ref)
    assert torch.equal(s.cpu(), s_ref)
    assert (sc.cpu() - sc_ref).abs().max().item() < 1e-3
    assert C.best_tokens(s.cpu(), v.cpu()) == C.best_tokens(s_ref, v_ref)


def test_beam_search_early_exit_and_max_length():
    """All beams finishing early truncates L exactly like the sampler's `if alive.sum()==0: return`; never finishing
    appends EOS at max_length."""
    from oracle import captioning as C
    from tennis_b200.models.captioning.gnmt import BeamSearchScorer
    from tennis_b200.utils.translation import BeamSearchTranslator
    cell, H, D, E, V = "gru", 128, 64, 100, 30
    model, p, _ = _build(cell, H, D, E, V, 0.35, seed=7)
    x, vl = C.synthetic_sources(3, 9, D, seed=2)
    # force EOS to dominate -> immediate termination
    p2 = dict(p)
    p2["tgt_proj.bias"] = p["tgt_proj.bias"].clone()
    p2["tgt_proj.bias"][3] += 50.0
    model.tgt_proj.bias._data = p2["tgt_proj.bias"].cuda()
    model.tgt_proj.bias._version += 1
    with torch.no_grad():
        s_ref, _, v_ref = C.translate(p2, x, vl, cell=cell, H=H, beam=3, max_length=25)
    s, _, v = BeamSearchTranslator(model, 3, BeamSearchScorer(1.0, 5), 25).translate(x.cuda(), vl.cuda())
    assert s.shape == s_ref.shape and torch.equal(s.cpu(), s_ref) and torch.equal(v.cpu(), v_ref)
    # forbid EOS -> runs to max_length, EOS appended
    p3 = dict(p)
    p3["tgt_proj.bias"] = p["tgt_proj.bias"].clone()
    p3["tgt_proj.bias"][3] -= 50.0
    model.tgt_proj.bias._data = p3["tgt_proj.bias"].cuda()
    model.tgt_proj.bias._version += 1
    with torch.no_grad():
        s_ref, _, v_ref = C.translate(p3, x, vl, cell=cell, H=H, beam=3, max_length=10)
    s, _, v = BeamSearchTranslator(model, 3, BeamSearchScorer(1.0, 5), 10).translate(x.cuda(), vl.cuda())
    assert s.shape == s_ref.shape == (3, 3, 12) and torch.equal(s.cpu(), s_ref) and torch.equal(v.cpu(), v_ref)


def test_masked_softmax_ce_matches_oracle():
    from oracle import captioning as C
    from tennis_b200.gluon import MaskedSoftmaxCELoss
    g = torch.Generator().manual_seed(4)
    pred = torch.randn(5, 9, 254, generator=g)
    label = torch.randint(0, 254, (5, 9), generator=g).float()
    vl = torch.tensor([9., 1., 4., 0., 7.])
    ref = C.masked_softmax_ce(pred, label, vl)
    out = MaskedSoftmaxCELoss()(pred.cuda(), label.cuda(), vl.cuda())
    assert (out.cpu() - ref).abs().max().item() < 1e-5
