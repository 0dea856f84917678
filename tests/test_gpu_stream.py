"""End-to-end stream (BASELINE.json configs[4], tennis_b200/stream.py): stage-wise parity against the CPU oracle on one short
synthetic video -- features within the stated bf16 tolerance, event classes equal to the oracle head run on the same features
wherever its margin exceeds the feature noise, caption token ids equal to the oracle captioner run on the same features."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_stream_one_video_matches_oracle_stage_by_stage(tmp_path):
    import bench
    import run_stream
    from oracle import captioning as C
    from oracle import vision as O
    from tennis_b200 import feature_store
    from tennis_b200 import synthetic as S
    from tennis_b200.dataset import window_frames
    from tennis_b200.stream import StreamPipeline, synthetic_video
    dev = torch.device("cuda", 0)
    det, cap, tr = run_stream.build(dev)
    F, W, SEG = 150, 32, 64
    pipe = StreamPipeline(det, cap, tr, window=W, segment=SEG, feat_dir=str(tmp_path), head_batch=64)
    frames = synthetic_video(F, 224, seed=901)
    r = pipe.run_video("V000", frames)
    feats = r["features"].cpu()
    # stage 1: features vs the fp32 oracle CNN on the same (normalised) frames
    p = O.synthetic_params("densenet121", seed=1234)
    with torch.no_grad():
        ref_f = O.FEATURES["densenet121"](S.normalize_u8(frames[:32]), p)
    assert (feats[:32] - ref_f).abs().max().item() < 1e-2 * ref_f.abs().max().item()
    # the on-disk hand-off holds exactly these rows
    store = feature_store.PackedVideo(str(tmp_path), "V000")
    assert store.features.shape == (F, 1024) and (torch.from_numpy(store.read(range(F))) == feats).all()
    # stage 2: events -- oracle bi-GRU head on OUR features, same windows (dataset.window_frames)
    O2, _, rp, cw, cb = bench.build_oracle_model()
    idx = torch.tensor([window_frames(i, W, 1, 1, F) for i in range(F)])
    with torch.no_grad():
        y = O.birnn_layer(feats[idx], rp, "gru", 128).max(dim=1).values
        ref_l = y @ cw.t() + cb
    got_l = r["event_logits"].cpu()
    assert got_l.shape == ref_l.shape == (F, 11)
    assert (got_l - ref_l).abs().max().item() < 2e-2  # bf16 input projection of the head
    top2 = ref_l.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 4e-2
    assert torch.equal(r["event_classes"][safe], ref_l.argmax(1)[safe])
    # stage 3: captions -- oracle GNMT on OUR features of every point
    pg = S.synthetic_gnmt_params(seed=10000, scale=0.35, cell="lstm", H=128, D_src=1024, E=100, V=254)
    bounds = [(s, min(F, s + SEG)) for s in range(0, F, SEG)]
    assert len(r["captions"]) == len(bounds) == 3
    src = torch.zeros(len(bounds), SEG, 1024)
    vl = torch.zeros(len(bounds))
    for k, (s, e) in enumerate(bounds):
        src[k, :e - s] = feats[s:e]
        vl[k] = e - s
    with torch.no_grad():
        s_ref, _, v_ref = C.translate(pg, src, vl, cell="lstm", H=128, beam=5, max_length=50, bos=2, eos=3, alpha=1.0, K=5)
    ref_t = C.best_tokens(s_ref, v_ref)
    same = sum(int(a == b) for a, b in zip(r["captions"], ref_t))
    print("stream captions: %d / %d points token-identical; first 12 tokens equal for all" % (same, len(ref_t)))
    for a, b in zip(r["captions"], ref_t):
        assert a[:12] == b[:12]
