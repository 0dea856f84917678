"""CPU-only: dataset sample assembly (D1), caption batching (8f-3), vocabulary, metrics (8f-4), host pipeline schedule."""
import math
import os

import numpy as np
import pytest
import torch

from tennis_b200.dataset import TennisSet, feature_path, image_path, window_frames
from tennis_b200.metrics.vision import PRF1, Accuracy, compute_bleu
from tennis_b200.utils.captioning import FixedBucketSampler, get_dataloaders, pad_stack
from tennis_b200.vocab import Vocab, count_tokens


def _reference_window(center, window, stride, every, video_length):
    """Independent restatement of dataset.py:190-201 written as nested conditionals."""
    out = []
    last = video_length - every
    while last % every != 0:
        last -= 1
    for off in range(int(-window / 2), int(math.ceil(window / 2))):
        f = center + off * stride
        if f < 0:
            f = 0
        if f > last:
            f = last
        out.append(f)
    return out


def test_window_offsets_and_clamping():
    for (c, w, s, e, L) in [(3, 32, 1, 1, 96), (90, 32, 1, 1, 96), (50, 30, 1, 1, 96), (10, 8, 2, 2, 21), (0, 15, 3, 1, 40),
                            (39, 15, 3, 1, 40), (7, 1, 1, 1, 9)]:
        got = window_frames(c, w, s, e, L)
        assert got == _reference_window(c, w, s, e, L)
        assert len(got) == w
    assert window_frames(50, 32, 1, 1, 96) == list(range(34, 66))          # -16 .. +15 (dataset.py:192)
    assert window_frames(2, 4, 1, 1, 96) == [0, 1, 2, 3]                   # clamped at 0 duplicates edge frames


def test_path_scheme_matches_reference_layout():
    assert image_path("data/frames", "V010", 12345) == "data/frames/V010.mp4/0000012000/0000012345.jpg"
    assert feature_path("data/features/0006", "V010", 999) == "data/features/0006/V010.mp4/0000000000/0000000999.npy"


def test_event_samples_shapes_and_determinism():
    ds = TennisSet(split='val', window=8, stride=1, every=4, synthetic={}, data_shape=32)
    x, y, i = ds[5]
    assert x.shape == (8, 3, 32, 32) and x.dtype == torch.float32 and 0 <= y < 11 and i == 5
    x2, _, _ = TennisSet(split='val', window=8, stride=1, every=4, synthetic={}, data_shape=32)[5]
    assert torch.equal(x, x2)
    feats = TennisSet(split='train', window=30, feats_model='0006', synthetic={})[0][0]
    assert feats.shape == (30, 1024)
    single = TennisSet(split='test', window=1, synthetic={}, data_shape=16)[0][0]
    assert single.shape == (3, 16, 16)
    assert ds.classes[0] == 'OTH' and len(ds.classes) == 11


def test_caption_samples_and_vocab_convention():
    dc = TennisSet(split='train', captions=True, synthetic={}, feats_model='0006', max_cap_len=5)
    v = dc.vocab
    assert [v.token_to_idx[t] for t in ('<unk>', '<pad>', '<bos>', '<eos>')] == [0, 1, 2, 3]
    frames, cap, n, ncap = dc[0]
    assert cap.dtype == np.int32 and cap[0] == 2 and cap[-1] == 3 and ncap == len(cap) <= 7
    assert frames.shape[0] == n and frames.shape[1] == 1024
    dv = TennisSet(split='val', captions=True, synthetic={}, feats_model='0006', vocab=v, inference=True)
    assert len(dv[0]) == 5 and dv[0][4] == 0
    lens = dc.get_data_lens()
    assert all(isinstance(a, int) and b >= 2 for a, b in lens)
    vv = Vocab(count_tokens("b a a c c c".split()))
    assert vv.idx_to_token[4:] == ['c', 'a', 'b']  # descending frequency
    vv.set_embedding({'a': np.ones(3, np.float32)})
    assert vv.embedding.idx_to_vec.shape == (7, 3) and vv.embedding.idx_to_vec[0].sum() == 0  # specials -> zero vectors


def test_bucketed_batches_cover_dataset_and_pad_with_zero():
    dc = TennisSet(split='train', captions=True, synthetic={'num_points': 11}, feats_model='0006')
    dv = TennisSet(split='val', captions=True, synthetic={'num_points': 7}, feats_model='0006', vocab=dc.vocab, inference=True)
    tr, va, te = get_dataloaders(dc, dv, dv, batch_size=4, test_batch_size=3, num_buckets=5)
    seen = []
    for src, tgt, sl, tl, idx in va:
        assert src.shape[0] == tgt.shape[0] == sl.shape[0] == idx.shape[0] <= 3
        assert sl.dtype == torch.float32 and int(sl.max()) == src.shape[1] and int(tl.max()) == tgt.shape[1]
        for b in range(src.shape[0]):
            assert (src[b, int(sl[b]):] == 0).all() and (tgt[b, int(tl[b]):] == 0).all()
        seen += idx.tolist()
    assert sorted(seen) == list(range(len(dv)))
    assert sum(b[0].shape[0] for b in tr) == len(dc)
    assert pad_stack([torch.ones(2, 3), torch.ones(4, 3)]).shape == (2, 4, 3)
    s = FixedBucketSampler([(5, 3), (50, 9), (7, 2), (48, 8)], batch_size=2, num_buckets=2)
    assert sorted(map(sorted, s)) == [[0, 2], [1, 3]]


def test_prf1_counts_and_swapped_names():
    m = PRF1(label_names=['OTH', 'A', 'B'])
    labels = torch.tensor([0, 1, 2, 2, 1, 0])
    preds = torch.tensor([0, 1, 1, 2, 1, 1])
    m.update([labels], [preds])
    assert m.mat.astype(int).tolist() == [[1, 1, 0], [0, 2, 0], [0, 1, 1]]
    d = dict(m.get())
    # reference quirk (metrics/vision.py:73-74): "prec" = matches / positives(labels), "rec" = matches / predictions
    assert abs(d['A_prec'] - 2 / 2) < 1e-9 and abs(d['A_rec'] - 2 / 4) < 1e-9
    assert abs(d['AVG_NB_f1'] - np.mean([d['A_f1'], d['B_f1']])) < 1e-12
    a = Accuracy(top_k=2)
    a.update([torch.tensor([2, 0])], [torch.tensor([[0.1, 0.5, 0.4], [0.2, 0.7, 0.1]])])
    assert a.get()[1] == 1.0


def test_bleu_known_values():
    ref = [[["the", "cat", "sat", "on", "the", "mat"]]]
    assert abs(compute_bleu(ref, [["the", "cat", "sat", "on", "the", "mat"]])[0] - 1.0) < 1e-12
    b, prec, bp, rl, tl = compute_bleu(ref, [["the", "cat", "sat", "on", "mat"]])
    assert prec[0] == 1.0 and abs(bp - math.exp(1 - 6 / 5)) < 1e-12 and 0 < b < 1


def test_host_pipeline_chunk_schedule():
    from tennis_b200.parallel import HostPipeline, modelled_makespan, plan_chunks
    h = HostPipeline.__new__(HostPipeline)
    h.chunks = 8
    for B in (1, 3, 16, 17, 64, 100, 256):
        for model in (None, (0.35, 0.8, 0.32), (0.09, 0.8, 0.32), (1.0, 0.0, 0.1)):
            cuts = h._schedule(B, model)
            assert cuts[0] == 0 and cuts[-1] == B and all(b > a for a, b in zip(cuts, cuts[1:]))
    h.chunks = 1
    assert h._schedule(64, (0.35, 0.8, 0.32)) == [0, 64]

    def span(cuts, m):
        return modelled_makespan([b - a for a, b in zip(cuts, cuts[1:])], *m)

    # copy about as slow as compute (fp32 frames): the plan beats one big chunk, the doubling schedule and uniform chunks,
    # and does not end on a large chunk (its kernels cannot start before the whole copy has landed)
    m = (0.349, 0.79, 0.3226)
    cuts = plan_chunks(64, *m)
    assert span(cuts, m) <= min(span([0, 64], m), span([0, 4, 12, 28, 64], m), span(list(range(0, 65, 8)), m)) + 1e-9
    assert cuts[-1] - cuts[-2] <= 16
    # copy 4x faster than compute (uint8 frames): few chunks, small first one
    m8 = (0.087, 0.79, 0.3226)
    cuts8 = plan_chunks(64, *m8)
    assert len(cuts8) - 1 <= 4 and cuts8[1] <= 16
    assert span(cuts8, m8) <= span(cuts, m8)


def test_packed_feature_store_roundtrip_and_dataset_read(tmp_path):
    """feature_store.py: reference per-frame layout -> packed -> reference layout is bit-identical; window reads (with the
    clamped, repeated frames of dataset.py:196-201) equal per-frame reads; TennisSet picks the packed video up."""
    import os
    from tennis_b200 import feature_store as FS
    from tennis_b200.dataset import TennisSet, feature_path, window_frames
    root = str(tmp_path / "data")
    feat_dir = os.path.join(root, "features", "0006")
    rng = np.random.RandomState(0)
    frames = list(range(0, 2400, 3))          # spans three 1000-frame chunk directories
    feats = rng.randn(len(frames), 16).astype(np.float32)
    for f, row in zip(frames, feats):
        path = feature_path(feat_dir, "V007", f)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.save(path, row)
    assert FS.per_frame_path(feat_dir, "V007", 1203) == feature_path(feat_dir, "V007", 1203)
    assert FS.pack_video(feat_dir, "V007") == len(frames)
    store = FS.PackedVideo(feat_dir, "V007")
    assert store.dim == 16 and np.array_equal(store.frames, np.array(frames))
    win = [0, 0, 3, 6, 2397, 2397]            # clamped window: repeats at both ends
    assert np.array_equal(store.read(win), np.stack([feats[frames.index(f)] for f in win]))
    try:
        store.read([4])
        raise AssertionError("a frame that was never extracted must not be served")
    except KeyError:
        pass
    # compatibility writer reproduces the per-frame files bit for bit
    other = os.path.join(root, "features", "copy")
    FS.write_packed(other, "V007", frames[::-1], feats[::-1])     # unsorted input is sorted on write
    assert FS.unpack_video(other, "V007") == len(frames)
    for f in (0, 999, 1002, 2397):
        assert np.array_equal(np.load(feature_path(other, "V007", f)), np.load(feature_path(feat_dir, "V007", f)))
    # the dataset serves windows from the packed video
    os.makedirs(os.path.join(root, "splits"), exist_ok=True)
    ds = TennisSet.__new__(TennisSet)
    ds.feat_dir, ds._synthetic, ds._load_feats = feat_dir, None, True
    fr = window_frames(6, 4, 3, 3, 2400)
    got = ds._load_frames("V007", fr)
    assert got.shape == (4, 16) and np.array_equal(got.numpy(), np.stack([feats[frames.index(f)] for f in fr]))
    assert np.array_equal(ds._load_frame("V007", 1203).numpy(), feats[frames.index(1203)])


def test_real_mode_reads_the_reference_file_layout(tmp_path):
    """A miniature on-disk dataset in the reference's layout (splits / labels / points / captions / per-frame features):
    samples, events, balancing, save_feats padding, caption points and windowed feature reads (dataset.py:302-437)."""
    import os
    from tennis_b200.dataset import TennisSet, feature_path
    root = str(tmp_path / "data")
    os.makedirs(os.path.join(root, "splits", "02"))
    os.makedirs(os.path.join(root, "annotations", "labels"))
    frames = list(range(100, 160))
    lab = lambda f: 'SFI' if 110 <= f < 120 else 'HFL' if 130 <= f < 135 else 'OTH'
    with open(os.path.join(root, "splits", "02", "train.txt"), "w") as f:
        f.write("".join("V001 %d\n" % fr for fr in frames))
    with open(os.path.join(root, "annotations", "labels", "V001.txt"), "w") as f:
        f.write("".join("%d %s\n" % (fr, lab(fr)) for fr in range(0, 400)))
    with open(os.path.join(root, "annotations", "points.txt"), "w") as f:
        f.write("P1 V001 105 125 x\nP2 V001 300 320 x\nP3 V009 10 20 x\n")
    with open(os.path.join(root, "annotations", "captions.txt"), "w") as f:
        f.write("P1\tthe player serves into the net\nP2\tout of split\nP3\tother video\n")
    feat_dir = os.path.join(root, "features", "0006")
    rng = np.random.RandomState(1)
    table = {}
    for fr in range(0, 420):
        table[fr] = rng.randn(8).astype(np.float32)
        path = feature_path(feat_dir, "V001", fr)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.save(path, table[fr])
    ds = TennisSet(root=root, split='train', balance=False, window=4, stride=2, feats_model='0006')
    assert len(ds) == 60 and ds.class_counts()[:2] == [45, 10] and ds.class_counts()[ds.classes.index('HFL')] == 5
    assert [e[1:] for e in ds._events] == [[100, 109, 'OTH'], [110, 119, 'SFI'], [120, 129, 'OTH'], [130, 134, 'HFL'],
                                           [135, 159, 'OTH']]
    x, label, idx = ds[12]                                  # frame 112, offsets -2..1 with stride 2
    assert label == ds.classes.index('SFI') and idx == 12
    assert np.array_equal(x.numpy(), np.stack([table[f] for f in (108, 110, 112, 114)]))
    assert "SFI    10       1" in ds.stats()
    bal = TennisSet(root=root, split='train', balance=True, feats_model='0006')
    assert 15 <= len(bal) <= 60 and bal.class_counts()[1] == 10        # only OTH frames are thinned
    padded = TennisSet(root=root, split='train', balance=False, feats_model='0006', save_feats=True)
    assert len(padded) == 60 + 2 * 255 and min(s[1] for s in padded._samples) == 100 - 255
    cap = TennisSet(root=root, split='train', captions=True, feats_model='0006', max_cap_len=4)
    assert list(cap._points) == ['P1'] and len(cap) == 1
    feats, ids, n, ln = cap[0]
    assert n == 20 and feats.shape == (20, 8) and np.array_equal(feats[0].numpy(), table[105])
    assert ln == 6 and ids[0] == cap.vocab[cap.vocab.bos_token] and ids[-1] == cap.vocab[cap.vocab.eos_token]
    assert cap.get_captions(split=True) == [["the", "player", "serves", "into", "the", "net"]]


REF_DATA = "/root/reference/data"


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference tree not mounted (GPU box): nothing to compare against")
def test_host_tables_match_the_reference_fixtures():
    """The only data artefacts the reference ships: the 11 class names and the 250 x 100 caption-word
    embeddings.  The class table compiled into dataset.py must equal data/classes.names, and the embedding reader + Vocab glue
    must reproduce train_gnmt.py:211-218 on the real file (unit-norm rows, specials get zero vectors)."""
    from tennis_b200.dataset import CLASSES
    from tennis_b200.vocab import Vocab, count_tokens, load_embedding_file
    with open(os.path.join(REF_DATA, "classes.names")) as f:
        assert [line.strip() for line in f if line.strip()] == CLASSES
    table = load_embedding_file(os.path.join(REF_DATA, "embeddings-ex.txt"))
    assert len(table) == 250 and all(v.shape == (100,) for v in table.values())
    norms = np.array([np.linalg.norm(v) for v in table.values()])
    assert np.abs(norms - 1.0).max() < 1e-3
    vocab = Vocab(count_tokens(list(table.keys())))
    assert len(vocab) == 254 and vocab.idx_to_token[:4] == ['<unk>', '<pad>', '<bos>', '<eos>']
    vocab.set_embedding(table)
    emb = np.asarray(vocab.embedding.idx_to_vec)
    assert emb.shape == (254, 100) and not emb[:4].any()
    some = vocab.idx_to_token[10]
    assert np.array_equal(emb[10], table[some])


def test_image_transforms_match_the_reference_recipe():
    """transforms.py: Resize(S+32) is a SQUARE bilinear resize (A.1), CenterCrop takes the middle S x S, output stays uint8 HWC
    for the device-side ToTensor+Normalize; the training transform is reproducible and clips/feature dumps reuse the test one."""
    cv2 = pytest.importorskip("cv2")
    from tennis_b200 import transforms as TF
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(90, 160, 3)).astype(np.uint8)           # a 16:9 frame
    S = 32
    out = TF.TestTransform(S)(torch.from_numpy(img))
    assert out.dtype == torch.uint8 and tuple(out.shape) == (S, S, 3)
    ref = cv2.resize(img, (S + 32, S + 32), interpolation=cv2.INTER_LINEAR)[16:16 + S, 16:16 + S]
    assert np.array_equal(out.numpy(), ref)
    assert TF.center_crop(img, 60).shape == (60, 60, 3) and np.array_equal(TF.center_crop(img, 60), img[15:75, 50:110])
    a = TF.TrainTransform(S, seed=3)(img)
    b = TF.TrainTransform(S, seed=3)(img)
    c = TF.TrainTransform(S, seed=4)(img)
    assert a.dtype == torch.uint8 and tuple(a.shape) == (S, S, 3) and torch.equal(a, b) and not torch.equal(a, c)
    train, test = TF.build_transforms(S, window=8)
    assert train is test
    train, test = TF.build_transforms(S, window=1)
    assert isinstance(train, TF.TrainTransform) and isinstance(test, TF.TestTransform)
    with pytest.raises(ValueError):
        TF.TestTransform(S)(torch.zeros(3, 8, 8))
    # host-side ToTensor + Normalize equals the oracle's recipe (train.py:145-146)
    from oracle.vision import normalize_u8
    f32 = TF.TestTransform(S, to_tensor=True)(img)
    assert f32.dtype == torch.float32 and tuple(f32.shape) == (3, S, S)
    assert (f32 - normalize_u8(out.unsqueeze(0))[0]).abs().max().item() < 1e-6
