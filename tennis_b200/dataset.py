"""TennisSet: sample assembly of the reference's dataset.py for the hot path.

Real mode (data/splits, data/annotations/... present): split / label / point / caption files, events, OTH balancing, the
frame / feature path scheme and windowing as the reference reads them (dataset.py:135-150, 186-233, 268-287, 302-437); frame
extraction from the videos is not part of the path.  Synthetic mode (no dataset on disk, `synthetic=` given or TENNIS_SYNTHETIC=1): a seeded
procedural stand-in with the same sample tuples, so the scripts, the batching code and the kernels see the shapes
and index semantics of the real thing:
    events : (img (T,3,S,S) | (3,S,S) | feats (T,D), label int, idx int)
    captions: (frames (n,3,S,S) | feats (n,D), cap ids int32 [bos ... eos], n, len(cap)[, idx])
"""
import math
import os

import numpy as np
import torch

from .vocab import Vocab, count_tokens

CLASSES = ['OTH', 'SFI', 'SFF', 'SFL', 'SNI', 'SNF', 'SNL', 'HFL', 'HFR', 'HNL', 'HNR']  # data/classes.names


def window_frames(center, window, stride, every, video_length):
    """Frame indices of a clip (reference dataset.py:190-201): offsets int(-W/2) .. ceil(W/2)-1, each clamped to
    [0, last frame that is a multiple of `every`]; clamping duplicates edge frames (Appendix C #16)."""
    offsets = list(range(int(-window / 2), int(math.ceil(window / 2))))
    max_frame = video_length - every
    for i in range(every):
        if (max_frame - i) % every == 0:
            max_frame -= i
            break
    return [min(max(0, center + off * stride), int(max_frame)) for off in offsets]


def image_path(root_dir, video_name, frame_number, chunk_size=1000):
    chunk = int(frame_number / chunk_size) * chunk_size
    return os.path.join(root_dir, video_name + '.mp4', '{:010d}'.format(chunk), '{:010d}.jpg'.format(frame_number))


def feature_path(feat_dir, video_name, frame_number, chunk_size=1000):
    chunk = int(frame_number / chunk_size) * chunk_size
    return os.path.join(feat_dir, video_name + '.mp4', '{:010d}'.format(chunk), '{:010d}.npy'.format(frame_number))


_WORDS = ("the player serves ball into net far near left right forehand backhand hits return and it goes out wide long "
          "down line cross court fault in play point won by winner rally short deep body t first second").split()


class TennisSet(object):
    def __init__(self, root='data', captions=False, transform=None, split='train', every=1, balance=True, padding=1,
                 stride=1, window=1, model_id='0000', split_id='02', flow=False, max_cap_len=-1, vocab=None,
                 inference=False, feats_model=None, save_feats=False, synthetic=None, data_shape=224):
        if flow:
            raise NotImplementedError("optical-flow inputs (TwoStreamModel) are outside the hot path")
        self._root, self._captions, self._split = root, captions, split
        self._balance, self._every, self._padding, self._stride, self._window = balance, every, padding, stride, window
        self._transform, self._inference, self._save_feats = transform, inference, save_feats
        self._frames_dir = os.path.join(root, "frames")
        self._splits_dir = os.path.join(root, "splits")
        self.output_dir = os.path.join(root, "outputs", model_id, split)
        self._load_feats = feats_model is not None
        self.feat_dir = os.path.join(root, "features", feats_model if feats_model is not None else model_id)
        self.classes = list(CLASSES)
        self._shape = data_shape
        splits_file = os.path.join(self._splits_dir, split_id, split + '.txt')
        if synthetic is None and (os.environ.get("TENNIS_SYNTHETIC") == "1" or not os.path.exists(splits_file)):
            synthetic = {}
        self._synthetic = synthetic
        if synthetic is not None:
            self._init_synthetic(**synthetic)
        else:
            self._init_real(splits_file)
        if self._captions:
            self._samples = list(self._points.keys())
            words = ' '.join(p[4] for p in self._points.values()).split()
            self.vocab = vocab if vocab is not None else Vocab(count_tokens(words))
            for pid in self._samples:
                toks = self._points[pid][4].split()
                ids = self.vocab[toks[:max_cap_len]] if max_cap_len >= 0 else self.vocab[toks]
                ids = [self.vocab[self.vocab.bos_token]] + ids + [self.vocab[self.vocab.eos_token]]
                self._points[pid] = self._points[pid][:5] + [np.array(ids, dtype=np.int32)]

    # ------------------------------------------------------------------ sources
    def _init_synthetic(self, num_videos=2, frames_per_video=96, num_points=6, feat_dim=1024, seed=0):
        self._seed = {'train': 1, 'val': 2, 'test': 3}.get(self._split, 4) * 1000 + seed
        self._feat_dim = feat_dim
        rng = np.random.RandomState(self._seed)
        self._videos = ['V%03d' % i for i in range(num_videos)]
        self._video_lengths = {v: frames_per_video for v in self._videos}
        self._samples = []
        for v in self._videos:
            for f in range(0, frames_per_video, self._every):
                self._samples.append([v, f, self.classes[int(rng.randint(0, len(self.classes)))]])
        self._points = {}
        for i in range(num_points):
            v = self._videos[i % num_videos]
            n = int(rng.randint(8, max(9, frames_per_video // 2)))
            start = int(rng.randint(0, frames_per_video - n))
            cap = ' '.join(_WORDS[int(j)] for j in rng.randint(0, len(_WORDS), size=int(rng.randint(4, 12))))
            self._points['P%04d' % i] = [v, start, start + n, 'point', cap]

    def _init_real(self, splits_file):
        """The on-disk dataset as the reference reads it (dataset.py:302-437):
            splits/<split_id>/<split>.txt          "video frame" per line
            annotations/labels/<video>.txt         "frame CLASS" per line
            annotations/points.txt                 "point_id video start end tag" per line
            annotations/captions.txt               "point_id<TAB>caption" per line
        Samples are [video, frame, class]; events are maximal runs of one class; points are kept when their video and start
        frame belong to the split.  `save_feats` pads every video with 255 'OTH' frames on both sides (:333-345) so that
        every window of the captioner finds its features.  Frames whose image is missing are dropped when images are what is
        loaded (the reference extracts them from the videos first; video decoding is outside the hot path)."""
        root = self._root
        labels_dir = os.path.join(root, "annotations", "labels")
        ann_dir = os.path.join(root, "annotations")
        with open(splits_file) as f:
            samples = [[t[0], int(t[1])] for t in (line.split() for line in f) if len(t) >= 2]
        videos = sorted({s[0] for s in samples})
        labels = {v: {} for v in videos}
        if self._save_feats:
            for v in videos:
                own = [s[1] for s in samples if s[0] == v]
                lo, hi = min(own), max(own)
                for i in range(1, 256):
                    for fr in (lo - i, hi + i):
                        samples.append([v, fr])
                        labels[v][fr] = 'OTH'
        if not self._load_feats:
            present = [s for s in samples if os.path.exists(image_path(self._frames_dir, s[0], s[1]))]
            if len(present) != len(samples):
                import logging
                logging.info("%d of %d frames of split %s have no image under %s and are ignored", len(samples) - len(present),
                             len(samples), self._split, self._frames_dir)
            samples = present
        for v in videos:
            path = os.path.join(labels_dir, v + '.txt')
            if os.path.exists(path):
                with open(path) as f:
                    for t in (line.split() for line in f):
                        if len(t) >= 2:
                            labels[v][int(t[0])] = t[1]
        in_set = {v: [] for v in videos}
        for s in samples:
            s.append(labels[s[0]].get(s[1], 'OTH'))
            in_set[s[0]].append(s[1])
        # events: consecutive in-split frames with the same label (dataset.py:395-410, including its initial 'OTH' event)
        events = []
        for v in videos:
            cur, start, last = 'OTH', -1, -1
            for fr in sorted(in_set[v]):
                if start < 0:
                    start = last = fr
                lab = labels[v].get(fr, 'OTH')
                if lab != cur:
                    events.append([v, start, last, cur])
                    cur, start = lab, fr
                last = fr
            events.append([v, start, last, cur])
        points = {}
        ppath, cpath = os.path.join(ann_dir, 'points.txt'), os.path.join(ann_dir, 'captions.txt')
        if os.path.exists(ppath) and os.path.exists(cpath):
            caps = {}
            with open(cpath) as f:
                for line in f:
                    t = line.rstrip('\n').split('\t')
                    if len(t) >= 2:
                        caps[t[0]] = t[1]
            member = {v: set(fr) for v, fr in in_set.items()}
            with open(ppath) as f:
                for t in (line.split() for line in f):
                    if len(t) >= 4 and t[1] in member and int(t[2]) in member[t[1]] and t[0] in caps:
                        rec = t[1:5] + ['point'] * max(0, 5 - len(t))
                        points[t[0]] = rec + [caps[t[0]]]
        elif self._captions:
            raise FileNotFoundError("captions=True needs %s and %s" % (ppath, cpath))
        self._samples, self._videos, self._events, self._points = samples, videos, events, points
        self._video_lengths = self._real_video_lengths()
        if not self._captions and self._balance:
            self._samples = self._balance_classes()

    def _real_video_lengths(self):
        """Largest frame number on disk per video (dataset.py:439-456: from the frames tree; from the feature tree or its packed
        store when only features exist)."""
        out = {}
        for v in self._videos:
            for tree, ext in ((self._frames_dir, '.jpg'), (self.feat_dir, '.npy')):
                vdir = os.path.join(tree, v + '.mp4')
                if os.path.isdir(vdir):
                    chunks = sorted(d for d in os.listdir(vdir) if d.isdigit())
                    if chunks:
                        files = sorted(fn for fn in os.listdir(os.path.join(vdir, chunks[-1])) if fn.endswith(ext))
                        if files:
                            out[v] = int(files[-1][:-len(ext)])
                            break
            if v not in out:
                packed = self._packed_video(v) if self._load_feats else None
                if packed is not None and len(packed.frames):
                    out[v] = int(packed.frames[-1])
                else:
                    out[v] = max(s[1] for s in self._samples if s[0] == v)
        return out

    def _balance_classes(self, seed=None):
        """Thin out the dominant 'OTH' class to about the size of the next largest one by uniform random sampling
        (dataset.py:268-287)."""
        import random
        rng = random.Random(seed) if seed is not None else random
        counts = self.class_counts()
        ratio = max(counts[1:]) / float(counts[0] + 1)
        return [s for s in self._samples if s[2] != 'OTH' or rng.uniform(0, 1) <= ratio]

    def stats(self):
        """Per-class frame / event counts, or point / frame counts for caption sets (dataset.py:97-130)."""
        out = 'Split: %s\n' % self._split
        if self._captions:
            frames = sum(int(self._points[s][2]) - int(self._points[s][1]) for s in self._samples)
            out += '{0: <8} {1: <8} {2: <5}\n'.format('# Points', '# Frames', 'FperP')
            out += '{0: <8} {1: <8} {2: <5}\n'.format(len(self._samples), frames, int(frames / max(1, len(self._samples))))
            return out
        fcounts = self.class_counts()
        ecounts = [0] * len(self.classes)
        for e in getattr(self, "_events", []):
            ecounts[self.classes.index(e[3])] += 1
        out += '{0: <6} {1: <8} {2: <8} {3: <5}\n'.format('Class', '# Frames', '# Events', 'FperE')
        for i, c in enumerate(self.classes):
            out += '{0: <6} {1: <8} {2: <8} {3: <5}\n'.format(c, fcounts[i], ecounts[i], int(fcounts[i] / (ecounts[i] + .00001)))
        return out

    def __str__(self):
        return '\n\n' + self.__class__.__name__ + '\n' + self.stats() + '\n'

    # ------------------------------------------------------------------ access
    def __len__(self):
        return len(self._samples)

    def _load_frame(self, video, frame):
        """-> fp32 (3,S,S) normalised tensor (what the reference's test transform yields) or a (D,) feature."""
        if self._synthetic is not None:
            g = torch.Generator().manual_seed(hash((self._seed, video, int(frame))) % (2 ** 31))
            if self._load_feats:
                return torch.randn(self._feat_dim, generator=g).relu()
            return torch.randn(3, self._shape, self._shape, generator=g)
        if self._load_feats:
            packed = self._packed_video(video)
            if packed is not None:
                return torch.from_numpy(packed.read([frame])[0])
            return torch.from_numpy(np.load(feature_path(self.feat_dir, video, frame)).astype(np.float32))
        import cv2
        img = cv2.cvtColor(cv2.imread(image_path(self._frames_dir, video, frame), 1), cv2.COLOR_BGR2RGB)
        t = torch.from_numpy(img)
        return self._transform(t) if self._transform is not None else t

    def _packed_video(self, video):
        """The packed feature store of `video` (feature_store.py) when one exists beside the per-frame files, else None."""
        cache = self.__dict__.setdefault("_packed_cache", {})
        if video not in cache:
            from .feature_store import PackedVideo
            cache[video] = PackedVideo(self.feat_dir, video) if PackedVideo.exists(self.feat_dir, video) else None
        return cache[video]

    def _load_frames(self, video, frames):
        """Stack of frames / features of one video; features of a packed video come from ONE gather."""
        if self._synthetic is None and self._load_feats:
            packed = self._packed_video(video)
            if packed is not None:
                return torch.from_numpy(packed.read(frames))
        return torch.stack([self._load_frame(video, f) for f in frames])

    def __getitem__(self, idx):
        sample = self._samples[idx]
        if self._captions:
            vid, start, end = self._points[sample][0], int(self._points[sample][1]), int(self._points[sample][2])
            cap = self._points[sample][5]
            imgs = self._load_frames(vid, [f for c, f in enumerate(range(start, end)) if c % self._every == 0])
            if self._inference:
                return imgs, cap, len(imgs), len(cap), idx
            return imgs, cap, len(imgs), len(cap)
        label = self.classes.index(sample[2])
        if self._window > 1:
            frames = window_frames(sample[1], self._window, self._stride, self._every, self._video_lengths[sample[0]])
            img = self._load_frames(sample[0], frames)
        else:
            img = self._load_frame(sample[0], sample[1])
        return img, label, idx

    def get_data_lens(self):
        assert self._captions
        return [(int((int(self._points[s][2]) - int(self._points[s][1]) + 1) / self._every), len(self._points[s][5]))
                for s in self._samples]

    def get_captions(self, ids=False, split=False):
        out = []
        for s in self._samples:
            cap = self._points[s][4]
            out.append(self._points[s][5] if ids else (cap.split() if split else cap))
        return out

    def save_feature_path(self, idx, chunk_size=1000):
        s = self._samples[idx]
        return feature_path(self.feat_dir, s[0], s[1], chunk_size)

    def class_counts(self):
        counts = [0] * len(self.classes)
        for s in self._samples:
            counts[self.classes.index(s[2])] += 1
        return counts

    @property
    def num_class(self):
        return len(self.classes)
