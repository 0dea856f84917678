"""End-to-end stream: frames -> DenseNet-121 features -> per-frame events -> point segments -> captions
(BASELINE.json configs[4]; the reference chains three scripts for this: the detector dumps features (train.py:530-545 /
evaluate.py:306-321), the feature files are the captioner's source sequences (dataset.py:141-150,166-168,202-204) and
train_gnmt.py:280-294 / evaluate_gnmt.py translate them).

One process per GPU.  Per video:
  1. features   the video's frames are sharded contiguously over the ranks (any frame count: parallel.balanced_range), each rank
                runs the CNN on its shard and ONE ragged all-gather assembles the (F, D) feature matrix on every rank;
  2. store      rank 0 writes the packed per-video feature store (feature_store.py) -- the on-disk hand-off of the reference --
                and every later stage reads the features it needs back from the device copy;
  3. events     every frame's stride-1 window of `window` frames (dataset.window_frames: clamped at the video ends) goes through
                the temporal head (bi-GRU + max + Dense); windows are sharded over the ranks, logits all-gathered;
  4. captions   the video is cut into `segment`-frame points (214 = the dataset's average point length, README.md:52); each
                segment's features are one source sequence of the GNMT captioner; segments are sharded over the ranks (the
                captioner is replicated), token ids gathered on every rank.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import feature_store
from .dataset import window_frames
from .parallel import all_gather_ragged, balanced_range


class StreamPipeline(object):
    def __init__(self, detector, captioner, translator, window=32, segment=214, feat_dir=None, head_batch=1024, group=None):
        """detector: CNNRNN over a FrameModel backbone; captioner: NMTModel; translator: BeamSearchTranslator(captioner)."""
        self.det, self.cap, self.tr = detector, captioner, translator
        self.window, self.segment, self.feat_dir, self.head_batch = int(window), int(segment), feat_dir, int(head_batch)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # ------------------------------------------------------------------ stages
    def features(self, frames_u8, chunk=512):
        """frames_u8: (F,H,W,3) uint8 host tensor (pinned) -> (F,D) fp32 features on the device, identical on every rank."""
        F = frames_u8.shape[0]
        lo, hi = balanced_range(F, self.rank, self.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        parts = []
        for c0 in range(lo, hi, chunk):
            x = frames_u8[c0:min(hi, c0 + chunk)].to(dev, non_blocking=True)
            parts.append(self.det.td.model(x))
        if parts:
            local = torch.cat(parts, 0)
        else:  # more ranks than frames
            D = self.det.td.model.feature_dim(frames_u8.shape[1], frames_u8.shape[2])
            local = torch.empty((0, D), dtype=torch.float32, device=dev)
        return all_gather_ragged(local, F, self.world, self.group)

    def store(self, video, feats):
        """The detector -> captioner hand-off on disk (packed layout; tools/pack_features.py converts to the per-frame files)."""
        if self.feat_dir is None or self.rank != 0:
            return None
        return feature_store.write_packed(self.feat_dir, video, np.arange(feats.shape[0]), feats.cpu().numpy())

    def window_index(self, F):
        idx = [window_frames(i, self.window, 1, 1, F) for i in range(F)]
        return torch.tensor(idx, dtype=torch.long)

    def events(self, feats):
        """(F,D) features -> (F,C) event logits: frame i is classified from its window of `window` frames."""
        F, D = feats.shape
        lo, hi = balanced_range(F, self.rank, self.world)
        idx = self.window_index(F)[lo:hi].to(feats.device)
        outs = []
        for c0 in range(0, hi - lo, self.head_batch):
            win = feats[idx[c0:c0 + self.head_batch]]  # (n, window, D) gather
            y = self.det.rnn.forward_max(win)
            outs.append(self.det.classes(y) if self.det.classes else y)
        if outs:
            local = torch.cat(outs, 0)
        else:
            local = self.det.classes(self.det.rnn.forward_max(feats[:1].reshape(1, 1, D).expand(1, self.window, D).contiguous()))[:0]
        return all_gather_ragged(local.contiguous(), F, self.world, self.group)

    def captions(self, feats):
        """(F,D) features -> list (per `segment`-frame point) of token-id lists (best beam, BOS/EOS stripped)."""
        F, D = feats.shape
        bounds = [(s, min(F, s + self.segment)) for s in range(0, F, self.segment)]
        lo, hi = balanced_range(len(bounds), self.rank, self.world)
        mine = bounds[lo:hi]
        toks = []
        if mine:
            src = torch.zeros((len(mine), self.segment, D), dtype=torch.float32, device=feats.device)
            vl = torch.zeros(len(mine), dtype=torch.float32, device=feats.device)
            for k, (s, e) in enumerate(mine):
                src[k, :e - s] = feats[s:e]
                vl[k] = e - s
            samples, _, valid = self.tr.translate(src, vl)
            samples, valid = samples.cpu(), valid.cpu()
            for k in range(len(mine)):
                toks.append([int(t) for t in samples[k, 0, 1:int(valid[k, 0]) - 1]])
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, toks, group=self.group)
            toks = [t for part in gathered for t in part]
        return toks

    # ------------------------------------------------------------------ one video, host to host
    def run_video(self, video, frames_u8):
        def sync():
            torch.cuda.synchronize()
            return time.perf_counter()
        t0 = sync()
        feats = self.features(frames_u8)
        t1 = sync()
        self.store(video, feats)
        t2 = sync()
        logits = self.events(feats)
        classes = logits.argmax(dim=1).cpu()
        t3 = sync()
        toks = self.captions(feats)
        t4 = sync()
        return {"video": video, "frames": int(frames_u8.shape[0]), "event_classes": classes, "event_logits": logits,
                "captions": toks, "features": feats, "tokens": sum(len(t) for t in toks),
                "t_features": t1 - t0, "t_store": t2 - t1, "t_events": t3 - t2, "t_captions": t4 - t3, "t_total": t4 - t0}


def synthetic_video(frames, size=224, seed=0):
    """(frames, size, size, 3) uint8 pinned host tensor: a smooth colour field drifting over time + pixel noise (content that
    changes along the video, so events and captions differ from segment to segment)."""
    g = torch.Generator().manual_seed(seed)
    keys = torch.rand(max(2, frames // 64 + 2), 3, 7, 7, generator=g) * 1.6 - 0.3
    pos = torch.linspace(0, keys.shape[0] - 1.001, frames)
    i0 = pos.floor().long()
    w = (pos - i0.float()).reshape(-1, 1, 1, 1)
    low = keys[i0] * (1 - w) + keys[i0 + 1] * w
    out = torch.empty((frames, size, size, 3), dtype=torch.uint8).pin_memory()
    for c0 in range(0, frames, 256):
        f = torch.nn.functional.interpolate(low[c0:c0 + 256], size=(size, size), mode="bilinear", align_corners=False)
        f = f + (torch.rand(f.shape, generator=g) - 0.5) * 0.2
        out[c0:c0 + 256] = (f.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1)
    return out
