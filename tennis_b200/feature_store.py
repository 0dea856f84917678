"""Packed per-video feature store: the hand-off between the detector (`save_features`, reference
train.py:530-545) and everything that trains on features (dataset.py:141-150,202-204).

The reference writes ONE .npy PER FRAME, `<feat_dir>/<video>.mp4/<frame//1000*1000:010d>/<frame:010d>.npy` (fp32 (D,)) —
786 k tiny files for the full dataset, and a 32-frame window costs 32 open/read/close round trips.  Packed layout, next to it:

    <feat_dir>/<video>.mp4.packed/features.npy   float32 (n, D), rows in ascending frame order, readable with np.load(mmap_mode)
    <feat_dir>/<video>.mp4.packed/frames.npy     int64   (n,)    the frame number of every row

A window is then one fancy-index gather from a memory map.  `pack_video` converts a reference-layout directory, `unpack_video`
writes the reference layout back (compatibility writer), and TennisSet reads a packed video when it exists and falls back to
the per-frame files otherwise, so both layouts stay valid inputs.
"""
import os

import numpy as np


def packed_dir(feat_dir, video_name):
    return os.path.join(feat_dir, video_name + '.mp4.packed')


def per_frame_path(feat_dir, video_name, frame_number, chunk_size=1000):
    chunk = int(frame_number) // chunk_size * chunk_size
    return os.path.join(feat_dir, video_name + '.mp4', '{:010d}'.format(chunk), '{:010d}.npy'.format(int(frame_number)))


def write_packed(feat_dir, video_name, frame_numbers, features):
    """frame_numbers (n,) ints, features (n, D) float32 -> sorted by frame, duplicates rejected."""
    frames = np.asarray(frame_numbers, dtype=np.int64).reshape(-1)
    feats = np.ascontiguousarray(np.asarray(features, dtype=np.float32))
    if feats.ndim != 2 or feats.shape[0] != frames.shape[0]:
        raise ValueError("features must be (n, D) with one row per frame number")
    order = np.argsort(frames, kind="stable")
    frames, feats = frames[order], feats[order]
    if frames.size and (np.diff(frames) == 0).any():
        raise ValueError("duplicate frame numbers for video %s" % video_name)
    d = packed_dir(feat_dir, video_name)
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, 'frames.npy'), frames)
    np.save(os.path.join(d, 'features.npy'), feats)
    return d


def pack_video(feat_dir, video_name):
    """Gather every `<chunk>/<frame>.npy` of a video in the reference layout into the packed layout; returns the row count."""
    root = os.path.join(feat_dir, video_name + '.mp4')
    frames, rows = [], []
    for chunk in sorted(os.listdir(root)):
        cdir = os.path.join(root, chunk)
        if not os.path.isdir(cdir):
            continue
        for fn in sorted(os.listdir(cdir)):
            if fn.endswith('.npy'):
                frames.append(int(fn[:-4]))
                rows.append(np.load(os.path.join(cdir, fn)).astype(np.float32).reshape(-1))
    if not rows:
        raise FileNotFoundError("no per-frame features under %s" % root)
    write_packed(feat_dir, video_name, frames, np.stack(rows))
    return len(rows)


def unpack_video(feat_dir, video_name, chunk_size=1000):
    """Compatibility writer: packed layout -> the reference's one-file-per-frame layout (bit-identical rows)."""
    store = PackedVideo(feat_dir, video_name)
    for row, frame in enumerate(store.frames):
        path = per_frame_path(feat_dir, video_name, frame, chunk_size)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.save(path, np.asarray(store.features[row]))
    return len(store.frames)


class PackedVideo(object):
    """Memory-mapped view of one packed video."""

    def __init__(self, feat_dir, video_name):
        d = packed_dir(feat_dir, video_name)
        self.frames = np.load(os.path.join(d, 'frames.npy'))
        self.features = np.load(os.path.join(d, 'features.npy'), mmap_mode='r')
        if self.features.shape[0] != self.frames.shape[0]:
            raise ValueError("corrupt packed store %s: %d rows vs %d frame numbers" % (d, self.features.shape[0], self.frames.shape[0]))

    @staticmethod
    def exists(feat_dir, video_name):
        d = packed_dir(feat_dir, video_name)
        return os.path.exists(os.path.join(d, 'frames.npy')) and os.path.exists(os.path.join(d, 'features.npy'))

    @property
    def dim(self):
        return int(self.features.shape[1])

    def rows_of(self, frame_numbers):
        want = np.asarray(frame_numbers, dtype=np.int64).reshape(-1)
        pos = np.searchsorted(self.frames, want)
        pos = np.clip(pos, 0, max(0, len(self.frames) - 1))
        if len(self.frames) == 0 or (self.frames[pos] != want).any():
            missing = want[(self.frames[pos] != want)] if len(self.frames) else want
            raise KeyError("frames %s are not in the packed store" % missing[:5].tolist())
        return pos

    def read(self, frame_numbers):
        """(T,) frame numbers (repeats allowed: windows clamp at the video ends) -> float32 (T, D) in the given order."""
        return np.ascontiguousarray(self.features[self.rows_of(frame_numbers)])
