"""Convert a feature directory written by `evaluate.py --save_feats` / the reference's save_features (one .npy per frame) into
the packed per-video store of tennis_b200/feature_store.py, or back.
usage: python tools/pack_features.py <feat_dir> [--unpack]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tennis_b200 import feature_store as FS  # noqa: E402

feat_dir = sys.argv[1]
unpack = "--unpack" in sys.argv[2:]
suffix = ".mp4.packed" if unpack else ".mp4"
for entry in sorted(os.listdir(feat_dir)):
    if not entry.endswith(suffix) or not os.path.isdir(os.path.join(feat_dir, entry)):
        continue
    video = entry[:-len(suffix)]
    n = FS.unpack_video(feat_dir, video) if unpack else FS.pack_video(feat_dir, video)
    print("%s: %d frames %s" % (video, n, "unpacked" if unpack else "packed"))
