#!/usr/bin/env python
"""End-to-end stream (BASELINE.json configs[4]): DenseNet-121 frame features -> GRU events -> GNMT captions on a synthetic
5-video stream, on N GPUs of one box.

    python run_stream.py [--gpus N] [--videos 5] [--frames 2048] [--window 32] [--segment 214] [--feat_dir DIR]

The reference runs this chain as three scripts with the per-frame feature files in between (train.py --save_feats ->
dataset.py:141-150 -> train_gnmt.py:280-294 / evaluate_gnmt.py); tennis_b200/stream.py does it per video in one process per GPU:
frames sharded over the ranks through the CNN, ONE all-gather of features, windows sharded through the temporal head, 214-frame
points sharded through the (replicated) captioner.  Prints one JSON line on rank 0: frames/s, caption tokens/s and the host-to-host
latency per video.  Weights are the seeded synthetic ones of the parity tests (no checkpoints offline)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def build(device):
    import torch
    import bench
    from tennis_b200 import synthetic as S
    from tennis_b200.gluon import Dropout, Embedding, HybridSequential
    from tennis_b200.models.captioning.gnmt import BeamSearchScorer, NMTModel, get_gnmt_encoder_decoder
    from tennis_b200.utils.translation import BeamSearchTranslator
    from tennis_b200.vocab import Vocab, count_tokens
    det = bench.build_model(device)
    D, H, E, V = 1024, 128, 100, 254
    p = S.synthetic_gnmt_params(seed=10000, scale=0.35, cell="lstm", H=H, D_src=D, E=E, V=V)
    vocab = Vocab(count_tokens(["w%03d" % i for i in range(V - 4)]))
    src_embed = HybridSequential()
    src_embed.add(Dropout(0.0))
    enc, dec = get_gnmt_encoder_decoder(cell_type="lstm", hidden_size=H, dropout=0.0, num_layers=2, num_bi_layers=1)
    cap = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=enc, decoder=dec, embed_size=E, prefix="gnmt_", src_embed=src_embed,
                   tgt_embed=Embedding(V, E))
    params = cap.collect_params()
    for k, v in p.items():
        params[k].shape, params[k]._data = tuple(v.shape), v.to(device)
        params[k]._version += 1
    tr = BeamSearchTranslator(cap, beam_size=5, scorer=BeamSearchScorer(alpha=1.0, K=5), max_length=50)
    return det, cap, tr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--videos", type=int, default=5)
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--window", type=int, default=32)
    ap.add_argument("--segment", type=int, default=214)
    ap.add_argument("--feat_dir", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    import torch
    import torch.distributed as dist
    from tennis_b200.stream import StreamPipeline, synthetic_video
    if not torch.cuda.is_available():
        raise SystemExit("run_stream.py needs a CUDA device: tennis_b200 has no CPU fallback")
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    det, cap, tr = build(device)
    tmp = None
    feat_dir = args.feat_dir
    if feat_dir is None and rank == 0:
        tmp = tempfile.TemporaryDirectory()
        feat_dir = tmp.name
    pipe = StreamPipeline(det, cap, tr, window=args.window, segment=args.segment, feat_dir=feat_dir)
    videos = [("V%03d" % i, synthetic_video(args.frames, 224, seed=900 + i)) for i in range(args.videos)]
    pipe.run_video("warmup", videos[0][1][:max(args.window, min(args.frames, 256))])  # engines, workspaces, NCCL channels
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = [pipe.run_video(name, fr) for name, fr in videos]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if rank == 0:
        frames = sum(r["frames"] for r in res)
        tokens = sum(r["tokens"] for r in res)
        t_det = sum(r["t_features"] + r["t_events"] for r in res)
        t_cap = sum(r["t_captions"] for r in res)
        hist = torch.bincount(torch.cat([r["event_classes"] for r in res]), minlength=11).tolist()
        print(json.dumps({
            "workload": "configs[4]: %d synthetic videos x %d frames @224x224 -> DenseNet-121 features -> BiGRU(128) events on "
                        "stride-1 windows of %d -> %d-frame points -> GNMT LSTM captions (beam 5, max 50 tokens)" % (
                            args.videos, args.frames, args.window, args.segment),
            "n_gpus": world, "frames": frames, "caption_tokens": tokens, "segments": sum(len(r["captions"]) for r in res),
            "wall_s": wall, "frames_per_s_end_to_end": frames / wall,
            "detector_frames_per_s": frames / t_det, "caption_tokens_per_s": tokens / max(t_cap, 1e-9),
            "host_to_host_latency_s_per_video": {"mean": sum(r["t_total"] for r in res) / len(res), "max": max(r["t_total"] for r in res)},
            "stage_seconds": {"features": sum(r["t_features"] for r in res), "store": sum(r["t_store"] for r in res),
                              "events": sum(r["t_events"] for r in res), "captions": t_cap},
            "event_class_histogram": hist, "data": "synthetic", "weights": "seeded synthetic (tennis_b200/synthetic.py)"}))
    if world > 1:
        dist.destroy_process_group()
    if tmp is not None:
        tmp.cleanup()


if __name__ == "__main__":
    main()
