"""Multi-GPU and host-pipelined execution of the CNN -> temporal-head path.

One process per GPU (torch.distributed, NCCL over NVLink).  Frames are independent through the CNN
(inference BN uses running statistics), so the flattened frame axis is sharded contiguously across ranks
(SURVEY.md §8e): rank r runs the backbone on its shard, ONE collective — an all-gather of the per-frame
features — assembles (F, D) on every rank, and the (tiny) temporal head runs on the gathered features.
This replaces the reference's single-process `split_and_load` loop over contexts (train.py:410-419).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` for `rank`; the last rank takes the remainder
    (mirrors gluon.utils.split_and_load(even_split=False))."""
    per = n_items // world
    lo = rank * per
    hi = n_items if rank == world - 1 else lo + per
    return lo, hi


def all_gather_rows(local, world=None, group=None):
    """All-gather equal-sized row blocks (n_local, D) -> (world*n_local, D) on every rank."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class ShardedCNNRNN(object):
    """Frame-sharded forward of a CNNRNN / TemporalPooling-style model: `model.td.model` is the per-frame
    feature extractor, `head(features (B,T,D)) -> logits` the temporal head.

    forward(local_clips): local_clips is this rank's (B_local, T, 3, H, W) slice of the global batch; returns the
    logits of the GLOBAL batch (B_local*world, C) on every rank.
    """

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def features_local(self, clips):
        B, T = clips.shape[:2]
        feats = self.model.td.model(clips.reshape((B * T,) + tuple(clips.shape[2:])))
        return feats, getattr(feats, "_tn_bf16", None)

    def head(self, feats_bt, twin_bt=None):
        m = self.model
        if twin_bt is not None:
            feats_bt._tn_bf16 = twin_bt
        if hasattr(m, "rnn"):
            y = m.rnn.forward_max(feats_bt)
        else:
            from . import ops
            y = ops.temporal_pool(feats_bt, 'mean' if m.pool == 'mean' else 'max')
        return m.classes(y) if m.classes else y

    def forward(self, clips):
        B, T = clips.shape[:2]
        feats, twin = self.features_local(clips)
        if self.world > 1:
            # the only forward collective: per-frame features, exchanged in the dtype the head consumes
            if twin is not None:
                twin = all_gather_rows(twin, self.world, self.group)
                feats = twin  # fp32 copy is not needed by the head
            else:
                feats = all_gather_rows(feats, self.world, self.group)
        Bg = B * self.world
        D = feats.shape[1]
        f_bt = feats.reshape(Bg, T, D)
        t_bt = None if twin is None else twin.reshape(Bg, T, D)
        if twin is not None and self.world > 1:
            return self.head(t_bt, None)
        return self.head(f_bt, t_bt)

    __call__ = forward


class HostPipeline(object):
    """End-to-end step from HOST memory: pinned host clips -> chunked H2D on a copy stream overlapped with the
    backbone of the previous chunk -> temporal head -> logits back on the host.  This is the call a user of the
    reference makes when they write `split_and_load(batch) ; model(x) ; out.asnumpy()` (train.py:410-431)."""

    def __init__(self, sharded, chunks=4):
        self.sh = sharded
        self.chunks = chunks
        self.copy_stream = torch.cuda.Stream()
        self._dev_bufs = None

    def _bufs(self, shape, dtype, device):
        key = (tuple(shape), dtype)
        if self._dev_bufs is None or self._dev_bufs[0] != key:
            self._dev_bufs = (key, [torch.empty(shape, dtype=dtype, device=device) for _ in range(2)])
        return self._dev_bufs[1]

    def _schedule(self, B):
        if B < 16 or self.chunks <= 1:
            n = max(1, min(self.chunks, B))
            per = (B + n - 1) // n
            return [min(B, i * per) for i in range(n)] + [B]
        cuts, size, lo = [0], max(1, B // 16), 0
        while lo < B:
            lo = lo + size if B - (lo + size) >= size else B  # fold a short tail into the last chunk
            cuts.append(lo)
            size *= 2
        return cuts

    def forward(self, clips_host):
        """clips_host: pinned (B,T,3,H,W) fp32 [or (B,T,H,W,3) uint8] host tensor -> (B*world, C) fp32 host tensor."""
        assert not clips_host.is_cuda
        B, T = clips_host.shape[:2]
        dev = torch.device("cuda", torch.cuda.current_device())
        # chunk schedule: a small first chunk so compute starts after ~1/16 of the copy, then growing chunks so the
        # kernels keep large grids (H2D of chunk i+1 overlaps the backbone of chunk i)
        bounds = self._schedule(B)
        nch = len(bounds) - 1
        per = max(bounds[i + 1] - bounds[i] for i in range(nch))
        main = torch.cuda.current_stream()
        bufs = self._bufs((per,) + tuple(clips_host.shape[1:]), clips_host.dtype, dev)
        feats_all, twin_all = [], []
        ready = [torch.cuda.Event() for _ in range(nch)]
        consumed = [torch.cuda.Event() for _ in range(nch)]
        for i in range(nch):
            lo, hi = bounds[i], bounds[i + 1]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(consumed[i - 2])
                bufs[i % 2][: hi - lo].copy_(clips_host[lo:hi], non_blocking=True)
                ready[i].record(self.copy_stream)
            main.wait_event(ready[i])
            f, t = self.sh.features_local(bufs[i % 2][: hi - lo])
            consumed[i].record(main)
            feats_all.append(f)
            twin_all.append(t)
        feats = torch.cat(feats_all, 0)
        twin = torch.cat(twin_all, 0) if twin_all[0] is not None else None
        if self.sh.world > 1:
            src = twin if twin is not None else feats
            g = all_gather_rows(src, self.sh.world, self.sh.group)
            Bg = B * self.sh.world
            logits = self.sh.head(g.reshape(Bg, T, -1), None)
        else:
            logits = self.sh.head(feats.reshape(B, T, -1), None if twin is None else twin.reshape(B, T, -1))
        return logits.cpu()  # device -> host read of the step's result (sync point, train.py:427-431)

    __call__ = forward
