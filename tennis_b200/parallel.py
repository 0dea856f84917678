"""Multi-GPU and host-pipelined execution of the CNN -> temporal-head path.

One process per GPU (torch.distributed, NCCL over NVLink).  Frames are independent through the CNN
(inference BN uses running statistics), so the flattened frame axis is sharded contiguously across ranks
(SURVEY.md §8e): rank r runs the backbone on its shard, ONE collective — an all-gather of the per-frame
features — assembles (F, D) on every rank, and the (tiny) temporal head runs on the gathered features.
This replaces the reference's single-process `split_and_load` loop over contexts (train.py:410-419).
"""
import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated afterwards are
    node-local and H2D copies do not cross the socket interconnect (one process per GPU: eight ranks otherwise share whatever
    node the launcher started them on).  Best effort: returns the node id, or None when the topology cannot be read."""
    import os
    try:
        props = torch.cuda.get_device_properties(int(device_index))  # CUDA ordinal (honours CUDA_VISIBLE_DEVICES)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # no sysfs topology (containers) or no PCI ids: leave the affinity alone
        return None


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` for `rank`; the last rank takes the remainder
    (mirrors gluon.utils.split_and_load(even_split=False))."""
    per = n_items // world
    lo = rank * per
    hi = n_items if rank == world - 1 else lo + per
    return lo, hi


def balanced_range(n_items, rank, world):
    """Contiguous shard [lo, hi) whose sizes differ by at most one item (first n % world ranks get the extra one): the split used
    for FRAME sharding, where any frame count must work (odd B*T, world sizes that do not divide it)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_ragged(local, n_total, world=None, group=None):
    """All-gather row blocks of UNEQUAL size: rank r holds rows balanced_range(n_total, r, world) of a (n_total, ...) tensor.  Blocks
    are padded to the largest one so that ONE all_gather_into_tensor moves them (NVSwitch collectives are latency-bound: one
    padded exchange beats `world` broadcasts), then the padding rows are dropped.  Returns (n_total, ...) on every rank."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    per = -(-n_total // world)
    tail = tuple(local.shape[1:])
    if local.shape[0] == per and n_total % world == 0:
        return all_gather_rows(local, world, group)
    padded = local.new_zeros((per,) + tail)
    padded[:local.shape[0]].copy_(local)
    out = local.new_empty((world * per,) + tail)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = balanced_range(n_total, r, world)
        parts.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(parts, 0)


def all_gather_rows(local, world=None, group=None):
    """All-gather equal-sized row blocks (n_local, D) -> (world*n_local, D) on every rank."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def sum_gradients_across_ranks(grads, group=None):
    """Sum each gradient tensor over the ranks, in place (the reference's Trainer on KVStore('device') sums the per-device
    gradients before `rescale_grad = 1/batch_size` is applied: train.py:298-299,424).  One flat bucket per call:
    NVSwitch collectives are latency-, not link-bound, so the ~3 MB of head gradients go out as a single all-reduce."""
    grads = [g for g in grads if g is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].reshape(g.shape))
        off += n
    return grads


class ShardedCNNRNN(object):
    """Frame-sharded forward of a CNNRNN / TemporalPooling-style model: `model.td.model` is the per-frame
    feature extractor, `head(features (B,T,D)) -> logits` the temporal head.

    forward(local_clips): local_clips is this rank's (B_local, T, 3, H, W) slice of the global batch -> logits of the GLOBAL
    batch (B_local*world, C) on every rank.  forward_frames(frames_local, B, T): the general form -- this rank holds frames
    balanced_range(B*T, rank, world) of the flattened (B*T, ...) batch, so shards may cut through clips and B need not divide.

    Collectives per step: ONE all-gather of the bf16 per-frame features (the exchange north_star names), then every rank runs the
    temporal head on ITS OWN clips only and the (B, C) logits are all-gathered (a few KB) -- the head is not recomputed W times.
    """

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._even_checked = set()

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

    def head_sharded(self, gathered, B, T):
        """gathered: (B*T, D) features of the global batch (bf16 twin or fp32) on every rank -> (B, C) global logits: each rank
        runs the head on clips balanced_range(B, rank, world) and the logits are exchanged."""
        D = gathered.shape[1]
        lo, hi = balanced_range(B, self.rank, self.world)
        mine = gathered[lo * T:hi * T].reshape(hi - lo, T, D)
        if hi > lo:
            logits = self.head(mine, None)
        else:  # more ranks than clips: nothing to do here, but the exchange below needs the logit width
            logits = self.head(gathered[:T].reshape(1, T, D), None)[:0]
        return all_gather_ragged(logits.contiguous(), B, self.world, self.group)

    def forward_frames(self, frames_local, B, T):
        feats = self.model.td.model(frames_local)
        twin = getattr(feats, "_tn_bf16", None)
        src = twin if twin is not None else feats
        if self.world == 1:
            D = feats.shape[1]
            return self.head(feats.reshape(B, T, D), None if twin is None else twin.reshape(B, T, D))
        return self.head_sharded(all_gather_ragged(src, B * T, self.world, self.group), B, T)

    def forward(self, clips):
        B, T = clips.shape[:2]
        feats, twin = self.features_local(clips)
        if self.world == 1:
            D = feats.shape[1]
            return self.head(feats.reshape(B, T, D), None if twin is None else twin.reshape(B, T, D))
        # forward() is the EVEN-shard form (every rank holds B clips); unequal shards would enter differently sized collectives and
        # hang every rank, so the first call with a given B checks it (one 2-element all-reduce) -- ragged batches: forward_frames()
        if B not in self._even_checked:
            chk = torch.tensor([B, -B], dtype=torch.int64, device=clips.device)
            dist.all_reduce(chk, op=dist.ReduceOp.MAX, group=self.group)
            if int(chk[0]) != -int(chk[1]):
                raise ValueError("ShardedCNNRNN.forward needs the same number of clips on every rank (%d here, %d..%d over the ranks); "
                                 "use forward_frames() for ragged shards" % (B, -int(chk[1]), int(chk[0])))
            self._even_checked.add(B)
        # the only large forward collective: per-frame features, exchanged in the dtype the head consumes
        src = twin if twin is not None else feats
        return self.head_sharded(all_gather_rows(src, self.world, self.group), B * self.world, T)

    __call__ = forward


def modelled_makespan(sizes, copy_ms, fixed_ms, compute_ms):
    """Makespan of a chunked copy/compute pipeline: chunk k's kernels start when its copy has landed and chunk k-1's
    kernels are done; copies run back to back at `copy_ms` per item, a chunk of n items computes in fixed_ms + n*compute_ms."""
    t_copy = t_done = 0.0
    for n in sizes:
        t_copy += copy_ms * n
        t_done = max(t_done, t_copy) + fixed_ms + compute_ms * n
    return t_done


def plan_chunks(B, copy_ms, fixed_ms, compute_ms, max_chunks=8, units=32):
    """Chunk boundaries [0, ..., B] minimising `modelled_makespan` (hill climbing over a coarse grid of `units`)."""
    U = max(1, min(units, B))

    def cost(sz):
        return modelled_makespan([B * u / float(U) for u in sz], copy_ms, fixed_ms, compute_ms)

    best = None
    for k in range(1, min(max_chunks, U) + 1):
        seeds = [[U // k + (1 if i < U % k else 0) for i in range(k)]]
        w = list(range(1, k + 1))
        arith = [max(1, int(U * x / float(sum(w)))) for x in w]
        arith[-1] += U - sum(arith)
        if arith[-1] >= 1:
            seeds.append(arith)
        for sz in seeds:
            sz = list(sz)
            c = cost(sz)
            improved = True
            while improved:
                improved = False
                for i in range(len(sz)):
                    for j in range(len(sz)):
                        if i == j or sz[i] <= 1:
                            continue
                        sz[i] -= 1
                        sz[j] += 1
                        c2 = cost(sz)
                        if c2 < c - 1e-9:
                            c, improved = c2, True
                        else:
                            sz[i] += 1
                            sz[j] -= 1
            if best is None or c < best[0] - 1e-9:
                best = (c, list(sz))
    cuts, acc = [0], 0
    for u in best[1]:
        acc += u
        nxt = B if acc == U else int(round(B * acc / float(U)))
        if nxt > cuts[-1]:
            cuts.append(nxt)
    if cuts[-1] != B:
        cuts.append(B)
    return cuts


class HostPipeline(object):
    """End-to-end step from HOST memory: pinned host clips -> chunked H2D on a copy stream overlapped with the
    backbone of the previous chunks -> temporal head -> logits back on the host.  This is the call a user of the
    reference makes when they write `split_and_load(batch) ; model(x) ; out.asnumpy()` (train.py:410-431).

    `forward(x)` is that synchronous call.  `submit(x)` / `result(h)` split it so a loader loop can keep one step in
    flight (the H2D of step s+1 overlaps the kernels of step s, like MXNet's asynchronous engine does for the
    reference's loop); at most two steps may be outstanding."""

    def __init__(self, sharded, chunks=8):
        self.sh = sharded
        self.chunks = chunks
        self.copy_stream = torch.cuda.Stream()
        self._dev_bufs = None
        self._consumed = [None, None]   # per device buffer: event recorded after its last consumer kernel
        self._turn = 0
        self._host_out = [None, None]   # pinned (B*world, C) result buffers, one per in-flight step
        self._model = {}                # (shape, dtype) -> (copy_ms, fixed_ms, compute_ms) per clip, measured once

    def _bufs(self, shape, dtype, device):
        key = (tuple(shape), dtype)
        if self._dev_bufs is None or self._dev_bufs[0] != key:
            torch.cuda.current_stream().synchronize()
            self._dev_bufs = (key, [torch.empty(shape, dtype=dtype, device=device) for _ in range(2)])
            self._consumed = [None, None]
        return self._dev_bufs[1]

    def _calibrate(self, clips_host, buf):
        """Measure, once per input format, the three numbers the chunk plan needs: host->device ms per clip and the
        backbone's fixed + per-clip ms (two chunk sizes).  Host->device of a 224x224 fp32 frame takes about as long as
        its DenseNet forward, a uint8 frame a quarter of that, so the best plan depends on the format."""
        B = clips_host.shape[0]
        main = torch.cuda.current_stream()

        def timed(fn):
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            fn()
            e1.record(main)
            e1.synchronize()
            return e0.elapsed_time(e1)

        nc = max(1, B // 8)
        copy_ms = timed(lambda: buf[:nc].copy_(clips_host[:nc], non_blocking=True)) / nc
        n1, n2 = max(1, B // 16), max(2, B // 4)
        if n2 <= n1 or B < 4:
            return copy_ms, 0.0, timed(lambda: self.sh.features_local(buf[:B])) / B
        t1 = timed(lambda: self.sh.features_local(buf[:n1]))
        t2 = timed(lambda: self.sh.features_local(buf[:n2]))
        per = max(1e-6, (t2 - t1) / (n2 - n1))
        return copy_ms, max(0.0, t1 - per * n1), per

    def _schedule(self, B, model=None):
        if model is None or self.chunks <= 1:
            n = max(1, min(self.chunks, B))
            return sorted(set(int(round(B * k / float(n))) for k in range(n + 1)))
        return plan_chunks(B, model[0], model[1], model[2], max_chunks=self.chunks)

    def submit(self, clips_host):
        """Queue one step; returns a handle for `result`.  clips_host: pinned (B,T,3,H,W) fp32 [or (B,T,H,W,3) uint8]."""
        assert not clips_host.is_cuda
        B, T = clips_host.shape[:2]
        dev = torch.device("cuda", torch.cuda.current_device())
        main = torch.cuda.current_stream()
        slot = self._turn & 1
        self._turn += 1
        buf = self._bufs(tuple(clips_host.shape), clips_host.dtype, dev)[slot]
        key = (tuple(clips_host.shape), clips_host.dtype)
        if key not in self._model:
            if self._consumed[slot] is not None:
                self._consumed[slot].synchronize()
            self._model[key] = self._calibrate(clips_host, buf) if self.chunks > 1 and B >= 4 else None
        bounds = self._schedule(B, self._model[key])
        self.last_plan = {"cuts": bounds, "model_ms_per_clip": self._model[key]}  # (copy, fixed, compute); for reports
        nch = len(bounds) - 1
        ready = [torch.cuda.Event() for _ in range(nch)]
        with torch.cuda.stream(self.copy_stream):
            if self._consumed[slot] is not None:  # the step that last used this buffer must have read it
                self.copy_stream.wait_event(self._consumed[slot])
            for i in range(nch):
                lo, hi = bounds[i], bounds[i + 1]
                buf[lo:hi].copy_(clips_host[lo:hi], non_blocking=True)
                ready[i].record(self.copy_stream)
        feats_all, twin_all = [], []
        for i in range(nch):
            main.wait_event(ready[i])
            f, t = self.sh.features_local(buf[bounds[i]:bounds[i + 1]])
            feats_all.append(f)
            twin_all.append(t)
        done = torch.cuda.Event()
        done.record(main)
        self._consumed[slot] = done
        feats = torch.cat(feats_all, 0)
        twin = torch.cat(twin_all, 0) if twin_all[0] is not None else None
        if self.sh.world > 1:
            src = twin if twin is not None else feats
            g = all_gather_rows(src, self.sh.world, self.sh.group)
            logits = self.sh.head_sharded(g, B * self.sh.world, T)
        else:
            logits = self.sh.head(feats.reshape(B, T, -1), None if twin is None else twin.reshape(B, T, -1))
        host = self._host_out[slot]  # pinned result buffers are allocated once per slot, not per step (cudaHostAlloc is slow)
        if host is None or host.shape != logits.shape or host.dtype != logits.dtype:
            host = self._host_out[slot] = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
        host.copy_(logits, non_blocking=True)  # device -> host read of the step's result, queued behind the head
        fin = torch.cuda.Event()
        fin.record(main)
        return (host, fin, logits)

    @staticmethod
    def result(handle):
        """Wait for a submitted step and return its (B*world, C) fp32 host logits (the sync point, train.py:427-431)."""
        host, fin, _ = handle
        fin.synchronize()
        return host

    def forward(self, clips_host):
        """clips_host: pinned (B,T,3,H,W) fp32 [or (B,T,H,W,3) uint8] host tensor -> (B*world, C) fp32 host tensor."""
        return self.result(self.submit(clips_host))

    __call__ = forward
