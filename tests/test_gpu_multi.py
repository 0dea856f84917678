"""Multi-GPU (NCCL) behaviour of the product path; needs >= 2 GPUs (skipped on the 1-GPU test box, run with `gpurun --gpus 2`).
  * frame-sharded inference at world 2 returns logits BIT-IDENTICAL to world 1, also when the frame count does not
    divide by the world size;
  * one SGD step of the temporal head at world 2 (clip shards + NCCL gradient sum in Trainer.step) equals the single-process step
    on the whole batch."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import bench
from tennis_b200 import autograd
from tennis_b200.gluon import SoftmaxCrossEntropyLoss, Trainer
from tennis_b200.parallel import ShardedCNNRNN, balanced_range, shard_range
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model = bench.build_model(dev)
sh = ShardedCNNRNN(model)
g = torch.Generator().manual_seed(5)
B, T = 8, 3
clips = torch.randn(B, T, 3, 224, 224, generator=g)
# (1) EVEN clip shards through forward() (its contract: 8 clips over 1/2/4/8 ranks); (2) ragged FRAME shards through
# forward_frames(): 7 clips x 3 frames = 21 frames never divide evenly
lo, hi = shard_range(B, rank, world)
out = sh(clips[lo:hi].to(dev))
odd = torch.randn(7, 3, 3, 224, 224, generator=g)
flat = odd.reshape(21, 3, 224, 224)
flo, fhi = balanced_range(21, rank, world)
out_odd = sh.forward_frames(flat[flo:fhi].to(dev), 7, 3)
# (3) one SGD step of the head on features: clip shards, gradients summed over ranks
feats = torch.randn(8, T, 1024, generator=g).relu()
labels = torch.arange(8) %% 11
from tennis_b200.models.vision.definitions import CNNRNN
head = CNNRNN(None, 11, hidden_size=128, type="gru")
head.initialize(ctx=dev)
head(feats[:1].to(dev))
tr = Trainer(head.collect_params(), "sgd", {"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4})
l2, h2 = shard_range(8, rank, world)
with autograd.record():
    loss = SoftmaxCrossEntropyLoss()(head(feats[l2:h2].to(dev)), labels[l2:h2].to(dev))
autograd.backward([loss])
tr.step(8)
w = torch.cat([p.data().reshape(-1).cpu() for _, p in sorted(head.collect_params().items())])
torch.cuda.synchronize()
if rank == 0:
    torch.save({"out": out.cpu(), "out_odd": out_odd.cpu(), "w": w}, sys.argv[1])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
"""


_CACHE = {}


def _run(world, out_path, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, TN_INIT_SEED="3")
    if world == 1:
        env.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
        cmd = [sys.executable, str(script), out_path]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
               "127.0.0.1", "--master-port", "29541", str(script), out_path]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return torch.load(out_path)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_world_n_inference_bit_identical_and_sgd_step_equal(world, tmp_path):
    """W in {1, 2, 4, 8} give bit-identical logits (also with ragged frame / clip shards and ranks without work: 6 and
    7 clips, 21 frames, 8 training clips over up to 8 ranks)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs >= %d GPUs" % world)
    if "one" not in _CACHE:  # the single-process result is the same for every world size
        _CACHE["one"] = _run(1, str(tmp_path / "w1.pt"), tmp_path)
    one = _CACHE["one"]
    two = _run(world, str(tmp_path / "wn.pt"), tmp_path)
    assert one["out"].shape == (8, 11) and torch.equal(one["out"], two["out"])
    assert one["out_odd"].shape == (7, 11) and torch.equal(one["out_odd"], two["out_odd"])
    # gradient sums associate differently across ranks (fp32): equal to rounding, not bit-identical
    assert (one["w"] - two["w"]).abs().max().item() < 1e-5 * max(1.0, one["w"].abs().max().item())
