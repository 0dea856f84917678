"""CPU-only: C-ABI surface, host-side mirrors of the reference interface, checkpoint codec, sharding logic."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tennis_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tennis_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from tennis_b200 import _build
        _build.build()
    handle = _lib.lib()  # binds every signature; AttributeError on a missing symbol
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), "library does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes table lacks %s" % n
    assert handle.tn_version() >= 100


def test_no_gpu_means_loud_failure_not_fallback():
    from tennis_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().tn_device_check(0) == _lib.TN_ERR_ARCH
    with pytest.raises(_lib.TennisB200Error):
        ops.dense(torch.zeros(2, 4), torch.zeros(3, 4))
    with pytest.raises(_lib.TennisB200Error):
        ops.Backbone("densenet121", torch.zeros(_lib.lib().tn_backbone_param_count(0)))


def test_param_count_and_feature_dim_queries():
    from tennis_b200 import _lib
    L = _lib.lib()
    assert L.tn_backbone_param_count(_lib.ARCH_DENSENET121) == 7037504
    assert L.tn_backbone_param_count(_lib.ARCH_RESNET18_V2) == 11182796
    assert L.tn_backbone_feature_dim(_lib.ARCH_DENSENET121, 224, 224) == 1024
    assert L.tn_backbone_feature_dim(_lib.ARCH_DENSENET121, 512, 512) == 4096  # train.py:259
    assert L.tn_backbone_feature_dim(_lib.ARCH_RESNET18_V2, 224, 224) == 512


def test_model_zoo_inventory_matches_oracle_order():
    from oracle import vision as O
    from tennis_b200 import model_zoo
    for name, arch in (("DenseNet121", "densenet121"), ("resnet18_v2", "resnet18_v2")):
        feats = model_zoo.get_model(name).features
        assert list(feats.collect_params().keys()) == [n for n, _ in O.PARAM_SHAPES[arch]()]
        feats.initialize(ctx="cpu")
        assert feats.flat_params().numel() == O.flatten_params(arch, O.synthetic_params(arch)).numel()


def test_reference_model_assembly_and_param_names(tmp_path):
    """train.py:204-236 assembly; structural names as Gluon's save_parameters writes them."""
    from tennis_b200 import model_zoo
    from tennis_b200.models.vision.definitions import CNNRNN, FrameModel, TemporalPooling
    backbone = model_zoo.get_model("DenseNet121", pretrained=False).features
    fm = FrameModel(backbone, 11)
    model = CNNRNN(fm, 11, hidden_size=128, type="gru")
    names = list(model.collect_params().keys())
    assert "td.model.conv0.weight" in names and "rnn.l0_i2h_weight" in names and "rnn.r0_h2h_bias" in names
    assert "classes.weight" in names and "classes.bias" in names
    assert model.feats is False and CNNRNN(None, 11).feats is True
    tp = TemporalPooling(fm, num_classes=0, pool="mean")
    assert tp.classes is fm.classes and tp.td.model is fm.backbone  # definitions.py:53-55
    # freeze_backbone idiom (train.py:231-233)
    for prm in fm.backbone.collect_params().values():
        prm.grad_req = "null"
    model.initialize(ctx="cpu")
    # deferred shapes are resolved on first use; give them explicitly to round-trip a checkpoint
    for d in ("l0", "r0"):
        model.rnn._reg_params[d + "_i2h_weight"]._finish_deferred((384, 1024))
    model.classes.weight._finish_deferred((11, 256))
    path = str(tmp_path / "0003.params")
    model.save_parameters(path)
    model2 = CNNRNN(FrameModel(model_zoo.get_model("DenseNet121").features, 11), 11, hidden_size=128, type="gru")
    model2.load_parameters(path, ctx="cpu")
    a, b = model.collect_params(), model2.collect_params()
    for k in a:
        assert torch.equal(a[k].data().cpu(), b[k].data().cpu()), k


def test_params_codec_roundtrip_and_layout(tmp_path):
    from tennis_b200 import params_io
    arrays = {"a.weight": np.arange(12, dtype=np.float32).reshape(3, 4), "b": np.array([1, 2, 3], dtype=np.int32)}
    path = str(tmp_path / "x.params")
    params_io.save(path, arrays)
    raw = open(path, "rb").read()
    assert raw[:8] == (0x112).to_bytes(8, "little")  # NDArray-list magic
    back = params_io.load(path)
    assert list(back) == list(arrays)
    for k in arrays:
        assert back[k].dtype == arrays[k].dtype and np.array_equal(back[k], arrays[k])


def test_params_reader_parses_hand_assembled_ndarray_list(tmp_path):
    """The reader against bytes assembled field by field from the layout in SURVEY.md 8f-1 (independent of our writer):
    list header, one V2 and one V3 dense record, Gluon-style names with the legacy `arg:` / `aux:` prefixes."""
    import struct
    from tennis_b200 import params_io
    w = np.arange(6, dtype="<f4").reshape(2, 3)
    m = np.array([5, 7], dtype="<i8")
    raw = struct.pack("<QQQ", 0x112, 0, 2)
    raw += struct.pack("<Ii", 0xF993FAC9, 0) + struct.pack("<I", 2) + struct.pack("<2q", 2, 3) + struct.pack("<iii", 1, 0, 0) + w.tobytes()
    raw += struct.pack("<Ii", 0xF993FACA, 0) + struct.pack("<I", 1) + struct.pack("<1q", 2) + struct.pack("<iii", 2, 3, 6) + m.tobytes()
    raw += struct.pack("<Q", 2)
    for name in (b"arg:classes.weight", b"aux:bn0.running_mean"):
        raw += struct.pack("<Q", len(name)) + name
    path = str(tmp_path / "hand.params")
    open(path, "wb").write(raw)
    got = params_io.load(path)
    assert list(got) == ["classes.weight", "bn0.running_mean"]
    assert got["classes.weight"].dtype == np.float32 and np.array_equal(got["classes.weight"], w)
    assert got["bn0.running_mean"].dtype == np.int64 and np.array_equal(got["bn0.running_mean"], m)
    bad = bytearray(raw)
    bad[0] = 0x13
    open(path, "wb").write(bytes(bad))
    import pytest as _pt
    with _pt.raises(ValueError):
        params_io.load(path)


def test_time_distributed_folds_and_unfolds():
    from tennis_b200.gluon import Block
    from tennis_b200.utils.layers import TimeDistributed

    class Fake(Block):
        def forward(self, x):
            return x.sum(dim=(2, 3), keepdim=True)[:, :1].repeat(1, 4, 1, 1), x.mean(dim=(1, 2, 3))

    td = TimeDistributed(Fake())
    out = td(torch.ones(3, 2, 3, 2, 2))
    assert isinstance(out, tuple) and out[0].shape == (3, 2, 4, 1, 1)  # shape KAT of definitions.py:166-167
    assert out[1].shape == (3, 2)  # tuple outputs are unfolded element-wise (layers.py:41-42)


def test_shard_range_matches_split_and_load():
    from tennis_b200.parallel import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 10)]  # last takes the remainder
    assert [shard_range(8192, r, 8) for r in range(8)][3] == (3072, 4096)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from tennis_b200.parallel import all_gather_rows, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["MASTER_PORT"], rank=rank, world_size=world)
F, D = 12, 5
full = torch.arange(F * D, dtype=torch.float32).reshape(F, D)
lo, hi = shard_range(F, rank, world)
got = all_gather_rows(full[lo:hi].clone(), world)
assert torch.equal(got, full), got
dist.destroy_process_group()
print("ok", rank)
"""


def test_feature_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    env = dict(os.environ, WORLD_SIZE="2", MASTER_PORT="29611", MASTER_ADDR="127.0.0.1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


_GLOO_GRAD_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from tennis_b200.parallel import sum_gradients_across_ranks
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
g = torch.Generator().manual_seed(7)
base = [torch.randn(5, 3, generator=g), torch.randn(4, generator=g), torch.randn(2, 2, 2, generator=g)]
grads = [b * (rank + 1) for b in base] + [None]        # rank r holds (r+1) * base; a parameter without gradient is skipped
out = sum_gradients_across_ranks(grads)
assert len(out) == 3
for o, b in zip(out, base):
    assert torch.allclose(o, b * sum(range(1, world + 1))), (rank, o, b)
assert grads[0] is out[0]                               # summed in place
dist.destroy_process_group()
"""


def test_gradient_sum_world_size_2_gloo(tmp_path):
    """Trainer.step's cross-rank step: gradients are SUMMED over the ranks before rescale_grad (reference KVStore('device'))."""
    script = tmp_path / "g.py"
    script.write_text(_GLOO_GRAD_WORKER % ROOT)
    env = dict(os.environ, WORLD_SIZE="2", MASTER_PORT="29613", MASTER_ADDR="127.0.0.1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    from tennis_b200.parallel import sum_gradients_across_ranks
    g = [torch.ones(3)]
    assert sum_gradients_across_ranks(g)[0] is g[0] and torch.equal(g[0], torch.ones(3))  # single process: untouched


_GLOO_RAGGED_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from tennis_b200.parallel import all_gather_ragged, balanced_range, ShardedCNNRNN
from tennis_b200 import cli
from tennis_b200.metrics.vision import PRF1, Accuracy
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# (1) frame shards that do not divide: 7 clips x 5 frames = 35 rows over 3 ranks (12/12/11), and fewer rows than ranks
for F in (35, 2, 3, 0 + world):
    full = torch.arange(F * 4, dtype=torch.float32).reshape(F, 4)
    lo, hi = balanced_range(F, rank, world)
    got = all_gather_ragged(full[lo:hi].clone(), F, world)
    assert torch.equal(got, full), (F, rank, got)
# (2) the sharded head: every rank runs the head on its own clips only, logits are exchanged -> identical to one process
class _Head(object):
    classes = None
    pool = 'max'
class _Sh(ShardedCNNRNN):
    def head(self, f_bt, twin=None):
        return f_bt.max(dim=1).values[:, :3] * 2.0      # stand-in temporal head (B,T,D)->(B,3)
sh = _Sh(_Head())
B, T, D = 7, 5, 4
g = torch.Generator().manual_seed(3)
feats = torch.randn(B * T, D, generator=g)
want = feats.reshape(B, T, D).max(dim=1).values[:, :3] * 2.0
got = sh.head_sharded(feats, B, T)
assert torch.equal(got, want), (rank, got, want)
# (3) script plumbing: contiguous batch shards (last rank takes the remainder) and metric accumulators summed over ranks
items = list(range(10))
mine = cli.rank_shard(items)
assert sum(dist_len for dist_len in [len(mine)]) >= 10 // world
labels = torch.tensor(items) %% 3
preds = torch.nn.functional.one_hot((torch.tensor(items) * 2) %% 3, 3).float()
acc, prf = Accuracy(), PRF1(label_names=['a', 'b', 'c'])
lo = items.index(mine[0]) if mine else 0
sl = slice(lo, lo + len(mine))
if mine:
    acc.update([labels[sl]], [preds[sl]]); prf.update([labels[sl]], [preds[sl]])
cli.sync_metrics([acc, prf])
acc1, prf1 = Accuracy(), PRF1(label_names=['a', 'b', 'c'])
acc1.update([labels], [preds]); prf1.update([labels], [preds])
assert acc.get() == acc1.get() and (prf.mat == prf1.mat).all() and (prf.scores == prf1.scores).all()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 3])
def test_ragged_gather_sharded_head_and_metric_sync_gloo(tmp_path, world):
    """Frame sharding for frame counts the world size does not divide, the head on the rank's own clips +
    logits exchange, and the scripts' rank sharding / metric reduction, on gloo."""
    script = tmp_path / "r.py"
    script.write_text(_GLOO_RAGGED_WORKER % ROOT)
    env = dict(os.environ, WORLD_SIZE=str(world), MASTER_PORT=str(29620 + world), MASTER_ADDR="127.0.0.1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_sequence_reverse_index_matches_oracle():
    """models/captioning/train_graph.py::_reverse_index is SequenceReverse(use_sequence_length=True) as a gather (A.4)."""
    from oracle.captioning import _sequence_reverse
    from tennis_b200.models.captioning.train_graph import _reverse_index, _seq_reverse
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 7, 3, generator=g)
    lens = torch.tensor([7, 1, 4, 6], dtype=torch.int32)
    idx = _reverse_index(lens, 7)
    got = _seq_reverse(x, idx)
    assert torch.equal(got, _sequence_reverse(x, lens.float()))
    assert torch.equal(_seq_reverse(got, idx), x)  # an involution: the same gather routes gradients back


def test_numa_binding_is_best_effort_without_gpu():
    from tennis_b200.parallel import bind_to_gpu_numa_node
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), int)
    assert os.sched_getaffinity(0) <= before


def test_checkpoint_directory_conventions(tmp_path):
    """cli.latest_params / best_epoch: newest NNNN.params wins and valid_best.params is never taken for a resume point
    (train.py:286-295, train_gnmt.py:236-247); the best epoch is the highest AVG_NB_f1 in scores.txt (train.py:334-346)."""
    from tennis_b200 import cli
    d = tmp_path / "exp"
    assert cli.latest_params(str(d)) == (None, 0) and cli.best_epoch(str(d)) is None
    d.mkdir()
    for name in ("0000.params", "0003.params", "0011.params", "valid_best.params", "log.txt"):
        (d / name).write_bytes(b"x")
    path, start = cli.latest_params(str(d))
    assert path.endswith("0011.params") and start == 12
    (d / "scores.txt").write_text("0000 0.1000\n0003 0.4100\n0011 0.3900\n")
    assert cli.best_epoch(str(d)) == 3


@pytest.mark.parametrize("arch,first_bn", [("densenet121", "4.0.1.0.gamma"), ("resnet18_v2", "5.0.bn1.gamma")])
def test_gluoncv_structural_names_cover_the_inventory_and_load(arch, first_bn, tmp_path):
    """A checkpoint keyed the way the reference's save_parameters keys it (GluonCV child indices, e.g. the documented
    `td.model.4.0.1.0.gamma`) loads into our models: the name table is a bijection onto the inventory, in order."""
    from oracle import vision as O
    from tennis_b200 import gluoncv_names, model_zoo, params_io
    from tennis_b200.models.vision.definitions import CNNRNN, FrameModel
    table = gluoncv_names.NAME_MAPS[arch]()
    ours = [n for n, _ in O.PARAM_SHAPES[arch]()]
    assert list(table.values()) == ours and len(set(table)) == len(ours)
    assert first_bn in table and table[first_bn].endswith("bn1.gamma")
    p = O.synthetic_params(arch, seed=5)
    inv = {v: k for k, v in table.items()}
    for prefix, build in (("backbone.", lambda bb: FrameModel(bb, -1)), ("td.model.", lambda bb: CNNRNN(FrameModel(bb, -1), -1))):
        model = build(model_zoo.get_model(arch).features)
        model.initialize(ctx=torch.device("cpu"))
        ckpt = {prefix + inv[k]: v.numpy() for k, v in p.items()}
        if prefix == "td.model.":
            for k, prm in model.rnn._reg_params.items():       # deferred input width: give the checkpoint a concrete one
                shape = tuple(s if s > 0 else 16 for s in prm.shape)
                ckpt["rnn." + k] = np.zeros(shape, dtype=np.float32)
        path = str(tmp_path / ("%s_%s.params" % (arch, prefix.strip("."))))
        params_io.save(path, ckpt)
        model.load_parameters(path, ctx=torch.device("cpu"))
        got = model.collect_params()
        for k in ("conv0.weight", ours[5], ours[-1]):
            assert torch.equal(got[prefix + k].data(), p[k]), (prefix, k)


_GLOO_MISMATCH_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from tennis_b200.gluon import Parameter, Trainer
from tennis_b200.parallel import ShardedCNNRNN
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# (1) rank 1 has not materialised its second parameter (deferred shape): Trainer.step must fail on EVERY rank instead of hanging
a, b = Parameter("a", shape=(3,)), Parameter("b", shape=(2,))
a._data = torch.zeros(3)
if rank == 0:
    b._data = torch.zeros(2)
tr = Trainer({"a": a, "b": b}, "sgd", {"learning_rate": 0.1})
try:
    tr.step(1)
    raise SystemExit("rank %%d: no error" %% rank)
except RuntimeError as e:
    assert "different parameter sets" in str(e), e
# (2) ShardedCNNRNN.forward is the even-shard form: unequal clip counts are refused on every rank
class _M(object):
    class td(object):
        @staticmethod
        def model(x):
            return x.reshape(x.shape[0], -1)[:, :4].float()
sh = ShardedCNNRNN(_M())
try:
    sh(torch.zeros(2 + rank, 3, 1, 2, 2))
    raise SystemExit("rank %%d: no error" %% rank)
except ValueError as e:
    assert "same number of clips" in str(e), e
dist.destroy_process_group()
print("ok", rank)
"""


def test_mismatched_ranks_fail_loudly_instead_of_hanging_gloo(tmp_path):
    """Two ways a multi-GPU run used to hang in a mismatched collective (profiles: 22 GPU-minutes lost to the first one): a rank
    whose deferred-shape parameters do not exist yet entering Trainer.step, and unequal clip shards through the even-shard forward."""
    script = tmp_path / "m.py"
    script.write_text(_GLOO_MISMATCH_WORKER % ROOT)
    env = dict(os.environ, WORLD_SIZE="2", MASTER_PORT="29617", MASTER_ADDR="127.0.0.1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def _emu_planes(src, transpose=False, pad_hw=None, pitch=None, shift=0):
    """fp64 emulation of tn_split_bf16's INDEXING (include/tennis_b200.h): rows (n,y,x) -> (n,y+1,x+1) of the zero-padded grid of
    (h+2) rows of `pitch` pixels, optional transpose, every pixel `shift` positions later."""
    R, C = src.shape
    if pad_hw is None:
        out = src.clone()
    else:
        h, w = pad_hw
        pitch = pitch or (w + 2)
        n = R // (h * w)
        out = torch.zeros(n * (h + 2) * pitch, C, dtype=src.dtype)
        r = torch.arange(R)
        x, y, f = r % w, (r // w) % h, r // (w * h)
        out[(f * (h + 2) + y + 1) * pitch + x + 1 + shift] = src
    return out.t().contiguous() if transpose else out


def _emu_gemm(A, B, M, N, K, taps, tile_taps=False):
    """fp64 emulation of tn_gemm_tc: D[m, n] = sum_t sum_{k<K} A[m + ar, ak + k] * B[n + br, bk + k], reads outside a matrix are
    zeros; tile_taps: one (M, N) product per tap."""
    def win(X, r0, nr, k0):
        out = torch.zeros(nr, K, dtype=X.dtype)
        rows = torch.arange(r0, r0 + nr)
        cols = torch.arange(k0, k0 + K)
        rv = (rows >= 0) & (rows < X.shape[0])
        cv = (cols >= 0) & (cols < X.shape[1])
        if rv.any() and cv.any():
            out[rv.nonzero()[:, 0][:, None], cv.nonzero()[:, 0][None, :]] = X[rows[rv]][:, cols[cv]]
        return out
    outs = [win(A, ar, M, ak) @ win(B, br, N, bk).t() for ar, ak, br, bk in taps]
    return outs if tile_taps else sum(outs)


def _unpad(D, n, h, w):
    return D.reshape(n, h + 2, w + 2, -1)[:, 1:h + 1, 1:w + 1].reshape(n * h * w, -1)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 5, 6, 8, 16), (1, 7, 7, 16, 8), (3, 4, 9, 8, 8)])
def test_tap_lists_reproduce_conv3x3(n, h, w, cin, cout):
    """The tap lists the training graph hands to tn_gemm_tc (tcgemm.taps_conv3x3_*), run through an fp64 emulation of the
    documented tn_split_bf16 / tn_gemm_tc index semantics, reproduce torch's 3x3 / stride 1 / pad 1 convolution, its data gradient
    and its weight gradient -- including the 8-pixel row pitch and the three dx-shifted copies of the weight-gradient operand."""
    import torch.nn.functional as F
    from tennis_b200 import tcgemm
    g = torch.Generator().manual_seed(n * 100 + h)
    x = torch.randn(n, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(n, cout, h, w, generator=g, dtype=torch.float64)
    y = F.conv2d(x, wt, padding=1)
    (y * dy).sum().backward()
    x2 = x.detach().permute(0, 2, 3, 1).reshape(-1, cin)          # NHWC rows
    dy2 = dy.permute(0, 2, 3, 1).reshape(-1, cout)
    wk = wt.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin)    # (Cout, (r,s,c))
    Mp = tcgemm.padded_rows(n, h, w)
    # forward
    Y = _emu_gemm(_emu_planes(x2, pad_hw=(h, w)), wk, Mp, cout, cin, tcgemm.taps_conv3x3_forward(w, cin))
    assert torch.allclose(_unpad(Y, n, h, w), y.detach().permute(0, 2, 3, 1).reshape(-1, cout), atol=1e-10)
    # data gradient
    wT = wt.detach().permute(1, 2, 3, 0).reshape(cin, 9 * cout)    # (Cin, (t, n))
    DX = _emu_gemm(_emu_planes(dy2, pad_hw=(h, w)), wT, Mp, cin, cout, tcgemm.taps_conv3x3_dgrad(w, cout))
    assert torch.allclose(_unpad(DX, n, h, w), x.grad.permute(0, 2, 3, 1).reshape(-1, cin), atol=1e-10)
    # weight gradient: contraction over the padded pixels at pitch P8, dx baked into three copies of dY^T
    P8 = tcgemm.pitch8(w)
    assert P8 % 8 == 0 and P8 >= w + 2
    Mp8 = tcgemm.padded_rows(n, h, w, P8)
    xT = _emu_planes(x2, transpose=True, pad_hw=(h, w), pitch=P8)
    dyT = torch.cat([_emu_planes(dy2, transpose=True, pad_hw=(h, w), pitch=P8, shift=dx) for dx in (-1, 0, 1)], 0)
    taps = tcgemm.taps_conv3x3_wgrad(w, cout)
    assert all(t[1] % 8 == 0 and t[3] % 8 == 0 for t in taps)      # TMA: 16-byte aligned contraction coordinates
    Dt = _emu_gemm(xT, dyT, cin, cout, Mp8, taps, tile_taps=True)   # nine (Cin, Cout) products
    dwk = torch.stack([d.t() for d in Dt], 1).reshape(cout, 9 * cin)  # dwk[n, t*Cin + c] = D_t[c, n]
    assert torch.allclose(dwk.reshape(cout, 3, 3, cin).permute(0, 3, 1, 2), wt.grad, atol=1e-9)


def test_num_gpus_relaunch_command(monkeypatch):
    """`python train.py --num_gpus 4 ...` (reference train.py:103: a context per GPU) re-executes the SAME command line under
    torch.distributed.run with one process per GPU on the loopback address; a process already under torchrun does not relaunch."""
    from tennis_b200 import cli
    seen = {}

    def fake_call(cmd):
        seen["cmd"] = cmd
        return 7
    monkeypatch.setattr("subprocess.call", fake_call)
    monkeypatch.setattr(sys, "argv", ["train.py", "--num_gpus", "4", "--synthetic"])
    assert cli._relaunch_under_torchrun(4) == 7       # the child's exit code is passed on
    cmd = seen["cmd"]
    assert cmd[0] == sys.executable and cmd[1:3] == ["-m", "torch.distributed.run"]
    assert "--nnodes=1" in cmd and cmd[cmd.index("--nproc-per-node") + 1] == "4"
    assert cmd[cmd.index("--master-addr") + 1] == "127.0.0.1" and int(cmd[cmd.index("--master-port") + 1]) > 0
    assert cmd[-4:] == ["train.py", "--num_gpus", "4", "--synthetic"]


def test_device_metric_state_round_trip():
    """The accumulator layout that metrics are summed in across ranks (cli.sync_metrics): state_tensor / load_state_tensor of the
    host metrics is lossless and additive -- two ranks that each saw a part of the samples sum to the single-process metric."""
    import numpy as np
    from tennis_b200.metrics.vision import PRF1, Accuracy
    names = ["OTH", "SFI", "SFF", "HFR", "HNR"]
    rng = np.random.RandomState(0)
    labels = rng.randint(0, 5, size=40)
    preds = rng.rand(40, 5)
    full_a, full_p = Accuracy(), PRF1(label_names=names)
    full_a.update([labels], [preds])
    full_p.update([labels], [preds])
    sa = sp = None
    for lo, hi in ((0, 13), (13, 40)):
        a, p = Accuracy(), PRF1(label_names=names)
        a.update([labels[lo:hi]], [preds[lo:hi]])
        p.update([labels[lo:hi]], [preds[lo:hi]])
        sa = a.state_tensor() if sa is None else sa + a.state_tensor()
        sp = p.state_tensor() if sp is None else sp + p.state_tensor()
    a2, p2 = Accuracy(), PRF1(label_names=names)
    a2.load_state_tensor(sa)
    p2.load_state_tensor(sp)
    assert a2.get() == full_a.get()
    assert p2.get() == full_p.get() and np.array_equal(p2.mat, full_p.mat)


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line on stdout with the contract keys, the CPU
    oracle port as `cpu_baseline.kind = port`, zero host<->device bytes, and the same metric / unit as the GPU arm."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"].startswith("frames/sec") and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
