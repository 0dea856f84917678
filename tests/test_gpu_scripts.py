"""Script-level smoke tests on the seeded synthetic dataset (the reference has no tests)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout + r.stderr


def test_evaluate_cnn_gru_synthetic(tmp_path):
    out = _run([os.path.join(ROOT, "evaluate.py"), "--backbone", "DenseNet121", "--temp_pool", "gru", "--window", "8",
                "--data_shape", "224", "--batch_size", "8", "--every", "16,16,16", "--synthetic", "--model_id", "t001"],
               str(tmp_path))
    assert "AVG_NB_f1" in out and "accuracy" in out


def test_evaluate_features_only_and_save_feats(tmp_path):
    out = _run([os.path.join(ROOT, "evaluate.py"), "--backbone", "resnet18_v2", "--data_shape", "224", "--batch_size", "8",
                "--every", "24,24,24", "--synthetic", "--save_feats", "--model_id", "t002"], str(tmp_path))
    feats = []
    for dp, _, fn in os.walk(os.path.join(str(tmp_path), "data", "features", "t002")):
        feats += [os.path.join(dp, f) for f in fn if f.endswith(".npy")]
    assert len(feats) == 8, out[-1500:]  # 2 videos x 96 frames / every 24
    import numpy as np
    assert np.load(feats[0]).shape == (512,)
    assert "/V000.mp4/0000000000/00000000" in feats[0].replace(os.sep, "/") or "/V001.mp4/" in feats[0].replace(os.sep, "/")


def test_evaluate_gnmt_synthetic(tmp_path):
    out = _run([os.path.join(ROOT, "evaluate_gnmt.py"), "--feats_model", "0006", "--cell_type", "lstm", "--beam_size", "5",
                "--test_batch_size", "4", "--tgt_max_len", "10", "--synthetic", "--model_id", "t101"], str(tmp_path))
    assert "tokens/sec" in out and "bleu=" in out
    assert os.path.exists(os.path.join(str(tmp_path), "models", "captioning", "experiments", "t101", "best_test_out.txt"))


def test_train_gnmt_synthetic_and_resume(tmp_path):
    """Captioner training on features (the published setting): epochs, checkpoints, valid_best, per-epoch outputs, resume."""
    args = [os.path.join(ROOT, "train_gnmt.py"), "--feats_model", "0006", "--cell_type", "gru", "--batch_size", "4",
            "--test_batch_size", "4", "--tgt_max_len", "10", "--epochs", "2", "--log_interval", "1", "--num_hidden", "32",
            "--synthetic", "--model_id", "t102"]
    out = _run(args, str(tmp_path))
    exp = os.path.join(str(tmp_path), "models", "captioning", "experiments", "t102")
    for f in ("0000.params", "0001.params", "valid_best.params", "epoch1_valid_out.txt", "best_test_out.txt", "val_gt.txt"):
        assert os.path.exists(os.path.join(exp, f)), f
    assert "[Epoch 1] valid Loss=" in out and "Best model test Loss=" in out and "Learning rate change" in out
    out2 = _run(args[:args.index("--epochs")] + ["--epochs", "3"] + args[args.index("--epochs") + 2:], str(tmp_path))
    assert "Loaded model params" in out2 and os.path.exists(os.path.join(exp, "0002.params"))
    # the trained checkpoint is what evaluate_gnmt.py picks up
    out3 = _run([os.path.join(ROOT, "evaluate_gnmt.py"), "--feats_model", "0006", "--cell_type", "gru", "--num_hidden", "32",
                 "--test_batch_size", "4", "--tgt_max_len", "10", "--synthetic", "--model_id", "t102"], str(tmp_path))
    assert "Loaded model params" in out3 and "valid_best.params" in out3


def test_train_head_on_features_synthetic_and_resume(tmp_path):
    """The published CNN-RNN setting (features -> BiGRU -> max -> Dense): two epochs, checkpoints, scores.txt, resume."""
    args = [os.path.join(ROOT, "train.py"), "--feats_model", "0006", "--temp_pool", "gru", "--window", "8", "--batch_size", "16",
            "--every", "4,8,8", "--epochs", "2", "--log_interval", "2", "--lr", "0.05", "--synthetic", "--model_id", "t042"]
    out = _run(args, str(tmp_path))
    exp = os.path.join(str(tmp_path), "models", "vision", "experiments", "t042")
    assert os.path.exists(os.path.join(exp, "0000.params")) and os.path.exists(os.path.join(exp, "0001.params"))
    assert len(open(os.path.join(exp, "scores.txt")).read().split()) == 4
    assert "test AVG_NB_f1" in out
    out2 = _run(args[:args.index("--epochs")] + ["--epochs", "3"] + args[args.index("--epochs") + 2:], str(tmp_path))
    assert "Loaded model params" in out2 and os.path.exists(os.path.join(exp, "0002.params"))


def test_train_cnn_framewise_trainable_backbone(tmp_path):
    """train.py's default flow with a TRAINABLE backbone: FrameModel(resnet18_v2) on single frames, one epoch."""
    out = _run([os.path.join(ROOT, "train.py"), "--backbone", "resnet18_v2", "--data_shape", "224", "--batch_size", "4",
                "--every", "24,48,48", "--epochs", "1", "--log_interval", "1", "--synthetic", "--model_id", "t044"], str(tmp_path))
    assert "validation" in out
    assert os.path.exists(os.path.join(str(tmp_path), "models", "vision", "experiments", "t044", "0000.params"))


def test_train_gnmt_trainable_cnn_source(tmp_path):
    """train_gnmt.py:150-170 of the reference: frames -> TimeDistributed(trainable CNN) -> GNMT; the encoder's source-feature
    gradient trains the backbone (one epoch on the synthetic stand-in)."""
    out = _run([os.path.join(ROOT, "train_gnmt.py"), "--backbone", "resnet18_v2", "--data_shape", "224", "--cell_type", "gru",
                "--batch_size", "2", "--test_batch_size", "2", "--tgt_max_len", "10", "--every", "8", "--epochs", "1", "--log_interval",
                "1", "--num_hidden", "32", "--synthetic", "--model_id", "t103"], str(tmp_path))
    assert "Training the CNN through the captioner" in out and "[Epoch 0] valid Loss=" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_scripts_with_num_gpus_2(tmp_path):
    """--num_gpus 2 (reference train.py:103): the script re-launches itself with one process per GPU, shards every batch, sums the
    gradients over NCCL and lets rank 0 write the artefacts exactly once."""
    args = [os.path.join(ROOT, "train.py"), "--feats_model", "0006", "--temp_pool", "gru", "--window", "8", "--batch_size", "16",
            "--every", "4,8,8", "--epochs", "2", "--log_interval", "2", "--lr", "0.05", "--synthetic", "--model_id", "t045",
            "--num_gpus", "2"]
    out = _run(args, str(tmp_path))
    exp = os.path.join(str(tmp_path), "models", "vision", "experiments", "t045")
    assert os.path.exists(os.path.join(exp, "0001.params")) and "test AVG_NB_f1" in out
    assert len(open(os.path.join(exp, "scores.txt")).read().split()) == 4   # written once, by rank 0
    out = _run([os.path.join(ROOT, "evaluate.py"), "--backbone", "resnet18_v2", "--temp_pool", "gru", "--window", "4", "--data_shape",
                "224", "--batch_size", "6", "--every", "24,48,48", "--synthetic", "--model_id", "t046", "--num_gpus", "2"], str(tmp_path))
    assert "AVG_NB_f1" in out
    out = _run([os.path.join(ROOT, "train_gnmt.py"), "--feats_model", "0006", "--cell_type", "gru", "--batch_size", "4",
                "--test_batch_size", "4", "--tgt_max_len", "10", "--epochs", "1", "--log_interval", "1", "--num_hidden", "32",
                "--synthetic", "--model_id", "t104", "--num_gpus", "2"], str(tmp_path))
    assert "[Epoch 0] valid Loss=" in out


def test_train_frozen_backbone_cnn_gru(tmp_path):
    out = _run([os.path.join(ROOT, "train.py"), "--backbone", "resnet18_v2", "--freeze_backbone", "--temp_pool", "gru", "--window", "4",
                "--data_shape", "224", "--batch_size", "4", "--every", "24,48,48", "--epochs", "1", "--synthetic", "--model_id",
                "t043"], str(tmp_path))
    assert "validation" in out


@pytest.mark.parametrize("env_extra", [{"TN_TILE_PAIR_MIN": "1"}, {"TN_NO_PDL": "1", "TN_TILE_PAIR_MIN": "1000000"},
                                       {"TN_RNN_NO_WREG": "1"},
                                       {"TN_CHUNK_B1": "3", "TN_CHUNK_B2": "5", "TN_CHUNK_B3": "3", "TN_CHUNK_B4": "7"},
                                       {"TN_L2_PREFETCH": "4", "TN_STAGE_CAP": "2"}],
                         ids=["paired_tiles_forced", "no_pdl_no_pairing", "rnn_smem_weights", "frame_chunked_blocks",
                              "l2_prefetch_short_ring"])
def test_smoke_parity_under_kernel_variants(env_extra):
    """Every launch-path switch must give the same oracle parity: DenseNet-121 + Bi-GRU logits vs the CPU oracle with
    tile pairing forced on for every streamed-weight 1x1 conv, with PDL and pairing off, with the shared-memory RNN scan, with
    the dense blocks run in ragged frame chunks, and with TMA L2 prefetch on a two-stage ring (the clamp prologue has its own
    parity tests in test_gpu_backbone.py)."""
    env = dict(os.environ, PYTHONPATH=ROOT, **env_extra)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "argmax equal: True" in r.stdout
