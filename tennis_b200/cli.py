"""Shared pieces of the four scripts: flag surfaces (names and defaults verbatim from the reference's train.py:32-93,
evaluate.py:30-75, train_gnmt.py:48-119, evaluate_gnmt.py:42-88), model assembly (train.py:195-241) and the
checkpoint directory conventions (train.py:286-295, 334-346)."""
import logging
import os

import torch
from absl import flags

from . import model_zoo
from .gluon import Dropout, Embedding, HybridSequential, Uniform
from .models.vision.definitions import CNNRNN, FrameModel, TemporalPooling

FLAGS = flags.FLAGS


def define_detector_flags(training):
    flags.DEFINE_string('backbone', 'resnet18_v2', 'Backbone CNN name: resnet18_v2, DenseNet121')
    flags.DEFINE_string('backbone_from_id', None, 'Load a backbone model from a model_id, used for Temporal Pooling with fine-tuned CNN')
    flags.DEFINE_bool('freeze_backbone', False, 'Freeze the backbone model')
    flags.DEFINE_string('model_id', '0000', 'model identification string')
    flags.DEFINE_string('split_id', '02', 'split identification string, 01: single test vid; 02: all videos have test sections')
    flags.DEFINE_integer('data_shape', 512, 'The width and height for the input image to be cropped to.')
    flags.DEFINE_list('every', '1, 1, 1', 'Use only every this many frames: [train, val, test] splits')
    flags.DEFINE_list('balance', 'True, False, False', 'Balance the play/not class samples: [train, val, test] splits')
    flags.DEFINE_integer('window', 1, 'Temporal window size of frames')
    flags.DEFINE_integer('padding', 1, 'Frame*every + and - padding around the marked event boundaries: [train, val, test] splits')
    flags.DEFINE_integer('stride', 1, 'Temporal stride of samples within a window')
    flags.DEFINE_integer('batch_size', 64, 'Batch size for detection: higher faster, but more memory intensive.')
    flags.DEFINE_integer('num_gpus', 1, 'Number of GPUs to use')
    flags.DEFINE_integer('num_workers', -1, 'The number of workers should be picked so that it is equal to number of cores on your machine for max parallelization. If this number is bigger than your number of cores it will use up a bunch of extra CPU memory. -1 is auto.')
    flags.DEFINE_bool('vis', False, 'Visualise testing results')
    flags.DEFINE_bool('save_feats', False, 'save CNN features as npy files')
    flags.DEFINE_string('feats_model', None, 'load CNN features as npy files from this model')
    flags.DEFINE_string('flow', '', 'How to use flow, "" for none, "only" for no rgb, "sixc" for six channel inp, "twos" for twostream')
    flags.DEFINE_string('temp_pool', None, 'mean, max or gru.')
    if training:
        flags.DEFINE_integer('log_interval', 100, 'Logging mini-batch interval.')
        flags.DEFINE_integer('epochs', 20, 'How many training epochs to complete')
        flags.DEFINE_float('lr', 0.001, 'Learning rate')
        flags.DEFINE_float('lr_factor', 0.75, 'lr factor')
        flags.DEFINE_list('lr_steps', '10, 20', 'Epochs at which learning rate factor applied')
        flags.DEFINE_float('momentum', 0.9, 'momentum')
        flags.DEFINE_float('wd', 0.0001, 'weight decay')
        flags.DEFINE_integer('max_batches', -1, 'for 0031 this number of batches per epoch')
    else:
        flags.DEFINE_string('split', 'test', 'the split to evaluate on: train, val, test')
    flags.DEFINE_bool('synthetic', False, 'use the seeded synthetic stand-in dataset (no data/ directory needed)')


def define_captioner_flags(training):
    flags.DEFINE_string('model_id', '0000', 'model identification string')
    flags.DEFINE_integer('epochs', 40, 'How many training epochs to complete')
    flags.DEFINE_integer('num_hidden', 128, 'Dimension of the states')
    flags.DEFINE_integer('emb_size', 100, 'Dimension of the embedding vectors')
    flags.DEFINE_float('dropout', 0.2, 'dropout applied to layers (0 = no dropout)')
    flags.DEFINE_integer('num_layers', 2, 'Number of layers in the encoder  and decoder')
    flags.DEFINE_integer('num_bi_layers', 1, 'Number of bidirectional layers in the encoder and decoder')
    flags.DEFINE_string('cell_type', 'gru', 'gru or lstm')
    flags.DEFINE_integer('batch_size', 128, 'Batch size for detection: higher faster, but more memory intensive.')
    flags.DEFINE_integer('beam_size', 4, 'Beam size.')
    flags.DEFINE_float('lp_alpha', 1.0, 'Alpha used in calculating the length penalty')
    flags.DEFINE_integer('lp_k', 5, 'K used in calculating the length penalty')
    flags.DEFINE_integer('test_batch_size', 32, 'Test batch size')
    flags.DEFINE_integer('num_buckets', 5, 'Bucket number')
    flags.DEFINE_string('bucket_scheme', 'constant', 'Strategy for generating bucket keys. It supports: "constant": all the buckets have the same width')
    flags.DEFINE_float('bucket_ratio', 0.0, 'Ratio for increasing the throughput of the bucketing')
    flags.DEFINE_integer('tgt_max_len', 50, 'Maximum length of the target sentence')
    flags.DEFINE_string('optimizer', 'adam', 'optimization algorithm')
    flags.DEFINE_float('lr', 1E-3, 'Initial learning rate')
    flags.DEFINE_float('lr_update_factor', 0.5, 'Learning rate decay factor')
    flags.DEFINE_float('clip', 5.0, 'gradient clipping (defined by the reference, never applied: train_gnmt.py:97-98)')
    flags.DEFINE_integer('log_interval', 100, 'Logging mini-batch interval.')
    flags.DEFINE_integer('num_gpus', 1, 'Number of GPUs to use')
    flags.DEFINE_string('backbone', 'DenseNet121', 'Backbone CNN name')
    flags.DEFINE_string('backbone_from_id', None, 'Load a backbone model from a model_id, used for Temporal Pooling with fine-tuned CNN')
    flags.DEFINE_bool('freeze_backbone', False, 'Freeze the backbone model')
    flags.DEFINE_integer('data_shape', 512, 'The width and height for the input image to be cropped to.')
    flags.DEFINE_integer('every', 1, 'Use only every this many frames: [train, val, test] splits')
    flags.DEFINE_string('feats_model', None, 'load CNN features as npy files from this model')
    flags.DEFINE_string('emb_file', 'embeddings-ex.txt', 'the word embedding file generated by train_embeddings.py')
    flags.DEFINE_bool('synthetic', False, 'use the seeded synthetic stand-in dataset (no data/ directory needed)')


def parse_list_flags():
    """train.py:97-99."""
    FLAGS.every = [int(s) for s in FLAGS.every]
    FLAGS.balance = [True if str(s).lower().strip() in ('true', 't', 'yes', 'y', '1') else False for s in FLAGS.balance]
    if 'lr_steps' in FLAGS:
        FLAGS.lr_steps = [int(s) for s in FLAGS.lr_steps]


def world():
    """(rank, world_size) of this process; (0, 1) outside torch.distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def is_main():
    return world()[0] == 0


def _relaunch_under_torchrun(n):
    """`python train.py --num_gpus N` (reference train.py:103: ctx = [mx.gpu(i) for i in range(num_gpus)]) -> one process per GPU:
    re-exec the same command line under torch.distributed.run and return its exit code."""
    import socket
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + sys.argv
    logging.info("--num_gpus %d: relaunching as %s", n, " ".join(cmd))
    return subprocess.call(cmd)


def context():
    """Device of this process.  --num_gpus N > 1 means N processes, one per GPU (frames / clips sharded by rank, gradients summed
    over NCCL in Trainer.step): a plain invocation relaunches itself under torch.distributed.run; under torchrun the process group
    is initialised here.  Only rank 0 writes checkpoints, scores.txt and log.txt (see is_main())."""
    import torch.distributed as dist
    if FLAGS.num_gpus <= 0:
        raise SystemExit("--num_gpus 0 selects the reference's MXNet CPU path; tennis_b200 is GPU-only (no CPU fallback)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device visible: tennis_b200 has no CPU fallback")
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if FLAGS.num_gpus > 1 and "RANK" not in os.environ:
        if torch.cuda.device_count() < FLAGS.num_gpus:
            raise SystemExit("--num_gpus %d but only %d CUDA devices are visible" % (FLAGS.num_gpus, torch.cuda.device_count()))
        raise SystemExit(_relaunch_under_torchrun(FLAGS.num_gpus))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if ws > 1 and not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return torch.device("cuda", local)


def setup_logging(exp_dir):
    os.makedirs(exp_dir, exist_ok=True)
    logger = logging.getLogger()
    logger.setLevel(logging.INFO if is_main() else logging.WARNING)
    if is_main():  # one log.txt per experiment, not one per rank
        logger.addHandler(logging.FileHandler(os.path.join(exp_dir, 'log.txt')))
    logging.basicConfig()


def barrier():
    import torch.distributed as dist
    if world()[1] > 1:
        dist.barrier()


def shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def broadcast_parameters(model):
    """Every rank continues from rank 0's parameters (initialisation is seeded per process; checkpoints are read by all)."""
    import torch.distributed as dist
    if world()[1] == 1:
        return
    for _, p in sorted(model.collect_params().items()):
        if p._data is not None:
            dist.broadcast(p._data, src=0)
            p._bump()


def rank_shard(items):
    """This rank's contiguous part of one global batch (list or tensor along dim 0): gluon.utils.split_and_load(batch, ctx_list,
    even_split=False) semantics (train.py:410-412) with one context per process."""
    from .parallel import shard_range
    r, w = world()
    if w == 1:
        return items
    lo, hi = shard_range(len(items), r, w)
    return items[lo:hi]


def sync_metrics(metrics):
    """Sum the metric accumulators over the ranks (each rank evaluated its shard of every batch)."""
    import torch.distributed as dist
    r, w = world()
    if w == 1:
        return metrics
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    for m in metrics:
        state = m.state_tensor().to(dev)
        dist.all_reduce(state)
        m.load_state_tensor(state.cpu())
    return metrics


def latest_params(exp_dir):
    """Newest NNNN.params (lexicographic), skipping valid_best.params (train.py:286-295, train_gnmt.py:236-247)."""
    if not os.path.isdir(exp_dir):
        return None, 0
    files = sorted([f for f in os.listdir(exp_dir) if f.endswith('.params') and f != 'valid_best.params'], reverse=True)
    if not files:
        return None, 0
    return os.path.join(exp_dir, files[0]), int(files[0].split('.')[0]) + 1


def best_epoch(exp_dir):
    """Best epoch by AVG_NB_f1 from scores.txt (train.py:334-346)."""
    path = os.path.join(exp_dir, 'scores.txt')
    if not os.path.exists(path):
        return None
    best, best_score = None, -1.0
    for line in open(path):
        parts = line.split()
        if len(parts) >= 2 and float(parts[1]) > best_score:
            best, best_score = int(parts[0]), float(parts[1])
    return best


def build_detector(ctx, num_classes=11):
    """Model assembly of train.py:195-241 / evaluate.py:116-162."""
    if FLAGS.feats_model is None:
        backbone = model_zoo.get_model(FLAGS.backbone, pretrained=False, ctx=ctx).features
        model = FrameModel(backbone, num_classes)
        if FLAGS.backbone_from_id:
            path, _ = latest_params(os.path.join('models', 'vision', 'experiments', FLAGS.backbone_from_id))
            if path is None:
                raise FileNotFoundError(os.path.join('models', 'vision', 'experiments', FLAGS.backbone_from_id))
            model.load_parameters(path, ctx=ctx, allow_missing=True)
            logging.info('Loaded backbone params: %s', path)
        if FLAGS.freeze_backbone:
            for p in model.collect_params().values():
                p.grad_req = 'null'
        if FLAGS.temp_pool in ('gru', 'lstm'):
            model = CNNRNN(model, num_classes, type=FLAGS.temp_pool, hidden_size=128)
        elif FLAGS.temp_pool in ('mean', 'max'):
            pass  # trained frame-wise, pooled at test time (train.py:349-351, evaluate.py:242-244)
    else:
        if FLAGS.temp_pool in ('gru', 'lstm'):
            model = CNNRNN(None, num_classes, type=FLAGS.temp_pool, hidden_size=128)
        else:
            model = TemporalPooling(None, num_classes, pool=FLAGS.temp_pool or 'max', feats=True) if FLAGS.window > 1 \
                else FrameModel(HybridSequential(), num_classes)
    model.initialize(init=Uniform(0.07), ctx=ctx)
    model.collect_params().reset_ctx(ctx)
    model.hybridize()
    return model


def build_captioner(ctx, vocab, embedding=None):
    """Model assembly of train_gnmt.py:148-233."""
    from .models.captioning.gnmt import NMTModel, get_gnmt_encoder_decoder
    from .utils.layers import TimeDistributed
    if FLAGS.feats_model is None:
        backbone = model_zoo.get_model(FLAGS.backbone, pretrained=False, ctx=ctx).features
        cnn = FrameModel(backbone, 11)
        if FLAGS.freeze_backbone:
            for p in cnn.collect_params().values():
                p.grad_req = 'null'
        src_embed = TimeDistributed(cnn.backbone)
    else:
        src_embed = HybridSequential(prefix='src_embed_')
        src_embed.add(Dropout(rate=0.0))
    tgt_embed = None
    if embedding is not None:
        tgt_embed = Embedding(embedding.shape[0], embedding.shape[1])
        tgt_embed.initialize(ctx=ctx)
        tgt_embed.weight.set_data(torch.as_tensor(embedding))
    encoder, decoder = get_gnmt_encoder_decoder(cell_type=FLAGS.cell_type, hidden_size=FLAGS.num_hidden,
                                                dropout=FLAGS.dropout, num_layers=FLAGS.num_layers,
                                                num_bi_layers=FLAGS.num_bi_layers)
    model = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=encoder, decoder=decoder, embed_size=FLAGS.emb_size,
                     prefix='gnmt_', src_embed=src_embed, tgt_embed=tgt_embed)
    model.initialize(init=Uniform(0.1), ctx=ctx)
    model.hybridize(static_alloc=True)
    return model
