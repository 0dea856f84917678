"""Event-detector training (same CLI surface as the reference's train.py).

Heads on pre-extracted features (`--feats_model <id>`, the published CNN-RNN 0042 setting) or on a frozen backbone
(`--freeze_backbone`) train on the fast path (tensor-core forward, fused BPTT).  A trainable backbone runs the fp32 training
graph of the CNN (models/vision/train_graph.py: batch-statistics BatchNorm, im2col convolutions; correct, not yet fast).

    python train.py --feats_model 0006 --temp_pool gru --window 30 --synthetic --epochs 2
    python train.py --backbone DenseNet121 --freeze_backbone --temp_pool gru --window 8 --data_shape 224 --synthetic
    python train.py --backbone resnet18_v2 --data_shape 224 --batch_size 8 --synthetic --epochs 1        # CNN trained frame-wise
"""
import logging
import os
import time

import numpy as np
import torch
from absl import app, flags

from tennis_b200 import autograd as ag
from tennis_b200 import cli
from tennis_b200.dataset import TennisSet
from tennis_b200.gluon import SoftmaxCrossEntropyLoss, Trainer
from tennis_b200.metrics.device import DeviceMetrics
from tennis_b200.metrics.vision import PRF1, Accuracy
from tennis_b200.models.vision.definitions import TemporalPooling

cli.define_detector_flags(training=True)
FLAGS = flags.FLAGS


def batches(dataset, batch_size, shuffle, seed):
    """Global batches in a rank-independent order; with --num_gpus N each process loads only ITS contiguous shard of every
    batch (reference train.py:410-412 split_and_load over the context list -> one context per process here).  A rank whose shard
    is empty (last, short batch) yields None and still takes part in the gradient all-reduce of Trainer.step."""
    order = list(range(len(dataset)))
    if shuffle:
        np.random.RandomState(seed).shuffle(order)
    for lo in range(0, len(order), batch_size):
        mine = cli.rank_shard(order[lo:lo + batch_size])
        if not mine:
            yield None
            continue
        items = [dataset[i] for i in mine]
        yield (torch.stack([it[0] for it in items]), torch.tensor([it[1] for it in items]),
               torch.tensor([it[2] for it in items]))


def test_model(net, dataset, ctx, metrics, batch_size):
    """reference train.py:503-527.  The metric counters are accumulated on the device (tennis_b200/metrics/device.py): no logits
    leave the GPU and nothing synchronises per batch; `metrics` = [Accuracy, Accuracy(top5), PRF1] is filled at the end."""
    dm = DeviceMetrics(dataset.classes, top_k=5, device=ctx)
    for batch in batches(dataset, batch_size, False, 0):
        if batch is None:
            continue
        data, labels, _ = batch
        dm.update(labels.to(ctx, non_blocking=True), net(data.to(ctx, non_blocking=True)))
    got = dm.finish()  # one D2H read; summed over ranks
    for m, g in zip(metrics, got):
        m.__dict__.update(g.__dict__)
    return metrics


def save_features(net, dataset, ctx, batch_size):
    """reference train.py:530-545: backbone features of every sample, one .npy per frame (existing files are kept)."""
    n = 0
    for batch in batches(dataset, batch_size, False, 0):
        if batch is None:
            continue
        data, _, idxs = batch
        feat = net.backbone(data.to(ctx)).cpu().numpy()
        for j, i in enumerate(idxs.tolist()):
            path = dataset.save_feature_path(i)
            if not os.path.exists(path):
                os.makedirs(os.path.dirname(path), exist_ok=True)
                np.save(path, feat[j])
                n += 1
    return n


def train_model(model, train_set, val_set, trainer, loss_fn, ctx, exp_dir, start_epoch):
    """reference train.py:388-500: epoch loop, LR steps, per-epoch validation, scores.txt, NNNN.params."""
    train_metrics = [Accuracy(), PRF1(label_names=train_set.classes)]
    val_metrics = [Accuracy(), Accuracy('top5', top_k=5), PRF1(label_names=val_set.classes)]
    lr_steps = sorted(FLAGS.lr_steps)
    lr_counter = sum(1 for s in lr_steps if s <= start_epoch)  # derived from the resume point (Appendix C #9)
    trainer.set_learning_rate(FLAGS.lr * (FLAGS.lr_factor ** lr_counter))
    for epoch in range(start_epoch, FLAGS.epochs):
        if lr_counter < len(lr_steps) and epoch == lr_steps[lr_counter]:
            trainer.set_learning_rate(trainer.learning_rate * FLAGS.lr_factor)
            lr_counter += 1
        for m in train_metrics:
            m.reset()
        tic, btic, train_loss, nb = time.time(), time.time(), 0.0, 0
        for i, batch in enumerate(batches(train_set, FLAGS.batch_size, True, epoch)):
            if FLAGS.max_batches > 0 and i >= FLAGS.max_batches:
                break
            if batch is not None:
                data, labels, _ = batch
                data, dl = data.to(ctx, non_blocking=True), labels.to(ctx)
                with ag.record():
                    out = model(data)
                    loss = loss_fn(out, dl)
                ag.backward([loss])
            trainer.step(FLAGS.batch_size)  # sums the per-rank gradients (NCCL), rescales by 1/global batch (train.py:424)
            if batch is None:
                continue
            nb += 1
            train_loss += loss.mean().item()  # device -> host sync point, as in the reference (train.py:427)
            for m in train_metrics:
                m.update([labels], [out.cpu()])
            if FLAGS.log_interval and (i + 1) % FLAGS.log_interval == 0:
                name, acc = train_metrics[0].get()
                logging.info('[Epoch %d] [Batch %d] Speed: %.3f samples/sec, %s=%.4f, lr=%.6f', epoch, i + 1,
                             FLAGS.batch_size * FLAGS.log_interval / max(1e-9, time.time() - btic), name, acc, trainer.learning_rate)
                btic = time.time()
        cli.sync_metrics(train_metrics)
        logging.info('[Epoch %d] training: loss=%.4f %s=%.4f time: %.1fs', epoch, train_loss / max(1, nb), *train_metrics[0].get(),
                     time.time() - tic)
        tic = time.time()
        test_model(model, val_set, ctx, val_metrics, FLAGS.batch_size)
        scores = dict(val_metrics[2].get())
        logging.info('[Epoch %d] validation: acc=%.4f AVG_NB_f1=%.4f time: %.1fs', epoch, val_metrics[0].get()[1], scores['AVG_NB_f1'],
                     time.time() - tic)
        if cli.is_main():  # one writer: ranks hold identical parameters after the summed-gradient update
            with open(os.path.join(exp_dir, 'scores.txt'), 'a') as f:
                f.write('%04d %.4f\n' % (epoch, scores['AVG_NB_f1']))
            model.save_parameters(os.path.join(exp_dir, '%04d.params' % epoch))
        cli.barrier()


def main(_argv):
    cli.parse_list_flags()
    ctx = cli.context()
    exp_dir = os.path.join('models', 'vision', 'experiments', FLAGS.model_id)
    cli.setup_logging(exp_dir)
    syn = {} if FLAGS.synthetic else None
    common = dict(padding=FLAGS.padding, stride=FLAGS.stride, window=FLAGS.window, model_id=FLAGS.model_id, split_id=FLAGS.split_id,
                  feats_model=FLAGS.feats_model, data_shape=FLAGS.data_shape, synthetic=syn, save_feats=FLAGS.save_feats)
    if FLAGS.save_feats:
        FLAGS.balance = [False, False, False]  # every frame gets a feature (train.py:159-160)
    train_tf = test_tf = None
    if FLAGS.feats_model is None:
        # decoded JPEGs (real data): Resize/CenterCrop/augmentation on the host (train.py:118-164); a trainable backbone gets
        # normalised fp32 frames for its fp32 training graph, everything else uint8 frames normalised on the GPU
        from tennis_b200 import transforms
        train_tf, test_tf = transforms.build_transforms(FLAGS.data_shape, FLAGS.window, FLAGS.save_feats,
                                                        to_tensor=not FLAGS.freeze_backbone)
    train_set = TennisSet(split='train', balance=FLAGS.balance[0], every=FLAGS.every[0], transform=train_tf, **common)
    val_set = TennisSet(split='val', balance=FLAGS.balance[1], every=FLAGS.every[1], transform=test_tf, **common)
    test_set = TennisSet(split='test', balance=FLAGS.balance[2], every=FLAGS.every[2], transform=test_tf, **common)
    logging.info('train/val/test: %d / %d / %d samples', len(train_set), len(val_set), len(test_set))
    model = cli.build_detector(ctx, len(train_set.classes))
    if FLAGS.save_feats:
        # train.py:266-284: load the best epoch by validation score and dump the backbone features of all three splits
        best = cli.best_epoch(exp_dir)
        if best is None:
            raise SystemExit("--save_feats needs a trained model: %s has no scores.txt" % exp_dir)
        model(train_set[0][0].unsqueeze(0).to(ctx))  # resolve deferred shapes before loading
        model.load_parameters(os.path.join(exp_dir, '%04d.params' % best), ctx=ctx)
        logging.info('Loaded model params: %s', os.path.join(exp_dir, '%04d.params' % best))
        for sett in (train_set, val_set, test_set):
            logging.info('saved %d new feature files for the %s split', save_features(model, sett, ctx, FLAGS.batch_size), sett._split)
        return
    path, start_epoch = cli.latest_params(exp_dir)
    if path is not None:
        x0 = train_set[0][0].unsqueeze(0).to(ctx)
        model(x0)  # resolve deferred shapes before loading
        model.load_parameters(path, ctx=ctx)
        logging.info('Loaded model params: %s', path)
    cli.broadcast_parameters(model)  # ranks start from rank 0's initialisation / checkpoint
    trainer = Trainer(model.collect_params(), 'sgd', {'learning_rate': FLAGS.lr, 'momentum': FLAGS.momentum, 'wd': FLAGS.wd})
    pooled_at_test = FLAGS.temp_pool in ('max', 'mean') and FLAGS.window > 1 and FLAGS.feats_model is None
    if pooled_at_test:
        # reference train.py:326,349-351: a max/mean-pooled model over a CNN is NOT trained here -- the frame-wise CNN comes from
        # --backbone_from_id and is only wrapped in TemporalPooling for the test below (its windows are 5-D clips the frame model
        # cannot train on)
        logging.info('--temp_pool %s --window %d without --feats_model: no training, pooled evaluation only (train.py:326)',
                     FLAGS.temp_pool, FLAGS.window)
    else:
        train_model(model, train_set, val_set, trainer, SoftmaxCrossEntropyLoss(), ctx, exp_dir, start_epoch)
    best = cli.best_epoch(exp_dir)
    if best is not None:
        model.load_parameters(os.path.join(exp_dir, '%04d.params' % best), ctx=ctx)
        logging.info('Testing best epoch %d', best)
    net = model
    if pooled_at_test:
        net = TemporalPooling(model, pool=FLAGS.temp_pool, num_classes=0, feats=False)  # train.py:349-351
    metrics = test_model(net, test_set, ctx, [Accuracy(), Accuracy('top5', top_k=5), PRF1(label_names=test_set.classes)],
                         FLAGS.batch_size)
    if cli.is_main():
        print(metrics[2].mat.astype(int))
        print('test %s: %.4f' % metrics[0].get())
        for k, v in metrics[2].get()[-6:]:
            print('test %s: %.4f' % (k, v))
    cli.shutdown()


if __name__ == '__main__':
    app.run(main)
