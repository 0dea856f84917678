"""Event-detector evaluation (same CLI surface as the reference's evaluate.py).

    python evaluate.py --backbone DenseNet121 --temp_pool gru --window 32 --data_shape 224 --synthetic
"""
import logging
import os
import sys
import time

import numpy as np
import torch
from absl import app, flags

from tennis_b200 import cli
from tennis_b200.dataset import TennisSet
from tennis_b200.metrics.vision import PRF1, Accuracy
from tennis_b200.models.vision.definitions import TemporalPooling

cli.define_detector_flags(training=False)
FLAGS = flags.FLAGS


def batches(dataset, batch_size):
    """Default batchify over (img, label, idx) samples, shuffle=False (evaluate.py:112-114)."""
    for lo in range(0, len(dataset), batch_size):
        mine = cli.rank_shard(list(range(lo, min(len(dataset), lo + batch_size))))  # --num_gpus N: this process's part
        if not mine:
            continue
        items = [dataset[i] for i in mine]
        yield (torch.stack([it[0] for it in items]), torch.tensor([it[1] for it in items]),
               torch.tensor([it[2] for it in items]))


def evaluate_model(net, dataset, ctx, metrics, batch_size):
    """reference evaluate.py:274-303 (one device per process; the per-sample bookkeeping bug of :287 is not copied)."""
    results, gts = {}, {}
    tic = time.time()
    n = 0
    for data, labels, idxs in batches(dataset, batch_size):
        out = net(data.to(ctx, non_blocking=True)).cpu()  # one D2H copy per batch (Appendix C #2)
        for m in metrics:
            m.update([labels], [out])
        for j, i in enumerate(idxs.tolist()):
            s = dataset._samples[i]
            key = '%s/%010d' % (s[0], s[1])
            results[key] = out[j].numpy()
            gts[key] = int(labels[j])
        n += data.shape[0]
    logging.info('evaluated %d samples in %.2fs (%.1f samples/sec)', n, time.time() - tic, n / max(1e-9, time.time() - tic))
    return results, gts


def save_features(net, dataset, ctx, batch_size):
    """reference evaluate.py:306-321: net.backbone(x) -> one .npy per frame under data/features/<model_id>/..."""
    for data, _, idxs in batches(dataset, batch_size):
        feat = net.backbone(data.to(ctx)).cpu().numpy()
        for j, i in enumerate(idxs.tolist()):
            path = dataset.save_feature_path(i)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.save(path, feat[j])


def main(_argv):
    cli.parse_list_flags()
    ctx = cli.context()
    exp_dir = os.path.join('models', 'vision', 'experiments', FLAGS.model_id)
    cli.setup_logging(exp_dir)
    split_i = {'train': 0, 'val': 1, 'test': 2}[FLAGS.split]
    test_tf = None
    if FLAGS.feats_model is None:
        from tennis_b200 import transforms  # real frames: Resize(S+32) -> CenterCrop(S) on the host, normalisation on the GPU
        test_tf = transforms.TestTransform(FLAGS.data_shape)
    dataset = TennisSet(split=FLAGS.split, transform=test_tf, balance=FLAGS.balance[split_i], every=FLAGS.every[split_i], padding=FLAGS.padding,
                        stride=FLAGS.stride, window=FLAGS.window, model_id=FLAGS.model_id, split_id=FLAGS.split_id,
                        feats_model=FLAGS.feats_model, save_feats=FLAGS.save_feats, data_shape=FLAGS.data_shape,
                        synthetic={} if FLAGS.synthetic else None)
    logging.info('%s set: %d samples', FLAGS.split, len(dataset))
    model = cli.build_detector(ctx, len(dataset.classes))
    path, _ = cli.latest_params(exp_dir)
    best = cli.best_epoch(exp_dir)
    if best is not None and os.path.exists(os.path.join(exp_dir, '%04d.params' % best)):
        path = os.path.join(exp_dir, '%04d.params' % best)
    if path is not None:
        model.load_parameters(path, ctx=ctx)
        logging.info('Loaded model params: %s', path)
    else:
        logging.warning('no checkpoint under %s: evaluating freshly initialised weights', exp_dir)
    if FLAGS.save_feats:
        save_features(model, dataset, ctx, FLAGS.batch_size)
        cli.shutdown()
        return
    if FLAGS.temp_pool in ('max', 'mean') and FLAGS.window > 1 and FLAGS.feats_model is None:
        model = TemporalPooling(model, pool=FLAGS.temp_pool, num_classes=0, feats=False)
    metrics = [Accuracy(), Accuracy('top5', top_k=5), PRF1(label_names=dataset.classes)]
    evaluate_model(model, dataset, ctx, metrics, FLAGS.batch_size)
    cli.sync_metrics(metrics)
    if cli.is_main():
        print(metrics[2].mat.astype(int))
        for m in metrics[:2]:
            print('%s: %.4f' % m.get())
        for k, v in metrics[2].get():
            print('%s: %.4f' % (k, v))
    cli.shutdown()


if __name__ == '__main__':
    try:
        app.run(main)
    except SystemExit:
        raise
