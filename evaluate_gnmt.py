"""Captioner evaluation (same CLI surface as the reference's evaluate_gnmt.py; its two latent bugs -- undefined
transform_train, evaluate() arity -- are fixed).

    python evaluate_gnmt.py --feats_model 0006 --cell_type lstm --beam_size 5 --synthetic
"""
import logging
import os
import time

import numpy as np
import torch
from absl import app, flags

from tennis_b200 import cli
from tennis_b200.dataset import TennisSet
from tennis_b200.gluon import MaskedSoftmaxCELoss
from tennis_b200.metrics.vision import compute_bleu
from tennis_b200.models.captioning.gnmt import BeamSearchScorer
from tennis_b200.utils.captioning import evaluate_captioner, get_comp_str, get_dataloaders, write_sentences
from tennis_b200.utils.translation import BeamSearchTranslator
from tennis_b200.vocab import load_embedding_file

cli.define_captioner_flags(training=False)
FLAGS = flags.FLAGS
loss_function = MaskedSoftmaxCELoss()


def run_eval(data_loader, model, translator, vocab, ctx):
    tic = time.time()
    loss, real, ntok = evaluate_captioner(data_loader, model, loss_function, translator, vocab, ctx)
    dt = time.time() - tic
    logging.info('decoded %d caption tokens in %.2fs (%.1f tokens/sec)', ntok, dt, ntok / max(dt, 1e-9))
    return loss, real


def main(_argv):
    ctx = cli.context()
    exp_dir = os.path.join('models', 'captioning', 'experiments', FLAGS.model_id)
    cli.setup_logging(exp_dir)
    syn = {} if FLAGS.synthetic else None
    train_tf = test_tf = None
    if FLAGS.feats_model is None:  # frames through the (frozen) CNN: host-side geometry as in train_gnmt.py:172-188
        from tennis_b200 import transforms
        train_tf, test_tf = transforms.build_transforms(FLAGS.data_shape)
    data_train = TennisSet(split='train', transform=train_tf, captions=True, max_cap_len=FLAGS.tgt_max_len, every=FLAGS.every,
                           feats_model=FLAGS.feats_model, data_shape=FLAGS.data_shape, synthetic=syn)
    data_val = TennisSet(split='val', transform=test_tf, captions=True, vocab=data_train.vocab, every=FLAGS.every, inference=True,
                         feats_model=FLAGS.feats_model, data_shape=FLAGS.data_shape, synthetic=syn)
    data_test = TennisSet(split='test', transform=test_tf, captions=True, vocab=data_train.vocab, every=FLAGS.every, inference=True,
                          feats_model=FLAGS.feats_model, data_shape=FLAGS.data_shape, synthetic=syn)
    embedding = None
    emb_path = os.path.join('data', FLAGS.emb_file) if FLAGS.emb_file else None
    if emb_path and os.path.exists(emb_path):
        data_train.vocab.set_embedding(load_embedding_file(emb_path))
        embedding = data_train.vocab.embedding.idx_to_vec
    model = cli.build_captioner(ctx, data_train.vocab, embedding)
    path = os.path.join(exp_dir, 'valid_best.params')
    if not os.path.exists(path):
        path, _ = cli.latest_params(exp_dir)
    if path is not None:
        model.load_parameters(path, ctx=ctx)
        logging.info('Loaded model params: %s', path)
    else:
        logging.warning('no checkpoint under %s: evaluating freshly initialised weights', exp_dir)
    translator = BeamSearchTranslator(model=model, beam_size=FLAGS.beam_size,
                                      scorer=BeamSearchScorer(alpha=FLAGS.lp_alpha, K=FLAGS.lp_k),
                                      max_length=FLAGS.tgt_max_len + 100)
    logging.info('Use beam_size=%d, alpha=%s, K=%d', FLAGS.beam_size, FLAGS.lp_alpha, FLAGS.lp_k)
    _, val_loader, test_loader = get_dataloaders(data_train, data_val, data_test, FLAGS.batch_size, FLAGS.test_batch_size,
                                                 FLAGS.num_buckets)
    for name, loader, data in (('val', val_loader, data_val), ('test', test_loader, data_test)):
        refs = data.get_captions(split=True)
        loss, out = run_eval(loader, model, translator, data_train.vocab, ctx)
        bleu, _, bp, ref_len, out_len = compute_bleu([[r] for r in refs], out)
        logging.info('%s loss=%.4f ppl=%.4f bleu=%.2f (bp %.3f, ref %d, out %d)', name, loss, np.exp(min(loss, 50)), bleu * 100, bp,
                     ref_len, out_len)
        write_sentences(out, os.path.join(exp_dir, 'best_%s_out.txt' % name))
        write_sentences(refs, os.path.join(exp_dir, '%s_gt.txt' % name))
        print(get_comp_str(refs[:2], out[:2]))


if __name__ == '__main__':
    app.run(main)
