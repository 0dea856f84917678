"""Captioner training (same CLI surface and flow as the reference's train_gnmt.py:120-497; its latent defects -- per-epoch
outputs written under a directory that is never created, the mislabelled "Best model" format strings, the unused --clip --
are handled as listed in SURVEY.md Appendix C #6-#8).

    python train_gnmt.py --feats_model 0006 --cell_type lstm --epochs 2 --synthetic

Forward, MaskedSoftmaxCELoss, backward and the Adam update run in libtennis_b200.so (csrc/tn_seq_train.cu, tn_train.cu)
through tennis_b200/models/captioning/train_graph.py; validation/test translation uses the device beam search.
NLGEval / TensorBoard of the reference are optional third-party reporters and are not part of this path (BLEU is computed)."""
import logging
import os
import time

import numpy as np
import torch
from absl import app, flags

from tennis_b200 import autograd, cli
from tennis_b200.dataset import TennisSet
from tennis_b200.gluon import MaskedSoftmaxCELoss, Trainer
from tennis_b200.metrics.vision import compute_bleu
from tennis_b200.models.captioning.gnmt import BeamSearchScorer
from tennis_b200.utils.captioning import evaluate_captioner, get_comp_str, get_dataloaders, write_sentences
from tennis_b200.utils.translation import BeamSearchTranslator
from tennis_b200.vocab import load_embedding_file

cli.define_captioner_flags(training=True)
FLAGS = flags.FLAGS


def evaluate(data_loader, model, loss_function, translator, data_train, ctx):
    loss, out, _ = evaluate_captioner(data_loader, model, loss_function, translator, data_train.vocab, ctx)
    return loss, out


def train(data_train, data_val, data_test, model, loss_function, val_tgt_sentences, test_tgt_sentences, translator,
          start_epoch, ctx, exp_dir):
    """train_gnmt.py:298-497."""
    trainer = Trainer(model.collect_params(), FLAGS.optimizer, {'learning_rate': FLAGS.lr})
    train_data_loader, val_data_loader, test_data_loader = get_dataloaders(data_train, data_val, data_test, FLAGS.batch_size,
                                                                           FLAGS.test_batch_size, FLAGS.num_buckets)
    best_valid_bleu = 0.0
    if cli.world()[1] > 1:
        # deferred-shape parameters (the encoder's input size comes from the data) are created by the first forward; a rank
        # whose shard of the first batch is empty would otherwise enter Trainer.step with fewer parameters than its peers
        src0, tgt0, svl0, tvl0 = next(iter(train_data_loader))
        model(src0[:1].to(ctx).float(), tgt0[:1, :-1].to(ctx).float(), svl0[:1].to(ctx), tvl0[:1].to(ctx) - 1)
        cli.broadcast_parameters(model)
    for epoch_id in range(start_epoch, FLAGS.epochs):
        log_avg_loss, log_wc, nlog = 0.0, 0.0, 0
        log_start_time = time.time()
        for batch_id, (src_seq, tgt_seq, src_valid_length, tgt_valid_length) in enumerate(train_data_loader):
            # loss = loss_function(...).mean() * (T - 1) / (tgt_valid_length - 1).mean(); loss.backward()   (:330-334)
            # the normalisers are those of the GLOBAL batch; with --num_gpus N every process takes its contiguous shard of the
            # sentences (split_and_load, train_gnmt.py:322-327) and Trainer.step sums the gradients over the ranks
            n_global = tgt_seq.shape[0]
            scale = float((tgt_seq.shape[1] - 1) / (tgt_valid_length - 1).cpu().numpy().mean())
            src_seq, tgt_seq = cli.rank_shard(src_seq), cli.rank_shard(tgt_seq)
            src_valid_length, tgt_valid_length = cli.rank_shard(src_valid_length), cli.rank_shard(tgt_valid_length)
            if src_seq.shape[0] == 0:  # empty shard of a short batch: contribute zero gradients
                trainer.step(1)
                continue
            src_seq, tgt_seq = src_seq.to(ctx).float(), tgt_seq.to(ctx).float()
            src_valid_length, tgt_valid_length = src_valid_length.to(ctx), tgt_valid_length.to(ctx)
            with autograd.record():
                out, _ = model(src_seq, tgt_seq[:, :-1], src_valid_length, tgt_valid_length - 1)
                loss_vec = loss_function(out, tgt_seq[:, 1:], tgt_valid_length - 1)
            autograd.backward([loss_vec], [torch.full_like(loss_vec, scale / n_global)])
            trainer.step(1)
            step_loss = float(loss_vec.cpu().numpy().mean()) * scale
            log_avg_loss += step_loss
            nlog += 1
            log_wc += float(src_valid_length.cpu().numpy().sum() + (tgt_valid_length - 1).cpu().numpy().sum())
            if (batch_id + 1) % FLAGS.log_interval == 0 or batch_id + 1 == len(train_data_loader):
                wps = log_wc / max(time.time() - log_start_time, 1e-9)
                logging.info('[Epoch %d Batch %d/%d] loss=%.4f, ppl=%.4f  throughput=%.2fK wps, wc=%.2fK', epoch_id, batch_id + 1,
                             len(train_data_loader), log_avg_loss / nlog, np.exp(min(log_avg_loss / nlog, 50)), wps / 1000,
                             log_wc / 1000)
                log_start_time = time.time()
                log_avg_loss, log_wc, nlog = 0.0, 0.0, 0

        valid_loss, valid_translation_out = evaluate(val_data_loader, model, loss_function, translator, data_train, ctx)
        valid_bleu_score, _, _, _, _ = compute_bleu([[r] for r in val_tgt_sentences], valid_translation_out)
        logging.info('[Epoch %d] valid Loss=%.4f, valid ppl=%.4f, valid bleu=%.2f', epoch_id, valid_loss,
                     np.exp(min(valid_loss, 50)), valid_bleu_score * 100)
        test_loss, test_translation_out = evaluate(test_data_loader, model, loss_function, translator, data_train, ctx)
        test_bleu_score, _, _, _, _ = compute_bleu([[r] for r in test_tgt_sentences], test_translation_out)
        logging.info('[Epoch %d] test Loss=%.4f, test ppl=%.4f, test bleu=%.2f', epoch_id, test_loss, np.exp(min(test_loss, 50)),
                     test_bleu_score * 100)
        if cli.is_main():
            write_sentences(valid_translation_out, os.path.join(exp_dir, 'epoch%d_valid_out.txt' % epoch_id))
            write_sentences(test_translation_out, os.path.join(exp_dir, 'epoch%d_test_out.txt' % epoch_id))

        if valid_bleu_score > best_valid_bleu or not os.path.exists(os.path.join(exp_dir, 'valid_best.params')):
            best_valid_bleu = max(best_valid_bleu, valid_bleu_score)
            save_path = os.path.join(exp_dir, 'valid_best.params')
            logging.info('Save best parameters to %s', save_path)
            if cli.is_main():
                model.save_parameters(save_path)
        cli.barrier()
        if epoch_id + 1 >= (FLAGS.epochs * 2) // 3:
            new_lr = trainer.learning_rate * FLAGS.lr_update_factor
            logging.info('Learning rate change to %s', new_lr)
            trainer.set_learning_rate(new_lr)
        if cli.is_main():
            model.save_parameters(os.path.join(exp_dir, '%04d.params' % epoch_id))
        cli.barrier()

    # load and evaluate the best model
    if os.path.exists(os.path.join(exp_dir, 'valid_best.params')):
        model.load_parameters(os.path.join(exp_dir, 'valid_best.params'), ctx=ctx)
    valid_loss, valid_translation_out = evaluate(val_data_loader, model, loss_function, translator, data_train, ctx)
    valid_bleu_score, _, _, _, _ = compute_bleu([[r] for r in val_tgt_sentences], valid_translation_out)
    logging.info('Best model valid Loss=%.4f, valid ppl=%.4f, valid bleu=%.2f', valid_loss, np.exp(min(valid_loss, 50)),
                 valid_bleu_score * 100)
    test_loss, test_translation_out = evaluate(test_data_loader, model, loss_function, translator, data_train, ctx)
    test_bleu_score, _, _, _, _ = compute_bleu([[r] for r in test_tgt_sentences], test_translation_out)
    logging.info('Best model test Loss=%.4f, test ppl=%.4f, test bleu=%.2f', test_loss, np.exp(min(test_loss, 50)),
                 test_bleu_score * 100)
    if cli.is_main():
        write_sentences(valid_translation_out, os.path.join(exp_dir, 'best_valid_out.txt'))
        write_sentences(test_translation_out, os.path.join(exp_dir, 'best_test_out.txt'))
        print(get_comp_str(test_tgt_sentences[:2], test_translation_out[:2]))
    cli.shutdown()


def main(_argv):
    ctx = cli.context()
    exp_dir = os.path.join('models', 'captioning', 'experiments', FLAGS.model_id)
    cli.setup_logging(exp_dir)
    if FLAGS.feats_model is None and not FLAGS.freeze_backbone:
        logging.info('Training the CNN through the captioner: the source-feature gradient of the encoder is routed into the '
                     'backbone (train_gnmt.py:150-170 of the reference); activations of every source frame are kept, use a '
                     'small --batch_size / --every')
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
    val_tgt_sentences = data_val.get_captions(split=True)
    test_tgt_sentences = data_test.get_captions(split=True)
    if cli.is_main():
        write_sentences(val_tgt_sentences, os.path.join(exp_dir, 'val_gt.txt'))
        write_sentences(test_tgt_sentences, os.path.join(exp_dir, 'test_gt.txt'))

    embedding = None
    emb_path = os.path.join('data', FLAGS.emb_file) if FLAGS.emb_file else None
    if emb_path and os.path.exists(emb_path):
        data_train.vocab.set_embedding(load_embedding_file(emb_path))
        embedding = data_train.vocab.embedding.idx_to_vec
    model = cli.build_captioner(ctx, data_train.vocab, embedding)

    path, start_epoch = cli.latest_params(exp_dir)
    if path is not None:
        model.load_parameters(path, ctx=ctx)
        logging.info('Loaded model params: %s', path)
    cli.broadcast_parameters(model)
    translator = BeamSearchTranslator(model=model, beam_size=FLAGS.beam_size,
                                      scorer=BeamSearchScorer(alpha=FLAGS.lp_alpha, K=FLAGS.lp_k),
                                      max_length=FLAGS.tgt_max_len + 100)
    logging.info('Use beam_size=%d, alpha=%s, K=%d', FLAGS.beam_size, FLAGS.lp_alpha, FLAGS.lp_k)
    loss_function = MaskedSoftmaxCELoss()
    train(data_train, data_val, data_test, model, loss_function, val_tgt_sentences, test_tgt_sentences, translator, start_epoch,
          ctx, exp_dir)


if __name__ == '__main__':
    app.run(main)
