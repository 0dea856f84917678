"""Caption batching (mirror of the reference's utils/captioning.py + the gluonnlp pieces it uses):
Pad(0)/Stack batchify, FixedBucketSampler with constant-width buckets, sentence IO."""
import io

import numpy as np
import torch


def pad_stack(seqs, pad_val=0):
    """gluonnlp batchify.Pad(axis=0, pad_val=0): pad along the first axis to the longest sample, then stack."""
    seqs = [torch.as_tensor(s) for s in seqs]
    n = max(s.shape[0] for s in seqs)
    out = torch.full((len(seqs), n) + tuple(seqs[0].shape[1:]), pad_val, dtype=seqs[0].dtype)
    for i, s in enumerate(seqs):
        out[i, : s.shape[0]] = s
    return out


def train_batchify(samples):
    src, tgt, sl, tl = zip(*samples)
    return pad_stack(src), pad_stack(tgt), torch.tensor(sl, dtype=torch.float32), torch.tensor(tl, dtype=torch.float32)


def test_batchify(samples):
    src, tgt, sl, tl, idx = zip(*samples)
    return (pad_stack(src), pad_stack(tgt), torch.tensor(sl, dtype=torch.float32), torch.tensor(tl, dtype=torch.float32),
            torch.tensor(idx))


class FixedBucketSampler(object):
    """gluonnlp FixedBucketSampler with ConstWidthBucket: `num_buckets` equal-width length buckets (keyed on the max of a
    sample's lengths when tuples are given), batches drawn bucket by bucket."""

    def __init__(self, lengths, batch_size, num_buckets=5, shuffle=False, seed=0):
        keys = [max(l) if isinstance(l, (tuple, list)) else l for l in lengths]
        lo, hi = min(keys), max(keys)
        width = max(1, int(np.ceil((hi - lo + 1) / float(num_buckets))))
        buckets = [[] for _ in range(num_buckets)]
        for i, k in enumerate(keys):
            buckets[min((k - lo) // width, num_buckets - 1)].append(i)
        self._buckets, self._batch_size = buckets, batch_size
        self._shuffle, self._rng = shuffle, np.random.RandomState(seed)
        self._n = sum((len(b) + batch_size - 1) // batch_size for b in buckets)

    def __iter__(self):
        """shuffle=True: like gluonnlp, the samples of every bucket are reshuffled each epoch before they are cut into batches
        (batch membership changes from epoch to epoch), and the batch order is shuffled as well."""
        batches = []
        for b in self._buckets:
            idx = list(b)
            if self._shuffle:
                self._rng.shuffle(idx)
            for j in range(0, len(idx), self._batch_size):
                batches.append(idx[j: j + self._batch_size])
        if self._shuffle:
            self._rng.shuffle(batches)
        for bt in batches:
            yield bt

    def __len__(self):
        return self._n


class DataLoader(object):
    def __init__(self, dataset, batch_sampler, batchify_fn):
        self._d, self._s, self._f = dataset, batch_sampler, batchify_fn

    def __iter__(self):
        for idxs in self._s:
            yield self._f([self._d[i] for i in idxs])

    def __len__(self):
        return len(self._s)


def get_dataloaders(data_train, data_val, data_test, batch_size=128, test_batch_size=32, num_buckets=5):
    """reference utils/captioning.py:28-86 (bucket_scheme 'constant')."""
    tr = DataLoader(data_train, FixedBucketSampler(data_train.get_data_lens(), batch_size, num_buckets, shuffle=True),
                    train_batchify)
    va = DataLoader(data_val, FixedBucketSampler([l[-1] for l in data_val.get_data_lens()], test_batch_size, num_buckets),
                    test_batchify)
    te = DataLoader(data_test, FixedBucketSampler([l[-1] for l in data_test.get_data_lens()], test_batch_size, num_buckets),
                    test_batchify)
    return tr, va, te


def write_sentences(sentences, file_path):
    with io.open(file_path, 'w', encoding='utf-8') as of:
        for sent in sentences:
            of.write((u' '.join(sent) if isinstance(sent, (list, tuple)) else sent) + u'\n')


def read_sentences(file_path):
    with io.open(file_path, 'r', encoding='utf-8') as f:
        return [line.rstrip('\n').split() for line in f]


def get_comp_str(tgts, prds):
    out = ''
    for tgt, prd in zip(tgts, prds):
        out += 'GT:\t' + (' '.join(tgt) if isinstance(tgt, (list, tuple)) else tgt) + '\n'
        out += '\nPD:\t' + (' '.join(prd) if isinstance(prd, (list, tuple)) else prd) + '\n\n\n'
    return out


def evaluate_captioner(data_loader, model, loss_function, translator, vocab, ctx):
    """The `evaluate` of train_gnmt.py:258-295 / evaluate_gnmt.py: teacher-forced MaskedSoftmaxCE loss + beam-search
    translation of every batch, sentences restored to instance order.  -> (avg loss, sentences, caption tokens decoded)."""
    translation_out, all_ids = [], []
    avg_loss, denom, ntok = 0.0, 0, 0
    for src, tgt, src_vl, tgt_vl, inst_ids in data_loader:
        src, tgt = src.to(ctx).float(), tgt.to(ctx).float()
        src_vl, tgt_vl = src_vl.to(ctx), tgt_vl.to(ctx)
        out, _ = model(src, tgt[:, :-1], src_vl, tgt_vl - 1)
        loss = float(loss_function(out, tgt[:, 1:], tgt_vl - 1).cpu().numpy().mean())
        all_ids.extend(inst_ids.tolist())
        avg_loss += loss * (tgt.shape[1] - 1)
        denom += tgt.shape[1] - 1
        samples, _, vlen = translator.translate(src_seq=src, src_valid_length=src_vl)
        best, vbest = samples[:, 0, :].cpu().numpy(), vlen[:, 0].cpu().numpy()
        for i in range(best.shape[0]):  # best beam, BOS/EOS stripped (train_gnmt.py:289-294)
            toks = [vocab.idx_to_token[e] for e in best[i][1:(vbest[i] - 1)]]
            ntok += len(toks)
            translation_out.append(toks)
    real = [None] * len(all_ids)
    for ind, sent in zip(all_ids, translation_out):
        real[ind] = sent
    return avg_loss / max(1, denom), real, ntok
