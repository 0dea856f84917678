"""Per-class precision / recall / F1 with a confusion matrix (the reference's metrics/vision.py semantics, integer
counts on the host).  Quirk kept for identical numbers: `prec` divides matches by the count of POSITIVES (labels) and
`rec` by the count of PREDICTIONS (reference metrics/vision.py:73-74, Appendix C #11); F1 is unaffected."""
import numpy as np


class PRF1(object):
    def __init__(self, axis=1, name='prf1', output_names=None, label_names=None):
        assert label_names is not None, 'label_names cant be None'
        self.name, self.axis, self.label_names = name, axis, label_names
        self.reset()

    def reset(self):
        n = len(self.label_names)
        self.scores = np.zeros((3, n))
        self.mat = np.zeros((n, n))

    def update(self, labels, preds):
        for label, pred in zip(labels, preds):
            label = np.asarray(label.detach().cpu() if hasattr(label, "detach") else label)
            pred = np.asarray(pred.detach().cpu() if hasattr(pred, "detach") else pred)
            if pred.shape != label.shape:
                pred = pred.argmax(axis=self.axis)
            pred, label = pred.astype('int32').reshape(-1), label.astype('int32').reshape(-1)
            np.add.at(self.mat, (label, pred), 1)
            for i in range(len(self.label_names)):
                predictions, positives = pred == i, label == i
                self.scores[0, i] += np.logical_and(predictions, positives).sum()
                self.scores[1, i] += positives.sum()
                self.scores[2, i] += predictions.sum()

    def state_tensor(self):
        """Accumulators as one float64 tensor (summed across ranks by cli.sync_metrics)."""
        import torch
        return torch.from_numpy(np.concatenate([self.scores.reshape(-1), self.mat.reshape(-1)]).astype(np.float64))

    def load_state_tensor(self, t):
        n = len(self.label_names)
        a = t.numpy()
        self.scores = a[:3 * n].reshape(3, n).copy()
        self.mat = a[3 * n:].reshape(n, n).copy()

    def get(self):
        eps = np.finfo(float).eps
        out, ps, rs, fs = [], [], [], []
        for i, c in enumerate(self.label_names):
            prec = self.scores[0][i] / (self.scores[1][i] + eps)
            rec = self.scores[0][i] / (self.scores[2][i] + eps)
            f1 = 2 * (prec * rec) / (prec + rec + eps)
            out += [(c + '_prec', prec), (c + '_rec', rec), (c + '_f1', f1)]
            ps.append(prec)
            rs.append(rec)
            fs.append(f1)
        out += [('AVG_prec', sum(ps) / len(ps)), ('AVG_rec', sum(rs) / len(rs)), ('AVG_f1', sum(fs) / len(fs))]
        out += [('AVG_NB_prec', sum(ps[1:]) / len(ps[1:])), ('AVG_NB_rec', sum(rs[1:]) / len(rs[1:])),
                ('AVG_NB_f1', sum(fs[1:]) / len(fs[1:]))]
        return out


class Accuracy(object):
    def __init__(self, name='accuracy', top_k=1):
        self.name, self.top_k = name, top_k
        self.reset()

    def reset(self):
        self.hit, self.n = 0, 0

    def update(self, labels, preds):
        for label, pred in zip(labels, preds):
            label = np.asarray(label.detach().cpu() if hasattr(label, "detach") else label).reshape(-1)
            pred = np.asarray(pred.detach().cpu() if hasattr(pred, "detach") else pred)
            top = np.argsort(-pred, axis=1, kind='stable')[:, : self.top_k]
            self.hit += int((top == label.reshape(-1, 1)).any(axis=1).sum())
            self.n += label.shape[0]

    def state_tensor(self):
        import torch
        return torch.tensor([self.hit, self.n], dtype=torch.float64)

    def load_state_tensor(self, t):
        self.hit, self.n = int(t[0].item()), int(t[1].item())

    def get(self):
        return self.name, self.hit / max(1, self.n)


def compute_bleu(references, translations, max_n=4, smooth=False):
    """Corpus BLEU with brevity penalty (Papineni et al.): references = list (per sentence) of list of reference token
    lists, translations = list of token lists.  Returns (bleu, precisions, bp, ref_len, trans_len)."""
    import collections
    import math
    match, total = [0] * max_n, [0] * max_n
    ref_len = trans_len = 0

    def ngrams(tokens, n):
        return collections.Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))

    for refs, hyp in zip(references, translations):
        trans_len += len(hyp)
        ref_len += min((abs(len(r) - len(hyp)), len(r)) for r in refs)[1]
        for n in range(1, max_n + 1):
            h = ngrams(hyp, n)
            mx = collections.Counter()
            for r in refs:
                mx |= ngrams(r, n)
            match[n - 1] += sum((h & mx).values())
            total[n - 1] += max(0, len(hyp) - n + 1)
    precisions = []
    for m, t in zip(match, total):
        if smooth:
            precisions.append((m + 1.0) / (t + 1.0))
        else:
            precisions.append(m / t if t > 0 else 0.0)
    bleu = math.exp(sum(math.log(p) for p in precisions) / max_n) if min(precisions) > 0 else 0.0
    bp = 1.0 if trans_len > ref_len else (math.exp(1 - ref_len / trans_len) if trans_len > 0 else 0.0)
    return bleu * bp, precisions, bp, ref_len, trans_len
