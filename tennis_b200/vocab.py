"""gluonnlp.Vocab stand-in with the index convention the captioner relies on:
<unk>=0, <pad>=1, <bos>=2, <eos>=3, then tokens by descending frequency (ties: alphabetical)."""
import collections

import numpy as np


class Vocab(object):
    def __init__(self, counter=None, unknown_token='<unk>', padding_token='<pad>', bos_token='<bos>', eos_token='<eos>'):
        self.unknown_token, self.padding_token, self.bos_token, self.eos_token = unknown_token, padding_token, bos_token, eos_token
        self.idx_to_token = [unknown_token, padding_token, bos_token, eos_token]
        if counter:
            special = set(self.idx_to_token)
            items = sorted(counter.items(), key=lambda kv: kv[0])
            items.sort(key=lambda kv: kv[1], reverse=True)
            self.idx_to_token += [t for t, _ in items if t not in special]
        self.token_to_idx = {t: i for i, t in enumerate(self.idx_to_token)}
        self.embedding = None

    def __len__(self):
        return len(self.idx_to_token)

    def __getitem__(self, tokens):
        if isinstance(tokens, (list, tuple)):
            return [self.token_to_idx.get(t, 0) for t in tokens]
        return self.token_to_idx.get(tokens, 0)

    def to_tokens(self, ids):
        return [self.idx_to_token[int(i)] for i in ids]

    def set_embedding(self, token_embedding):
        """token_embedding: dict token -> vector (or object with .idx_to_vec/.token_to_idx); tokens missing from the
        file (the four specials) get zero vectors, as gluonnlp does."""
        table = token_embedding if isinstance(token_embedding, dict) else token_embedding.as_dict()
        dim = len(next(iter(table.values())))
        mat = np.zeros((len(self), dim), dtype=np.float32)
        for i, t in enumerate(self.idx_to_token):
            if t in table:
                mat[i] = table[t]
        self.embedding = _Emb(mat)


class _Emb(object):
    def __init__(self, mat):
        self.idx_to_vec = mat


def count_tokens(tokens):
    return collections.Counter(tokens)


def load_embedding_file(path):
    """fastText/GloVe text format written by train_embeddings.py: `token v1 v2 ...` per line (data/embeddings-ex.txt)."""
    table = {}
    with open(path) as f:
        for line in f:
            parts = line.rstrip().split(' ')
            if len(parts) < 3:
                continue
            table[parts[0]] = np.asarray(parts[1:], dtype=np.float32)
    return table
