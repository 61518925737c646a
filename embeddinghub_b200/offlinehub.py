"""Mirror of embeddinghub/sdk/python/offlinehub.py (`Index`, lines 27-141) over ehb200.

Same constructor and methods (`set`, `get`, `multiset`, `multiget`,
`nearest_neighbor`, `size`), same key-mode semantics: the queried key is removed
from its own result and exactly `num` keys are returned (offlinehub.py:110-130).
hnswlib.Index("l2", dims) becomes an ehb200 index with metric "l2"; the capacity
bookkeeping (`_add_capacity_to_fit`, offlinehub.py:136-141) is not needed because
the library grows by doubling on its own.
"""
from collections.abc import Mapping

import numpy as np

from ._native import NativeIndex


class Index:
    def __init__(self, key_emb_iter, dims, metric="l2", device=0):
        self._data = {}
        self._dims = dims
        self._idx = NativeIndex(dims, metric=metric, capacity=1024, device=device)
        self._key_to_idx = {}
        self._idx_to_key = {}
        self.multiset(key_emb_iter)

    def _to_idx(self, key):
        idx = self._key_to_idx.get(key)
        if idx is None:
            idx = len(self._key_to_idx)
            self._key_to_idx[key] = idx
            self._idx_to_key[idx] = key
        return idx

    def set(self, key, embedding):
        self.multiset([(key, embedding)])

    def get(self, key):
        return self._data[key]

    def multiset(self, embedding_tuples):
        if isinstance(embedding_tuples, Mapping):
            embedding_tuples = embedding_tuples.items()
        embeddings, idxs = [], []
        for key, embedding in embedding_tuples:
            embeddings.append(embedding)
            idxs.append(self._to_idx(key))
            self._data[key] = embedding
        if not idxs:  # offlinehub.py:86-88
            return
        self._idx.add(np.asarray(embeddings, np.float32), np.asarray(idxs, np.uint64))

    def multiget(self, keys):
        return [self._data[key] for key in keys]

    def nearest_neighbor(self, num, key=None, embedding=None):
        has_key = key is not None
        if has_key:
            embedding = self._data[key]
            num_retrieve = num + 1
        else:
            num_retrieve = num
        labels, _, counts = self._idx.search(np.asarray(embedding, np.float32)[None, :], num_retrieve)
        results = [int(l) for l in labels[0][: counts[0]]]
        if has_key:
            idx = self._key_to_idx[key]
            results = [r for r in results if r != idx]
            if len(results) > num:
                results = results[:-1]
        return [self._idx_to_key[r] for r in results]

    def size(self):
        return len(self._data)
