"""Local (in-process) index with the surface of the reference SDK's offline index
(`embeddinghub/sdk/python/offlinehub.py`, class `Index`, lines 27-141): same constructor
arguments and the same method names and return conventions — `set`, `get`, `multiset`,
`multiget`, `nearest_neighbor(num, key=None, embedding=None)`, `size` — so the reference's
`sdk/python/test/offlinehub_test.py` cases read the same against it.

Differences in substance: the vectors live in the ehb200 index on the GPU (metric "l2" like the
reference's `hnswlib.Index("l2", dims)`), inserts go through the batched C-ABI call, and capacity
management is the library's business (the reference doubles `_cap` by hand, lines 136-141).
"""
from collections.abc import Mapping

import numpy as np

from ._native import NativeIndex


class Index:
    def __init__(self, key_emb_iter, dims, metric="l2", device=0):
        self._dims = int(dims)
        self._native = NativeIndex(self._dims, metric=metric, capacity=1024, device=device)
        self._vectors = {}      # key -> python list, exactly what the caller stored (get() returns it)
        self._label_of = {}     # key -> integer label used inside the native index
        self._key_of = []       # label -> key
        self.multiset(key_emb_iter)

    # -- writes ---------------------------------------------------------------------------------------
    def multiset(self, embedding_tuples):
        pairs = list(embedding_tuples.items()) if isinstance(embedding_tuples, Mapping) else list(embedding_tuples)
        if not pairs:
            return                       # the reference guards hnswlib's empty add_items the same way
        labels = np.fromiter((self._label(k) for k, _ in pairs), dtype=np.uint64, count=len(pairs))
        self._native.add(np.asarray([v for _, v in pairs], dtype=np.float32), labels)
        self._vectors.update(pairs)

    def set(self, key, embedding):
        self.multiset([(key, embedding)])

    def _label(self, key):
        label = self._label_of.get(key)
        if label is None:
            label = self._label_of[key] = len(self._key_of)
            self._key_of.append(key)
        return label

    # -- reads ----------------------------------------------------------------------------------------
    def get(self, key):
        return self._vectors[key]

    def multiget(self, keys):
        return [self._vectors[k] for k in keys]

    def size(self):
        return len(self._vectors)

    def nearest_neighbor(self, num, key=None, embedding=None):
        """Keys of the `num` nearest stored embeddings, nearest first.  With `key`, that key's own
        embedding is the query and the key itself is left out of the answer (the reference asks
        hnswlib for num+1 and filters, offlinehub.py:110-130)."""
        by_key = key is not None
        query = self._vectors[key] if by_key else embedding
        want = num + 1 if by_key else num
        labels, _, counts = self._native.search(np.asarray(query, np.float32)[None, :], want)
        found = [self._key_of[int(l)] for l in labels[0][: counts[0]]]
        if by_key:
            found = [k for k in found if k != key][:num]
        return found
