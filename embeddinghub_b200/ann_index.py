"""Host-side mirror of featureform::embedding::ANNIndex
(embeddinghub/embeddingstore/index.h:19-33, index.cc:10-52) over the ehb200 C ABI.

Same surface and semantics: string keys, insert-or-update on a duplicate key,
`approx_nearest(value, num)` returning keys nearest-first, `num == 0 -> []`.
The metric is a per-index parameter defaulting to L2 (the reference hard-codes
hnswlib::L2Space, index.cc:12-13).  Capacity doubling (index.cc:29-32) happens
inside the library.  ef stays at hnswlib's default 10 unless set_ef is called,
exactly as the reference (which never calls setEf).
"""
import numpy as np

from ._native import NativeIndex


class ANNIndex:
    def __init__(self, dims, init_cap=128, metric="l2", device=0, **kw):
        self._nn = NativeIndex(dims, metric=metric, capacity=init_cap, device=device, **kw)
        self._dims = int(dims)
        self._key_to_label = {}
        self._label_to_key = {}
        self._next_label = 0

    # index.cc:20-37
    def set(self, key, value):
        self.multiset([(key, value)])

    def multiset(self, items):
        """Batched form of `set` (the reference loops set() per row, version.cc:69-72)."""
        items = list(items.items()) if hasattr(items, "items") else list(items)
        if not items:
            return
        labels = np.empty(len(items), np.uint64)
        vecs = np.empty((len(items), self._dims), np.float32)
        for i, (key, value) in enumerate(items):
            label = self._key_to_label.get(key)
            if label is None:
                label = self._next_label
                self._next_label += 1
                self._key_to_label[key] = label
                self._label_to_key[label] = key
            labels[i] = label
            vecs[i] = value
        self._nn.add(vecs, labels)

    # index.cc:39-52
    def approx_nearest(self, value, num):
        return self.approx_nearest_batch(np.asarray(value, np.float32)[None, :], num)[0]

    def approx_nearest_batch(self, values, num, ef=0):
        """Batched k-NN (docs/inference.md:14-22 promises multi_nearest_neighbor;
        the reference never implemented it)."""
        if num == 0:
            return [[] for _ in range(len(values))]
        labels, _, counts = self._nn.search(values, num, ef)
        return [[self._label_to_key[int(l)] for l in row[:c]] for row, c in zip(labels, counts)]

    def get(self, key):
        return self._nn.get(self._key_to_label[key])

    def set_ef(self, ef):
        self._nn.set_ef(ef)

    def keys(self):
        """Stored keys in insertion order (what Download streams, server.cc:212-233)."""
        return list(self._key_to_label)

    def __len__(self):
        return self._next_label

    def __contains__(self, key):
        return key in self._key_to_label
