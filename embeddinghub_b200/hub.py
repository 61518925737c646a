"""In-process mirror of the NearestNeighbor/Set/Get semantics of
EmbeddingHubService (embeddinghub/embeddingstore/server.cc:65-233) on top of ehb200.

Only the request semantics of the k-NN path are reproduced — spaces as named
collections, key XOR embedding validation, key-mode self removal, frozen spaces,
and the status the reference answers with — so a gRPC front end (a "next" row,
SURVEY.md §8f-2) can delegate to it.  Storage is the index itself: `get` reads
the fp32 row back from HBM (the reference reads RocksDB, storage.cc:32-36).
"""
import threading

import numpy as np

from .ann_index import ANNIndex


class HubError(Exception):
    """code mirrors grpc::StatusCode names used by server.cc."""

    def __init__(self, code, message):
        super().__init__(f"{code}: {message}")
        self.code, self.message = code, message


class _Space:
    def __init__(self, dims, metric, device):
        self.dims = dims
        self.index = ANNIndex(dims, metric=metric, device=device)
        self.immutable = False


class EmbeddingHub:
    def __init__(self, device=0):
        self._spaces = {}
        self._device = device
        self._mu = threading.Lock()  # guards the space table and host key maps; searches run concurrently

    # server.cc:65-75
    def create_space(self, name, dims, metric="l2"):
        with self._mu:
            if name in self._spaces:
                raise HubError("ALREADY_EXISTS", "Space already exists")
            self._spaces[name] = _Space(int(dims), metric, self._device)

    # server.cc:87-96
    def freeze_space(self, name):
        self._space(name).immutable = True

    def delete_space(self, name):
        with self._mu:
            self._spaces.pop(name, None)

    def _space(self, name):
        sp = self._spaces.get(name)
        if sp is None:
            raise HubError("NOT_FOUND", "Not found")  # server.cc:104,120,140,159,178
        return sp

    # server.cc:113-129 / 131-149
    def set(self, space, key, embedding):
        self.multiset(space, [(key, embedding)])

    def multiset(self, space, items):
        sp = self._space(space)
        if sp.immutable:
            raise HubError("FAILED_PRECONDITION", "Cannot write to immutable space")  # server.cc:124-127
        with self._mu:
            try:
                sp.index.multiset(items)
            except ValueError as e:          # wrong embedding length: the caller's fault, nothing was stored
                raise HubError("INVALID_ARGUMENT", str(e)) from None

    # docs/reading_and_writing_embeddings.md:49-66
    def delete(self, space, key):
        self.multidelete(space, [key])

    def multidelete(self, space, keys):
        sp = self._space(space)
        if sp.immutable:
            raise HubError("FAILED_PRECONDITION", "Cannot write to immutable space")
        with self._mu:
            try:
                sp.index.multidelete(keys)
            except KeyError:
                raise HubError("NOT_FOUND", "Key not found") from None

    def delete_all(self, space):
        sp = self._space(space)
        if sp.immutable:
            raise HubError("FAILED_PRECONDITION", "Cannot write to immutable space")
        with self._mu:
            sp.index.delete_all()

    # server.cc:98-111 / 151-170
    def get(self, space, key):
        sp = self._space(space)
        if key not in sp.index:
            raise HubError("NOT_FOUND", "Key not found")
        return sp.index.get(key).tolist()

    def multiget(self, space, keys):
        return [self.get(space, k) for k in keys]

    # server.cc:172-210
    def nearest_neighbor(self, space, num, key="", embedding=None):
        return self.multi_nearest_neighbor(space, num, keys=[key] if key else None,
                                           embeddings=None if embedding is None else [embedding])[0]

    def multi_nearest_neighbor(self, space, num, keys=None, embeddings=None, ef=0):
        """Batched NearestNeighbor (docs/inference.md:14-22: promised, never implemented upstream)."""
        sp = self._space(space)
        if num < 0 or num > 2047:
            raise HubError("INVALID_ARGUMENT", "num must be in 0..2047")
        has_key = bool(keys)
        has_vec = embeddings is not None and len(embeddings) != 0
        if has_key and has_vec:
            raise HubError("INVALID_ARGUMENT", "Key and embedding cannot both be set")   # server.cc:183-186
        if not has_key and not has_vec:
            raise HubError("INVALID_ARGUMENT", "Key or embedding must be set")           # server.cc:187-189
        if has_key:
            for k in keys:
                if k not in sp.index:
                    raise HubError("NOT_FOUND", "Key not found")
            q = np.stack([sp.index.get(k) for k in keys])
            res = sp.index.approx_nearest_batch(q, num + 1, ef)      # server.cc:198
            out = []
            for k, r in zip(keys, res):
                if k in r:                                           # server.cc:205-207
                    r.remove(k)
                else:
                    r = r[:-1] if len(r) > num else r
                out.append(r[:num])
            return out
        try:
            q = np.asarray(embeddings, np.float32)
        except ValueError:
            raise HubError("INVALID_ARGUMENT", "embeddings have different lengths") from None
        if q.ndim != 2 or q.shape[1] != sp.dims:
            raise HubError("INVALID_ARGUMENT", f"embedding length must be {sp.dims}")
        return sp.index.approx_nearest_batch(q, num, ef)
