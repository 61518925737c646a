"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/liboracle.so).

CPU restatement of the hnswlib subset behind embeddinghub's ANNIndex
(reference: embeddinghub/embeddingstore/index.cc:10-52).  See the header of
hnsw_oracle.cpp for what it restates and how it is pinned.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METRICS = {"l2": 0, "ip": 1, "cosine": 2}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_ef.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_count.restype = C.c_uint64
        L.orc_count.argtypes = [C.c_void_p]
        L.orc_capacity.restype = C.c_uint64
        L.orc_capacity.argtypes = [C.c_void_p]
        L.orc_max_level.argtypes = [C.c_void_p]
        L.orc_entry_point.restype = C.c_uint32
        L.orc_entry_point.argtypes = [C.c_void_p]
        L.orc_resize.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_add.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_search.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int]
        L.orc_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_mark_delete.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_deleted_count.restype = C.c_uint64
        L.orc_deleted_count.argtypes = [C.c_void_p]
        L.orc_get_vector.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_upper_rows.restype = C.c_uint64
        L.orc_upper_rows.argtypes = [C.c_void_p]
        L.orc_export_graph.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_vectors.restype = C.c_void_p
        L.orc_vectors.argtypes = [C.c_void_p]
        L.orc_import_graph.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6 + [C.c_uint32, C.c_int]
        L.orc_bruteforce.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                     C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_normalize.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleHNSW:
    """hnswlib.HierarchicalNSW<float> restatement (defaults of index.cc:14-15)."""

    def __init__(self, dim, metric="l2", max_elements=128, M=16, ef_construction=200, seed=100):
        self.dim, self.metric, self.M = int(dim), metric, int(M)
        self._h = lib().orc_create(dim, METRICS[metric], max_elements, M, ef_construction, seed)
        if not self._h:
            raise RuntimeError(lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def _chk(self, rc):
        if rc:
            raise RuntimeError(lib().orc_last_error().decode())

    def set_ef(self, ef):
        lib().orc_set_ef(self._h, ef)

    def resize(self, cap):
        self._chk(lib().orc_resize(self._h, cap))

    @property
    def count(self):
        return lib().orc_count(self._h)

    @property
    def capacity(self):
        return lib().orc_capacity(self._h)

    def add(self, rows, labels=None, threads=1):
        rows = _f32(rows).reshape(-1, self.dim)
        n = rows.shape[0]
        if labels is None:
            labels = np.arange(self.count, self.count + n, dtype=np.uint64)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        self._chk(lib().orc_add(self._h, n, _ptr(rows), _ptr(labels), threads))

    def search(self, q, k, ef=0, threads=1):
        q = _f32(q).reshape(-1, self.dim)
        nq = q.shape[0]
        labels = np.empty((nq, k), dtype=np.uint64)
        dists = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.uint32)
        self._chk(lib().orc_search(self._h, nq, _ptr(q), k, ef, _ptr(labels), _ptr(dists), _ptr(counts), threads))
        return labels, dists, counts

    def metrics(self, reset=False):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().orc_metrics(self._h, C.byref(a), C.byref(b), C.byref(c), int(reset))
        return {"hops_upper": a.value, "hops0": b.value, "evals": c.value}

    def mark_delete(self, label):
        """hnswlib markDelete: KeyError for an unknown label, RuntimeError for a double delete."""
        rc = lib().orc_mark_delete(self._h, int(label))
        if rc == 1:
            raise KeyError(label)
        if rc:
            raise RuntimeError(lib().orc_last_error().decode())

    @property
    def deleted_count(self):
        return lib().orc_deleted_count(self._h)

    def get(self, label):
        out = np.empty(self.dim, dtype=np.float32)
        if lib().orc_get_vector(self._h, label, _ptr(out)):
            raise KeyError(label)
        return out

    def export_graph(self):
        n, M = self.count, self.M
        rows = lib().orc_upper_rows(self._h)
        g = {
            "levels": np.empty(n, np.uint8),
            "links0": np.empty((n, 2 * M), np.uint32),
            "up_off": np.empty(n, np.uint32),
            "links_up": np.empty((max(rows, 1), M), np.uint32),
            "labels": np.empty(n, np.uint64),
        }
        lib().orc_export_graph(self._h, _ptr(g["levels"]), _ptr(g["links0"]), _ptr(g["up_off"]),
                               _ptr(g["links_up"]), _ptr(g["labels"]))
        g["links_up"] = g["links_up"][:rows]
        g["entry"] = lib().orc_entry_point(self._h)
        g["maxlevel"] = lib().orc_max_level(self._h)
        vp = lib().orc_vectors(self._h)
        g["vectors"] = np.ctypeslib.as_array(C.cast(vp, C.POINTER(C.c_float)), shape=(n, self.dim)).copy()
        return g

    def import_graph(self, g):
        v = _f32(g["vectors"])
        n = v.shape[0]
        lu = np.ascontiguousarray(g["links_up"], np.uint32)
        if lu.size == 0:
            lu = np.zeros((1, self.M), np.uint32)
        self._chk(lib().orc_import_graph(
            self._h, n, _ptr(v), _ptr(np.ascontiguousarray(g["labels"], np.uint64)),
            _ptr(np.ascontiguousarray(g["levels"], np.uint8)), _ptr(np.ascontiguousarray(g["links0"], np.uint32)),
            _ptr(np.ascontiguousarray(g["up_off"], np.uint32)), _ptr(lu), int(g["entry"]), int(g["maxlevel"])))


def set_thread_pinning(on=True):
    """Worker t of every parallel pass pins itself to the t-th CPU of the process affinity mask."""
    lib().orc_set_thread_pinning(1 if on else 0)


def bruteforce(base, q, k, metric="l2", threads=1):
    """Exact k-NN, canonical sequential-FMA fp32 arithmetic, order (dist, row)."""
    base, q = _f32(base), _f32(q)
    n, d = base.shape
    q = q.reshape(-1, d)
    idx = np.empty((q.shape[0], k), np.uint64)
    dist = np.empty((q.shape[0], k), np.float32)
    rc = lib().orc_bruteforce(METRICS[metric], n, d, _ptr(base), q.shape[0], _ptr(q), k, _ptr(idx), _ptr(dist),
                              threads)
    if rc:
        raise RuntimeError(lib().orc_last_error().decode())
    return idx, dist


def normalize(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_normalize(x.shape[0], x.shape[1], _ptr(x), _ptr(out))
    return out


class OracleANNIndex:
    """Twin of featureform::embedding::ANNIndex (index.h:19-33, index.cc:10-52):
    string keys <-> labels, capacity doubling from init_cap, insert-or-update,
    approx_nearest returning keys nearest-first.  ef stays at hnswlib's default
    10 because the reference never calls setEf."""

    def __init__(self, dims, init_cap=128, metric="l2"):
        self._cap = init_cap
        self._nn = OracleHNSW(dims, metric, init_cap)
        self._k2l, self._l2k, self._next = {}, {}, 0

    def set(self, key, value):
        if key not in self._k2l:
            label = self._next
            self._next += 1
            self._l2k[label] = key
            self._k2l[key] = label
            if self._next == self._cap:           # index.cc:29-32
                self._cap *= 2
                self._nn.resize(self._cap)
        else:
            label = self._k2l[key]
        self._nn.add(np.asarray(value, np.float32)[None, :], np.array([label], np.uint64))

    def approx_nearest(self, value, num):
        if num == 0:
            return []
        labels, _, counts = self._nn.search(np.asarray(value, np.float32)[None, :], num)
        # index.cc:42-50 would pop `num` entries even when fewer came back (UB);
        # the twin returns what exists.
        return [self._l2k[int(l)] for l in labels[0][: counts[0]]]
