"""Range-sharded multi-GPU search: one process per GPU, base vectors partitioned
by contiguous label range, every rank owns an independent graph over its range, every rank
searches all queries, ONE all-gather of the per-shard top-k, then the merge kernel
(ehb_merge_topk_dev).  There is no collective on the write path: inserts route by label.

torch.distributed is only the plumbing (process group + the all-gather over NCCL/NVLink).
"""
import ctypes as C

import numpy as np

from ._native import check, lib


def shard_bounds(n_total, world):
    """Contiguous label ranges [lo, hi) per rank: rank g owns [g*N/G, (g+1)*N/G)."""
    return [(n_total * g) // world for g in range(world + 1)]


def owner_of(labels, n_total, world):
    """Rank that owns each label (vectorised)."""
    b = np.asarray(shard_bounds(n_total, world)[1:], dtype=np.uint64)
    return np.searchsorted(b, np.asarray(labels, dtype=np.uint64), side="right").astype(np.int64)


def route_rows(vecs, labels, n_total, world, rank):
    """Rows of an insert batch that belong to this rank (no communication needed)."""
    mine = owner_of(labels, n_total, world) == rank
    return np.asarray(vecs)[mine], np.asarray(labels)[mine]


def gather_topk(local_labels, local_dists, world, group=None):
    """The single exchange step: all ranks contribute [nq, k] (int64-viewed u64 labels, f32
    distances) and receive [world, nq, k].  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    nq, k = local_labels.shape
    gl = torch.empty((world, nq, k), dtype=local_labels.dtype, device=local_labels.device)
    gd = torch.empty((world, nq, k), dtype=local_dists.dtype, device=local_dists.device)
    if world == 1:
        gl[0], gd[0] = local_labels, local_dists
    else:
        dist.all_gather_into_tensor(gl.view(-1), local_labels.contiguous().view(-1), group=group)
        dist.all_gather_into_tensor(gd.view(-1), local_dists.contiguous().view(-1), group=group)
    return gl, gd


class ShardedSearcher:
    """Search over a range-sharded index.  `index` is this rank's NativeIndex (global labels)."""

    def __init__(self, index, world, device, group=None):
        import torch

        self.ix, self.world, self.device, self.group = index, world, device, group
        self._torch = torch
        self._buf = {}

    def _bufs(self, nq, k):
        t = self._torch
        key = (nq, k)
        if key not in self._buf:
            dev = t.device("cuda", self.device)
            self._buf[key] = dict(
                l=t.empty((nq, k), dtype=t.int64, device=dev), d=t.empty((nq, k), dtype=t.float32, device=dev),
                c=t.empty(nq, dtype=t.int32, device=dev), ml=t.empty((nq, k), dtype=t.int64, device=dev),
                md=t.empty((nq, k), dtype=t.float32, device=dev), mc=t.empty(nq, dtype=t.int32, device=dev))
        return self._buf[key]

    def search_dev(self, q, k, ef, stream_ptr, bruteforce=False, precision=0):
        """q: CUDA float32 tensor [nq, dim].  Returns (labels int64-viewed-u64, dists, counts) CUDA tensors
        holding the global top-k on every rank.  Nothing synchronises the host."""
        nq = q.shape[0]
        b = self._bufs(nq, k)
        if bruteforce:
            self.ix.search_bruteforce_dev(q.data_ptr(), nq, k, precision, b["l"].data_ptr(), b["d"].data_ptr(),
                                          b["c"].data_ptr(), stream_ptr)
        else:
            self.ix.search_dev(q.data_ptr(), nq, k, ef, b["l"].data_ptr(), b["d"].data_ptr(), b["c"].data_ptr(),
                               stream_ptr)
        if self.world == 1:
            return b["l"], b["d"], b["c"]
        gl, gd = gather_topk(b["l"], b["d"], self.world, self.group)
        check(lib().ehb_merge_topk_dev(self.world, nq, k, C.c_void_p(gd.data_ptr()), C.c_void_p(gl.data_ptr()),
                                       C.c_void_p(b["md"].data_ptr()), C.c_void_p(b["ml"].data_ptr()),
                                       C.c_void_p(b["mc"].data_ptr()), self.device, C.c_void_p(stream_ptr)))
        return b["ml"], b["md"], b["mc"]
