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
    distances) and receive [world, nq, k] of each.  Labels and distances travel in ONE all-gather
    (packed per rank as [labels bytes | distance bytes]).  Works on CUDA tensors (NCCL) and CPU
    tensors (gloo)."""
    import torch
    import torch.distributed as dist

    nq, k = local_labels.shape
    if world == 1:
        return local_labels.unsqueeze(0), local_dists.unsqueeze(0)
    nl, nd = nq * k * 8, nq * k * 4
    send = torch.empty(nl + nd, dtype=torch.uint8, device=local_labels.device)
    send[:nl].view(torch.int64).copy_(local_labels.contiguous().view(-1))
    send[nl:].view(torch.float32).copy_(local_dists.contiguous().view(-1))
    recv = torch.empty((world, nl + nd), dtype=torch.uint8, device=local_labels.device)
    dist.all_gather_into_tensor(recv.view(-1), send, group=group)
    gl = recv[:, :nl].contiguous().view(torch.int64).view(world, nq, k)
    gd = recv[:, nl:].contiguous().view(torch.float32).view(world, nq, k)
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
            nl, nd = nq * k * 8, nq * k * 4
            send = t.empty(nl + nd, dtype=t.uint8, device=dev)     # [labels | distances], written by the kernels
            self._buf[key] = dict(
                send=send, l=send[:nl].view(t.int64).view(nq, k), d=send[nl:].view(t.float32).view(nq, k),
                recv=t.empty((self.world, nl + nd), dtype=t.uint8, device=dev),
                c=t.empty(nq, dtype=t.int32, device=dev), ml=t.empty((nq, k), dtype=t.int64, device=dev),
                md=t.empty((nq, k), dtype=t.float32, device=dev), mc=t.empty(nq, dtype=t.int32, device=dev))
        return self._buf[key]

    def search_dev(self, q, k, ef, stream_ptr, bruteforce=False, precision=0):
        """q: CUDA float32 tensor [nq, dim].  Returns (labels int64-viewed-u64, dists, counts) CUDA tensors
        holding the global top-k on every rank.  Nothing synchronises the host.  The per-shard kernels write
        straight into the packed send buffer; ONE all-gather; the merge kernel reads the gathered blocks in
        place."""
        import torch.distributed as dist

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
        dist.all_gather_into_tensor(b["recv"].view(-1), b["send"], group=self.group)
        check(lib().ehb_merge_topk_packed_dev(self.world, nq, k, C.c_void_p(b["recv"].data_ptr()),
                                              b["recv"].shape[1], C.c_void_p(b["md"].data_ptr()),
                                              C.c_void_p(b["ml"].data_ptr()), C.c_void_p(b["mc"].data_ptr()),
                                              self.device, C.c_void_p(stream_ptr)))
        return b["ml"], b["md"], b["mc"]
