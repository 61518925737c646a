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
    """Search over a range-sharded index, one process per GPU.  `index` is this rank's NativeIndex (global
    labels).  exchange="peer" (default on GPUs): the library's peer-memory exchange (ehb_exchange_*: CUDA IPC
    mappings, one push + flag + merge kernel per step, no collective); exchange="nccl": one
    all_gather_into_tensor of the packed per-shard top-k + the merge kernel (kept as the comparison baseline
    and for process groups without peer access)."""

    def __init__(self, index, world, device, group=None, exchange="peer"):
        import torch

        self.ix, self.world, self.device, self.group = index, world, device, group
        self._torch = torch
        self._buf = {}
        self.exchange = exchange if world > 1 else "none"
        self._ex, self._ex_cap = None, (0, 0)

    def close(self):
        if self._ex is not None:
            lib().ehb_exchange_destroy(self._ex)
            self._ex = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ensure_exchange(self, nq, k):
        """(Re)creates the exchange when (nq, k) outgrow it.  Collective: every rank takes the same decision
        because every rank searches the same batch.  The 64-byte IPC handles travel in one all_gather."""
        import torch.distributed as dist

        t = self._torch
        if self._ex is not None and nq <= self._ex_cap[0] and k <= self._ex_cap[1]:
            return
        t.cuda.synchronize()
        if self._ex is not None:
            dist.barrier(group=self.group)   # no peer may still be storing into the buffer that is about to go
        self.close()
        cap = (max(nq, self._ex_cap[0]), max(k, self._ex_cap[1]))
        rank = dist.get_rank(self.group)
        h = C.c_void_p()
        check(lib().ehb_exchange_create(self.device, self.world, rank, cap[0], cap[1], C.byref(h)))
        mine = np.zeros(64, np.uint8)
        check(lib().ehb_exchange_ipc_handle(h, mine.ctypes.data_as(C.c_void_p)))
        dev = t.device("cuda", self.device)
        send = t.from_numpy(mine).to(dev)
        recv = t.empty(self.world * 64, dtype=t.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        handles = np.ascontiguousarray(recv.cpu().numpy())
        check(lib().ehb_exchange_open(h, handles.ctypes.data_as(C.c_void_p)))
        dist.barrier(group=self.group)   # every rank has mapped every peer before the first push
        self._ex, self._ex_cap = h, cap

    def _bufs(self, nq, k):
        t = self._torch
        key = (nq, k)
        if key not in self._buf:
            dev = t.device("cuda", self.device)
            nl, nd = nq * k * 8, nq * k * 4
            send = t.empty(nl + nd, dtype=t.uint8, device=dev)     # [labels | distances], written by the kernels
            self._buf[key] = dict(
                send=send, l=send[:nl].view(t.int64).view(nq, k), d=send[nl:].view(t.float32).view(nq, k),
                recv=t.empty((self.world, nl + nd), dtype=t.uint8, device=dev) if self.exchange == "nccl" else None,
                c=t.empty(nq, dtype=t.int32, device=dev), ml=t.empty((nq, k), dtype=t.int64, device=dev),
                md=t.empty((nq, k), dtype=t.float32, device=dev), mc=t.empty(nq, dtype=t.int32, device=dev))
        return self._buf[key]

    def _local(self, q, nq, k, ef, stream_ptr, bruteforce, precision, lp, dp, cp):
        if bruteforce:
            self.ix.search_bruteforce_dev(q.data_ptr(), nq, k, precision, lp, dp, cp, stream_ptr)
        else:
            self.ix.search_dev(q.data_ptr(), nq, k, ef, lp, dp, cp, stream_ptr)

    def search_dev(self, q, k, ef, stream_ptr, bruteforce=False, precision=0):
        """q: CUDA float32 tensor [nq, dim].  Returns (labels int64-viewed-u64, dists, counts) CUDA tensors
        holding the global top-k on every rank.  Nothing synchronises the host."""
        nq = q.shape[0]
        b = self._bufs(nq, k)
        if self.world == 1:
            self._local(q, nq, k, ef, stream_ptr, bruteforce, precision, b["l"].data_ptr(), b["d"].data_ptr(),
                        b["c"].data_ptr())
            return b["l"], b["d"], b["c"]
        if self.exchange == "peer":
            self._ensure_exchange(nq, k)
            if not bruteforce:
                # fused: the walk's epilogue stores each query's top-k into every peer's receive buffer and raises
                # the slice flags; one kernel waits for the peers' flags and merges
                check(lib().ehb_exchange_search_dev(self._ex, self.ix._h, nq, C.c_void_p(q.data_ptr()), k, ef,
                                                    C.c_void_p(b["md"].data_ptr()), C.c_void_p(b["ml"].data_ptr()),
                                                    C.c_void_p(b["mc"].data_ptr()), C.c_void_p(b["c"].data_ptr()),
                                                    C.c_void_p(stream_ptr)))
                return b["ml"], b["md"], b["mc"]
            lp, dp = C.c_void_p(), C.c_void_p()
            check(lib().ehb_exchange_begin(self._ex, nq, k, C.byref(lp), C.byref(dp)))
            # the shard's kernels write straight into this rank's block of its own receive buffer ...
            self._local(q, nq, k, ef, stream_ptr, bruteforce, precision, lp.value, dp.value, b["c"].data_ptr())
            # ... and one kernel pushes it to the peers, flags, waits for theirs and merges
            check(lib().ehb_exchange_merge_dev(self._ex, C.c_void_p(b["md"].data_ptr()), C.c_void_p(b["ml"].data_ptr()),
                                               C.c_void_p(b["mc"].data_ptr()), C.c_void_p(stream_ptr)))
            return b["ml"], b["md"], b["mc"]
        import torch.distributed as dist

        self._local(q, nq, k, ef, stream_ptr, bruteforce, precision, b["l"].data_ptr(), b["d"].data_ptr(),
                    b["c"].data_ptr())
        dist.all_gather_into_tensor(b["recv"].view(-1), b["send"], group=self.group)
        check(lib().ehb_merge_topk_packed_dev(self.world, nq, k, C.c_void_p(b["recv"].data_ptr()),
                                              b["recv"].shape[1], C.c_void_p(b["md"].data_ptr()),
                                              C.c_void_p(b["ml"].data_ptr()), C.c_void_p(b["mc"].data_ptr()),
                                              self.device, C.c_void_p(stream_ptr)))
        return b["ml"], b["md"], b["mc"]
