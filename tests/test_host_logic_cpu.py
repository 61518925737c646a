"""Host-side logic on CPU: the mirrors of the reference interfaces (ANNIndex, offline Index, EmbeddingHub
request semantics, the gRPC front end) run against a test double of the native index — an exact numpy
search standing in for the GPU — so key/label bookkeeping, key-mode self removal, status codes, batching
and the wire contract are covered without a device.  (The same cases run against the real CUDA path in
tests/test_gpu_host.py.)"""
import numpy as np
import pytest

import embeddinghub_b200.ann_index as ann_mod
import embeddinghub_b200.offline as offline_mod
from embeddinghub_b200 import hub as hub_mod


class FakeNative:
    """Exact L2 k-NN with insert-or-update by label: the contract of NativeIndex the mirrors rely on."""

    def __init__(self, dim, metric="l2", capacity=128, device=0, **kw):
        self.dim, self.vec, self.order, self.dead = int(dim), {}, [], set()

    def add(self, vecs, labels=None):
        vecs = np.asarray(vecs, np.float32).reshape(-1, self.dim)
        labels = range(len(self.order), len(self.order) + len(vecs)) if labels is None else labels
        for l, v in zip(labels, vecs):
            l = int(l)
            if l not in self.vec:
                self.order.append(l)
            self.vec[l] = v.copy()
            self.dead.discard(l)

    def get(self, label):
        if int(label) in self.dead:
            raise KeyError(label)
        return self.vec[int(label)].copy()

    def remove(self, labels):
        for l in np.atleast_1d(labels):
            if int(l) not in self.vec:
                raise KeyError(l)
            self.dead.add(int(l))

    def search_bruteforce(self, q, k, precision=0):
        return self.search(q, k)

    def set_ef(self, ef):
        pass

    def search(self, q, k, ef=0):
        q = np.asarray(q, np.float32).reshape(-1, self.dim)
        labs = np.array([l for l in self.order if l not in self.dead], np.uint64)
        out_l = np.full((len(q), k), np.uint64(0xFFFFFFFFFFFFFFFF))
        out_d = np.full((len(q), k), np.inf, np.float32)
        cnt = np.zeros(len(q), np.uint32)
        if len(labs) and k:
            X = np.stack([self.vec[int(l)] for l in labs])
            d = ((q[:, None, :] - X[None]) ** 2).sum(-1)
            idx = np.argsort(d, axis=1, kind="stable")[:, :k]
            m = idx.shape[1]
            out_l[:, :m], out_d[:, :m], cnt[:] = labs[idx], np.take_along_axis(d, idx, 1), m
        return out_l, out_d, cnt


@pytest.fixture(autouse=True)
def fake_native(monkeypatch):
    monkeypatch.setattr(ann_mod, "NativeIndex", FakeNative)
    monkeypatch.setattr(offline_mod, "NativeIndex", FakeNative)


EMB = [("a", [1, 0]), ("b", [0, 1]), ("c", [-1, -1]), ("d", [1, 1])]


def test_annindex_reference_cases():
    # embeddinghub/embeddingstore/test/index_test.cc:17-60
    def abc():
        ix = ann_mod.ANNIndex(3)
        for k, v in [("a", [0, 1, 0]), ("b", [1, 1, 0]), ("c", [1, 0, 0])]:
            ix.set(k, v)
        return ix

    assert abc().approx_nearest([0, 1, 0], 1) == ["a"]
    assert abc().approx_nearest([0, 1, 0], 2) == ["a", "b"]
    ix = abc()
    ix.set("a", [0, -1, 0])
    assert ix.approx_nearest([0, 1, 0], 1) == ["b"] and len(ix) == 3
    assert abc().approx_nearest([0, 1, 0], 0) == []
    assert abc().approx_nearest([0, 1, 0], 10) == ["a", "b", "c"]          # fewer points than num: what exists


def test_offline_index_reference_cases():
    # embeddinghub/sdk/python/test/offlinehub_test.py
    index = offline_mod.Index(EMB, 2)
    assert index.get("a") == [1, 0]
    assert index.nearest_neighbor(2, key="a") == ["d", "b"]
    index.multiset({"a": [3, 3], "b": [4, 4]})
    assert index.multiget(["a", "b", "c"]) == [[3, 3], [4, 4], [-1, -1]]
    index = offline_mod.Index([], 2)
    for key in list(range(1028)) * 2:
        index.set(str(key), [1, 1])
    assert index.size() == 1028


def test_hub_semantics_and_status_codes():
    hub = hub_mod.EmbeddingHub()
    with pytest.raises(hub_mod.HubError) as e:
        hub.get("nope", "a")
    assert e.value.code == "NOT_FOUND"
    hub.create_space("s", 2)
    with pytest.raises(hub_mod.HubError) as e:
        hub.create_space("s", 2)
    assert e.value.code == "ALREADY_EXISTS"
    hub.multiset("s", EMB)
    assert hub.nearest_neighbor("s", 2, key="a") == ["d", "b"]                 # server.cc:193-207
    assert hub.nearest_neighbor("s", 3, key="a") == ["d", "b", "c"]
    assert hub.nearest_neighbor("s", 1, embedding=[0.9, 0.1]) == ["a"]
    for kw in (dict(key="a", embedding=[1, 0]), dict()):
        with pytest.raises(hub_mod.HubError) as e:
            hub.nearest_neighbor("s", 1, **kw)
        assert e.value.code == "INVALID_ARGUMENT"                              # server.cc:183-189
    with pytest.raises(hub_mod.HubError) as e:
        hub.nearest_neighbor("s", 1, key="zzz")
    assert e.value.code == "NOT_FOUND"
    hub.freeze_space("s")
    with pytest.raises(hub_mod.HubError) as e:
        hub.set("s", "e", [0, 0])
    assert e.value.code == "FAILED_PRECONDITION"                               # server.cc:124-127
    assert hub.multi_nearest_neighbor("s", 1, embeddings=[[1, 0], [0, 1]]) == [["a"], ["b"]]
    hub.delete_space("s")
    with pytest.raises(hub_mod.HubError):
        hub.get("s", "a")


def test_bad_rows_leave_the_index_untouched_and_deletes_follow_the_docs():
    """a wrong-length embedding must not register its key; docs/reading_and_writing_embeddings.md:49-66."""
    hub = hub_mod.EmbeddingHub()
    hub.create_space("s", 2)
    hub.multiset("s", EMB)
    with pytest.raises(hub_mod.HubError) as e:
        hub.multiset("s", [("e", [1, 2]), ("bad", [1, 2, 3])])
    assert e.value.code == "INVALID_ARGUMENT"
    with pytest.raises(hub_mod.HubError) as e:
        hub.set("s", "one", [7])                                   # no silent broadcast of a length-1 row
    assert e.value.code == "INVALID_ARGUMENT"
    sp = hub._space("s")
    assert "e" not in sp.index and "bad" not in sp.index and "one" not in sp.index and len(sp.index) == 4
    for bad in ([1, 2, 3], [1]):
        with pytest.raises(hub_mod.HubError) as e:
            hub.nearest_neighbor("s", 1, embedding=bad)
        assert e.value.code == "INVALID_ARGUMENT"
    with pytest.raises(hub_mod.HubError) as e:
        hub.nearest_neighbor("s", -1, embedding=[1, 0])
    assert e.value.code == "INVALID_ARGUMENT"
    assert len(hub.nearest_neighbor("s", 600, embedding=[1, 0])) == 4   # beyond the beam limit: exact scan answers
    hub.delete("s", "a")
    assert hub.nearest_neighbor("s", 1, embedding=[0.9, 0.1]) == ["d"]
    with pytest.raises(hub_mod.HubError) as e:
        hub.get("s", "a")
    assert e.value.code == "NOT_FOUND"
    with pytest.raises(hub_mod.HubError) as e:
        hub.delete("s", "a")
    assert e.value.code == "NOT_FOUND"
    assert sorted(sp.index.keys()) == ["b", "c", "d"] and len(sp.index) == 3
    hub.set("s", "a", [1, 0])                                      # re-set un-deletes
    assert hub.nearest_neighbor("s", 1, embedding=[0.9, 0.1]) == ["a"]
    hub.multidelete("s", ["b", "c"])
    hub.delete_all("s")
    assert sp.index.keys() == [] and hub.nearest_neighbor("s", 2, embedding=[1, 0]) == []


def test_grpc_roundtrip_over_the_wire():
    import concurrent.futures as cf

    import grpc

    from embeddinghub_b200 import grpc_server as gs

    server, port = gs.make_server("127.0.0.1:0", hub=hub_mod.EmbeddingHub())
    server.start()
    try:
        stub, M = gs.Stub(grpc.insecure_channel(f"127.0.0.1:{port}")), gs.M
        emb = lambda v: M["Embedding"](values=v)  # noqa: E731
        stub.CreateSpace(M["CreateSpaceRequest"](name="s", dims=2))
        stub.MultiSet(iter([M["MultiSetRequest"](key=k, embedding=emb(v), space="s") for k, v in EMB]))
        assert list(stub.NearestNeighbor(M["NearestNeighborRequest"](num=2, space="s", key="a")).keys) == ["d", "b"]
        with cf.ThreadPoolExecutor(8) as ex:   # concurrent RPC threads call the library directly
            outs = list(ex.map(lambda i: list(stub.NearestNeighbor(M["NearestNeighborRequest"](
                num=1, space="s", embedding=emb([1, 0] if i % 2 else [0, 1]))).keys), range(32)))
        assert outs == [["a"] if i % 2 else ["b"] for i in range(32)]
        with pytest.raises(grpc.RpcError) as e:
            stub.Get(M["GetRequest"](key="a", space="missing"))
        assert e.value.code() == grpc.StatusCode.NOT_FOUND
        assert sorted(r.key for r in stub.Download(M["DownloadRequest"](space="s"))) == ["a", "b", "c", "d"]
        # one malformed request fails alone (INVALID_ARGUMENT), its concurrent neighbours succeed
        def call(i):
            try:
                return list(stub.NearestNeighbor(M["NearestNeighborRequest"](
                    num=1, space="s", embedding=emb([1, 0, 5] if i == 7 else [1, 0]))).keys)
            except grpc.RpcError as err:
                return err.code()
        with cf.ThreadPoolExecutor(8) as ex:
            outs = list(ex.map(call, range(16)))
        assert outs[7] == grpc.StatusCode.INVALID_ARGUMENT and all(o == ["a"] for i, o in enumerate(outs) if i != 7)
        # the side service: batched nearest neighbour on the wire (docs/inference.md:14-22)
        reqs = [M["NearestNeighborRequest"](num=1, space="s", embedding=emb([1, 0])),
                M["NearestNeighborRequest"](num=2, space="s", key="a"),
                M["NearestNeighborRequest"](num=1, space="s", embedding=emb([0, 1]))]
        assert [list(r.keys) for r in stub.MultiNearestNeighbor(iter(reqs))] == [["a"], ["d", "b"], ["b"]]
    finally:
        server.stop(0)
