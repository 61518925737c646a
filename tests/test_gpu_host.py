"""Host-side mirrors of the reference interfaces, exercised on the GPU through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import embeddinghub_b200 as ehb  # noqa: E402
from embeddinghub_b200.offline import Index  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_annindex_twin_runs_reference_index_test_cases():
    exe = os.path.join(ROOT, "tests", "cpp", "ann_index_cases")
    assert os.path.exists(exe), "run make"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok ") == 4


# embeddinghub/sdk/python/test/offlinehub_test.py — same cases, same expectations
EMB = [("a", [1, 0]), ("b", [0, 1]), ("c", [-1, -1]), ("d", [1, 1])]


def test_offlinehub_get_set_multiset():
    index = Index([], 3)
    index.set("a", [1, 2, 3])
    assert index.get("a") == [1, 2, 3]
    index = Index(EMB, 2)
    assert index.get("a") == [1, 0]
    index.set("a", [5, 5])
    assert index.get("a") == [5, 5]
    index.multiset({"a": [3, 3], "b": [4, 4]})
    assert index.multiget(["a", "b", "c"]) == [[3, 3], [4, 4], [-1, -1]]


def test_offlinehub_nn():
    assert Index(EMB, 2).nearest_neighbor(2, key="a") == ["d", "b"]      # offlinehub_test.py:63-65


def test_offlinehub_capacity():
    index = Index([], 2)
    embs = [(str(key), [1, 1]) for key in list(range(1028)) * 2]         # offlinehub_test.py:68-86
    for i in range(0, len(embs), 4):
        index.multiset(embs[i:i + 4])
    assert index.size() == 1028
    assert len(index.nearest_neighbor(5, embedding=[1, 1])) == 5


def test_hub_nearest_neighbor_semantics():
    hub = ehb.EmbeddingHub()
    with pytest.raises(ehb.HubError) as e:
        hub.nearest_neighbor("nope", 1, key="a")
    assert e.value.code == "NOT_FOUND"
    hub.create_space("s", 2)
    hub.multiset("s", EMB)
    assert hub.get("s", "c") == [-1.0, -1.0]
    assert hub.nearest_neighbor("s", 2, key="a") == ["d", "b"]           # key mode drops the key itself
    assert hub.nearest_neighbor("s", 1, embedding=[1, 0]) == ["a"]
    with pytest.raises(ehb.HubError) as e:
        hub.nearest_neighbor("s", 1, key="a", embedding=[1, 0])
    assert e.value.code == "INVALID_ARGUMENT"
    with pytest.raises(ehb.HubError) as e:
        hub.nearest_neighbor("s", 1)
    assert e.value.code == "INVALID_ARGUMENT"
    hub.freeze_space("s")
    with pytest.raises(ehb.HubError) as e:
        hub.set("s", "z", [0, 0])
    assert e.value.code == "FAILED_PRECONDITION"
    res = hub.multi_nearest_neighbor("s", 2, keys=["a", "b"])
    assert res[0] == ["d", "b"] and len(res[1]) == 2 and "b" not in res[1]


def test_concurrent_searches_are_safe():
    import threading

    rng = np.random.default_rng(0)
    base = rng.standard_normal((5000, 32), dtype=np.float32)
    ix = ehb.NativeIndex(32, capacity=5000)
    ix.add(base)
    ix.build()
    ix.set_search_width(2)   # a fixed number of warps per query: concurrent calls are combined into batches of
                             # varying size, and the automatic width depends on the batch size
    q = rng.standard_normal((64, 32), dtype=np.float32)
    ref = ix.search(q, 5, ef=40)[0]
    errs = []

    def work():
        try:
            for _ in range(20):
                assert np.array_equal(ix.search(q, 5, ef=40)[0], ref)
        except Exception as ex:  # pragma: no cover
            errs.append(ex)

    ts = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


def test_grpc_server_speaks_the_reference_contract():
    import grpc

    from embeddinghub_b200 import grpc_server as gs

    server, port = gs.make_server("127.0.0.1:0")
    server.start()
    try:
        stub = gs.Stub(grpc.insecure_channel(f"127.0.0.1:{port}"))
        M = gs.M
        emb = lambda v: M["Embedding"](values=v)  # noqa: E731
        with pytest.raises(grpc.RpcError) as e:
            stub.NearestNeighbor(M["NearestNeighborRequest"](num=1, space="nope", key="a"))
        assert e.value.code() == grpc.StatusCode.NOT_FOUND                       # server.cc:178
        stub.CreateSpace(M["CreateSpaceRequest"](name="s", dims=2))
        stub.Set(M["SetRequest"](key="a", embedding=emb([1, 0]), space="s"))
        stub.MultiSet(iter([M["MultiSetRequest"](key=k, embedding=emb(v), space="s")
                            for k, v in [("b", [0, 1]), ("c", [-1, -1]), ("d", [1, 1])]]))
        assert list(stub.Get(M["GetRequest"](key="c", space="s")).embedding.values) == [-1.0, -1.0]
        got = [list(r.embedding.values) for r in stub.MultiGet(iter([M["MultiGetRequest"](key=k, space="s") for k in "ab"]))]
        assert got == [[1.0, 0.0], [0.0, 1.0]]
        nn = stub.NearestNeighbor(M["NearestNeighborRequest"](num=2, space="s", key="a"))
        assert list(nn.keys) == ["d", "b"]                                        # key mode drops the key itself
        nn = stub.NearestNeighbor(M["NearestNeighborRequest"](num=1, space="s", embedding=emb([1, 0])))
        assert list(nn.keys) == ["a"]
        with pytest.raises(grpc.RpcError) as e:
            stub.NearestNeighbor(M["NearestNeighborRequest"](num=1, space="s", key="a", embedding=emb([1, 0])))
        assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT                 # server.cc:183-186
        assert sorted(r.key for r in stub.Download(M["DownloadRequest"](space="s"))) == ["a", "b", "c", "d"]
        # concurrent embedding-mode calls are coalesced into batched searches
        import concurrent.futures as cf

        with cf.ThreadPoolExecutor(16) as ex:
            outs = list(ex.map(lambda i: list(stub.NearestNeighbor(
                M["NearestNeighborRequest"](num=1, space="s", embedding=emb([1, 0] if i % 2 else [0, 1]))).keys),
                range(64)))
        assert outs == [["a"] if i % 2 else ["b"] for i in range(64)]
        stub.FreezeSpace(M["FreezeSpaceRequest"](name="s"))
        with pytest.raises(grpc.RpcError) as e:
            stub.Set(M["SetRequest"](key="z", embedding=emb([0, 0]), space="s"))
        assert e.value.code() == grpc.StatusCode.FAILED_PRECONDITION              # server.cc:124-127
    finally:
        server.stop(0)
