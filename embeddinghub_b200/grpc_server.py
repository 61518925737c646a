"""gRPC front end speaking the reference's wire contract.

Service `featureform.embedding.proto.EmbeddingHub` with its nine RPCs and message field numbers
exactly as declared in embeddinghub/embeddingstore/embedding_store.proto:9-106, so the reference's
Python SDK (embeddinghub/sdk/python/embeddinghub.py) can talk to this server unchanged.  The .proto
file itself is not copied: the descriptors are restated programmatically below (grpcio is present in
this image, protoc-generated stubs are not usable with its protobuf runtime).

Request handling delegates to embeddinghub_b200.hub.EmbeddingHub, which mirrors the semantics and
status codes of EmbeddingHubService (embeddingstore/server.cc:65-233).  The reference serialises
every RPC under one mutex (server.cc:175) and answers one query per call; here concurrent
NearestNeighbor calls are coalesced by a micro-batcher into one batched GPU search (the batched
call the reference's docs promise, docs/inference.md:14-22).
"""
import queue
import threading
import time
from concurrent import futures

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from .hub import EmbeddingHub, HubError

PKG = "featureform.embedding.proto"
SERVICE = PKG + ".EmbeddingHub"
DEFAULT_ADDRESS = "0.0.0.0:7462"  # embeddingstore/main.cc:8

_T = descriptor_pb2.FieldDescriptorProto


def _file_descriptor():
    f = descriptor_pb2.FileDescriptorProto()
    f.name = "embeddingstore/embedding_store.proto"
    f.package = PKG
    f.syntax = "proto3"

    def msg(name, *fields):
        m = f.message_type.add()
        m.name = name
        for fname, num, ftype, label, tname in fields:
            fd = m.field.add()
            fd.name, fd.number, fd.type, fd.label = fname, num, ftype, label
            if tname:
                fd.type_name = "." + PKG + "." + tname
        return m

    opt, rep = _T.LABEL_OPTIONAL, _T.LABEL_REPEATED
    s, u32, i32, fl, ms = _T.TYPE_STRING, _T.TYPE_UINT32, _T.TYPE_INT32, _T.TYPE_FLOAT, _T.TYPE_MESSAGE
    emb = ("embedding", None, ms, opt, "Embedding")
    # field numbers: embedding_store.proto:21-106
    msg("DeleteSpaceRequest", ("name", 1, s, opt, None))
    msg("DeleteSpaceResponse")
    msg("CreateSpaceRequest", ("name", 1, s, opt, None), ("dims", 2, u32, opt, None))
    msg("CreateSpaceResponse")
    msg("FreezeSpaceRequest", ("name", 1, s, opt, None))
    msg("FreezeSpaceResponse")
    msg("SetRequest", ("key", 1, s, opt, None), (emb[0], 2, ms, opt, emb[4]), ("space", 3, s, opt, None))
    msg("SetResponse")
    msg("GetRequest", ("key", 1, s, opt, None), ("space", 2, s, opt, None))
    msg("GetResponse", (emb[0], 1, ms, opt, emb[4]))
    msg("MultiSetRequest", ("key", 1, s, opt, None), (emb[0], 2, ms, opt, emb[4]), ("space", 3, s, opt, None))
    msg("MultiSetResponse")
    msg("MultiGetRequest", ("key", 1, s, opt, None), ("space", 2, s, opt, None))
    msg("MultiGetResponse", (emb[0], 1, ms, opt, emb[4]))
    msg("NearestNeighborRequest", ("num", 1, i32, opt, None), ("space", 2, s, opt, None), ("key", 3, s, opt, None),
        (emb[0], 4, ms, opt, emb[4]))
    msg("NearestNeighborResponse", ("keys", 1, s, rep, None))
    msg("DownloadRequest", ("space", 1, s, opt, None))
    msg("DownloadResponse", ("key", 1, s, opt, None), (emb[0], 2, ms, opt, emb[4]))
    msg("Embedding", ("values", 1, fl, rep, None))

    svc = f.service.add()
    svc.name = "EmbeddingHub"
    for name, cs, ss in RPCS:
        m = svc.method.add()
        m.name = name
        m.input_type = "." + PKG + "." + name + "Request"
        m.output_type = "." + PKG + "." + name + "Response"
        m.client_streaming, m.server_streaming = cs, ss
    return f


# name, client streaming, server streaming — embedding_store.proto:9-19
RPCS = [("CreateSpace", False, False), ("DeleteSpace", False, False), ("FreezeSpace", False, False),
        ("Set", False, False), ("Get", False, False), ("MultiSet", True, False), ("MultiGet", True, True),
        ("NearestNeighbor", False, False), ("Download", False, True)]

_pool = descriptor_pool.DescriptorPool()
_pool.Add(_file_descriptor())


def message_class(name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName(PKG + "." + name))


M = {n: message_class(n) for n in
     [r[0] + s for r in RPCS for s in ("Request", "Response")] + ["Embedding"]}

_CODES = {"NOT_FOUND": grpc.StatusCode.NOT_FOUND, "INVALID_ARGUMENT": grpc.StatusCode.INVALID_ARGUMENT,
          "FAILED_PRECONDITION": grpc.StatusCode.FAILED_PRECONDITION, "ALREADY_EXISTS": grpc.StatusCode.ALREADY_EXISTS}


class _NNBatcher:
    """Coalesces concurrent embedding-mode NearestNeighbor calls on one space into one batched search."""

    def __init__(self, hub, space, max_batch=256, max_wait_s=0.0005):
        self.hub, self.space, self.max_batch, self.max_wait_s = hub, space, max_batch, max_wait_s
        self.q = queue.Queue()
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def submit(self, num, embedding):
        slot = {"num": num, "emb": embedding, "ev": threading.Event(), "res": None, "err": None}
        self.q.put(slot)
        slot["ev"].wait()
        if slot["err"] is not None:
            raise slot["err"]
        return slot["res"]

    def _run(self):
        while True:
            first = self.q.get()
            batch = [first]
            deadline = time.perf_counter() + self.max_wait_s
            while len(batch) < self.max_batch:
                left = deadline - time.perf_counter()
                try:
                    batch.append(self.q.get(timeout=max(left, 0)) if left > 0 else self.q.get_nowait())
                except queue.Empty:
                    break
            try:
                kmax = max(s["num"] for s in batch)
                res = self.hub.multi_nearest_neighbor(self.space, kmax, embeddings=[s["emb"] for s in batch])
                for s, r in zip(batch, res):
                    s["res"] = r[: s["num"]]
            except Exception as e:  # noqa: BLE001 - forwarded to every waiter
                for s in batch:
                    s["err"] = e
            for s in batch:
                s["ev"].set()


class EmbeddingHubServicer:
    def __init__(self, hub=None, device=0, batch=True):
        self.hub = hub or EmbeddingHub(device=device)
        self.batch = batch
        self._batchers = {}
        self._lock = threading.Lock()

    def _fail(self, context, e):
        context.abort(_CODES.get(e.code, grpc.StatusCode.UNKNOWN), e.message)

    # ---- unary RPCs --------------------------------------------------------------------------
    def CreateSpace(self, req, context):
        try:
            self.hub.create_space(req.name, req.dims)
        except HubError as e:
            self._fail(context, e)
        return M["CreateSpaceResponse"]()

    def DeleteSpace(self, req, context):
        self.hub.delete_space(req.name)
        return M["DeleteSpaceResponse"]()

    def FreezeSpace(self, req, context):
        try:
            self.hub.freeze_space(req.name)
        except HubError as e:
            self._fail(context, e)
        return M["FreezeSpaceResponse"]()

    def Set(self, req, context):
        try:
            self.hub.set(req.space, req.key, list(req.embedding.values))
        except HubError as e:
            self._fail(context, e)
        return M["SetResponse"]()

    def Get(self, req, context):
        try:
            vals = self.hub.get(req.space, req.key)
        except HubError as e:
            self._fail(context, e)
        return M["GetResponse"](embedding=M["Embedding"](values=vals))

    def MultiSet(self, req_iter, context):
        by_space = {}
        for r in req_iter:
            by_space.setdefault(r.space, []).append((r.key, list(r.embedding.values)))
        try:
            for space, items in by_space.items():
                self.hub.multiset(space, items)          # one batched ingest per space
        except HubError as e:
            self._fail(context, e)
        return M["MultiSetResponse"]()

    def MultiGet(self, req_iter, context):
        for r in req_iter:
            try:
                yield M["MultiGetResponse"](embedding=M["Embedding"](values=self.hub.get(r.space, r.key)))
            except HubError as e:
                self._fail(context, e)

    def NearestNeighbor(self, req, context):
        has_vec = len(req.embedding.values) != 0
        try:
            if self.batch and has_vec and not req.key:
                self.hub._space(req.space)  # NOT_FOUND before queueing
                with self._lock:
                    b = self._batchers.get(req.space)
                    if b is None:
                        b = self._batchers[req.space] = _NNBatcher(self.hub, req.space)
                keys = b.submit(req.num, list(req.embedding.values))
            else:
                keys = self.hub.nearest_neighbor(req.space, req.num, key=req.key,
                                                 embedding=list(req.embedding.values) if has_vec else None)
        except HubError as e:
            self._fail(context, e)
        return M["NearestNeighborResponse"](keys=keys)

    def Download(self, req, context):
        try:
            sp = self.hub._space(req.space)
        except HubError as e:
            self._fail(context, e)
        for key in sp.index.keys():
            yield M["DownloadResponse"](key=key, embedding=M["Embedding"](values=sp.index.get(key).tolist()))


def _handlers(servicer):
    h = {}
    for name, cs, ss in RPCS:
        fn = getattr(servicer, name)
        de, se = M[name + "Request"].FromString, M[name + "Response"].SerializeToString
        if cs and ss:
            h[name] = grpc.stream_stream_rpc_method_handler(fn, de, se)
        elif cs:
            h[name] = grpc.stream_unary_rpc_method_handler(fn, de, se)
        elif ss:
            h[name] = grpc.unary_stream_rpc_method_handler(fn, de, se)
        else:
            h[name] = grpc.unary_unary_rpc_method_handler(fn, de, se)
    return grpc.method_handlers_generic_handler(SERVICE, h)


def make_server(address=DEFAULT_ADDRESS, device=0, max_workers=32, hub=None):
    """Returns (grpc server, bound port).  RunServer of embeddingstore/server.cc:249-268."""
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    server.add_generic_rpc_handlers((_handlers(EmbeddingHubServicer(hub=hub, device=device)),))
    port = server.add_insecure_port(address)
    return server, port


class Stub:
    """Client stub over the same descriptors (what protoc would generate as EmbeddingHubStub)."""

    def __init__(self, channel):
        for name, cs, ss in RPCS:
            path = f"/{SERVICE}/{name}"
            se, de = M[name + "Request"].SerializeToString, M[name + "Response"].FromString
            kind = (channel.stream_stream if cs and ss else channel.stream_unary if cs else
                    channel.unary_stream if ss else channel.unary_unary)
            setattr(self, name, kind(path, request_serializer=se, response_deserializer=de))


def main(argv=None):
    import sys

    argv = argv or sys.argv
    address = argv[1] if len(argv) > 1 else DEFAULT_ADDRESS   # main.cc:8
    server, port = make_server(address)
    server.start()
    print(f"Server listening on {address} (port {port})", flush=True)  # server.cc:263
    server.wait_for_termination()


if __name__ == "__main__":
    main()
