"""gRPC front end speaking the reference's wire contract.

Service `featureform.embedding.proto.EmbeddingHub` with its nine RPCs and message field numbers
exactly as declared in embeddinghub/embeddingstore/embedding_store.proto:9-106, so the reference's
Python SDK (embeddinghub/sdk/python/embeddinghub.py) can talk to this server unchanged.  The .proto
file itself is not copied: the descriptors are restated programmatically below (grpcio is present in
this image, protoc-generated stubs are not usable with its protobuf runtime).

Request handling delegates to embeddinghub_b200.hub.EmbeddingHub, which mirrors the semantics and
status codes of EmbeddingHubService (embeddingstore/server.cc:65-233).  The reference serialises
every RPC under one mutex (server.cc:175) and answers one query per call; here RPC threads call the
library concurrently and its combining queue (csrc/api.cu) coalesces their single-query searches into
batched launches.  The batched call the reference's docs promise (docs/inference.md:14-22) is served
by a side service, EmbeddingHubBatch.MultiNearestNeighbor, so the reference proto stays untouched.
"""
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


class EmbeddingHubServicer:
    """Thread-per-RPC like the reference's sync server, but WITHOUT its global mutex (server.cc:175): every
    NearestNeighbor call goes straight into ehb_index_search, whose combining queue turns the concurrent
    single-query calls of many RPC threads into batched launches.  One malformed request fails alone."""

    def __init__(self, hub=None, device=0):
        self.hub = hub or EmbeddingHub(device=device)

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
            keys = self.hub.nearest_neighbor(req.space, req.num, key=req.key,
                                             embedding=list(req.embedding.values) if has_vec else None)
        except HubError as e:
            self._fail(context, e)
        return M["NearestNeighborResponse"](keys=keys)

    def MultiNearestNeighbor(self, req_iter, context):
        """Side service (EmbeddingHubBatch): docs/inference.md:14-22 promises multi_nearest_neighbor but
        embedding_store.proto has no RPC for it and must stay byte-for-byte, so the batched form lives beside it:
        a stream of NearestNeighborRequest in, the matching NearestNeighborResponse stream out (same order), all
        embedding-mode requests of one (space, num) answered by ONE batched search."""
        reqs = list(req_iter)
        out = [None] * len(reqs)
        groups = {}
        for i, r in enumerate(reqs):
            if len(r.embedding.values) != 0 and not r.key:
                groups.setdefault((r.space, r.num), []).append(i)
        try:
            for (space, num), idx in groups.items():
                res = self.hub.multi_nearest_neighbor(space, num, embeddings=[list(reqs[i].embedding.values) for i in idx])
                for i, keys in zip(idx, res):
                    out[i] = keys
            for i, r in enumerate(reqs):
                if out[i] is None:
                    has_vec = len(r.embedding.values) != 0
                    out[i] = self.hub.nearest_neighbor(r.space, r.num, key=r.key,
                                                       embedding=list(r.embedding.values) if has_vec else None)
        except HubError as e:
            self._fail(context, e)
        for keys in out:
            yield M["NearestNeighborResponse"](keys=keys)

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


BATCH_SERVICE = PKG + ".EmbeddingHubBatch"


def _batch_handlers(servicer):
    h = {"MultiNearestNeighbor": grpc.stream_stream_rpc_method_handler(
        servicer.MultiNearestNeighbor, M["NearestNeighborRequest"].FromString,
        M["NearestNeighborResponse"].SerializeToString)}
    return grpc.method_handlers_generic_handler(BATCH_SERVICE, h)


def make_server(address=DEFAULT_ADDRESS, device=0, max_workers=32, hub=None):
    """Returns (grpc server, bound port).  RunServer of embeddingstore/server.cc:249-268."""
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    servicer = EmbeddingHubServicer(hub=hub, device=device)
    server.add_generic_rpc_handlers((_handlers(servicer), _batch_handlers(servicer)))
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
        self.MultiNearestNeighbor = channel.stream_stream(
            f"/{BATCH_SERVICE}/MultiNearestNeighbor", request_serializer=M["NearestNeighborRequest"].SerializeToString,
            response_deserializer=M["NearestNeighborResponse"].FromString)


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
