"""CPU check of the gRPC wire contract: descriptors match the reference proto's names, field numbers and
streaming modes (embeddinghub/embeddingstore/embedding_store.proto:9-106), and known bytes round-trip."""
from embeddinghub_b200 import grpc_server as gs


def test_service_shape():
    names = {r[0]: (r[1], r[2]) for r in gs.RPCS}
    assert names == {"CreateSpace": (False, False), "DeleteSpace": (False, False), "FreezeSpace": (False, False),
                     "Set": (False, False), "Get": (False, False), "MultiSet": (True, False),
                     "MultiGet": (True, True), "NearestNeighbor": (False, False), "Download": (False, True)}
    assert gs.SERVICE == "featureform.embedding.proto.EmbeddingHub"


def test_field_numbers():
    def fields(n):
        return {f.name: f.number for f in gs.M[n].DESCRIPTOR.fields}

    assert fields("NearestNeighborRequest") == {"num": 1, "space": 2, "key": 3, "embedding": 4}
    assert fields("NearestNeighborResponse") == {"keys": 1}
    assert fields("SetRequest") == {"key": 1, "embedding": 2, "space": 3}
    assert fields("GetRequest") == {"key": 1, "space": 2}
    assert fields("CreateSpaceRequest") == {"name": 1, "dims": 2}
    assert fields("DownloadResponse") == {"key": 1, "embedding": 2}
    assert fields("Embedding") == {"values": 1}


def test_wire_bytes():
    # NearestNeighborRequest{num=2, space="s", embedding{values=[1.0]}} hand-encoded per the proto3 wire format
    raw = bytes([0x08, 0x02, 0x12, 0x01, 0x73, 0x22, 0x06, 0x0A, 0x04, 0x00, 0x00, 0x80, 0x3F])
    m = gs.M["NearestNeighborRequest"].FromString(raw)
    assert (m.num, m.space, m.key, list(m.embedding.values)) == (2, "s", "", [1.0])
    assert m.SerializeToString() == raw
