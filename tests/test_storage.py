"""On-disk index format (core/src/main/protobuf/index.proto): the hand-written proto2 codec in
gulon_b200/storage.py against google.protobuf with the same schema (built at run time from a
FileDescriptorProto restating index.proto -- no generated code, no protoc).  CPU only."""
import numpy as np
import pytest

from gulon_b200 import storage as S

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def _schema():
    """index.proto restated field by field (P/index.proto:12-68), scalapb options left out (they do
    not change the wire format)."""
    fd = descriptor_pb2.FileDescriptorProto(name="index_test.proto", package="gulon", syntax="proto2")

    def msg(parent, name):
        m = parent.message_type.add() if hasattr(parent, "message_type") else parent.nested_type.add()
        m.name = name
        return m

    def field(m, name, number, type_, label, type_name=None):
        f = m.field.add(name=name, number=number, type=type_, label=label)
        if type_name:
            f.type_name = type_name
        return f

    REQ, REP, OPT = F.LABEL_REQUIRED, F.LABEL_REPEATED, F.LABEL_OPTIONAL
    fv = msg(fd, "FloatVector")
    field(fv, "values", 1, F.TYPE_FLOAT, REP)
    pq = msg(fd, "ProductQuantizer")
    field(pq, "num_clusters", 1, F.TYPE_INT32, REQ)
    field(pq, "quantizers", 2, F.TYPE_MESSAGE, REP, ".gulon.ProductQuantizer.Quantizer")
    qz = msg(pq, "Quantizer")
    field(qz, "start_index", 1, F.TYPE_INT32, REQ)
    field(qz, "dimension", 2, F.TYPE_INT32, REQ)
    field(qz, "centroids", 3, F.TYPE_MESSAGE, REP, ".gulon.FloatVector")
    em = msg(fd, "EncodedMatrix")
    field(em, "code_width", 1, F.TYPE_INT32, REQ)
    field(em, "length", 2, F.TYPE_INT32, REQ)
    field(em, "encodings", 3, F.TYPE_BYTES, REP)
    me = fd.enum_type.add(name="Metric")
    me.value.add(name="L2", number=0)
    me.value.add(name="COSINE", number=1)
    pi = msg(fd, "PQIndex")
    field(pi, "product_quantizer", 1, F.TYPE_MESSAGE, REQ, ".gulon.ProductQuantizer")
    field(pi, "data", 2, F.TYPE_MESSAGE, REQ, ".gulon.EncodedMatrix")
    si = msg(fd, "SortedIndex")
    field(si, "sorted_words", 1, F.TYPE_STRING, REP)
    field(si, "vector_index", 2, F.TYPE_MESSAGE, REQ, ".gulon.PQIndex")
    field(si, "metric", 3, F.TYPE_ENUM, REQ, ".gulon.Metric")
    gi = msg(fd, "GroupedIndex")
    field(gi, "grouped_words", 1, F.TYPE_STRING, REP)
    field(gi, "vector_index", 2, F.TYPE_MESSAGE, REQ, ".gulon.PQIndex")
    field(gi, "metric", 3, F.TYPE_ENUM, REQ, ".gulon.Metric")
    field(gi, "centroids", 4, F.TYPE_MESSAGE, REP, ".gulon.FloatVector")
    field(gi, "offsets", 5, F.TYPE_INT32, REP)
    se = gi.enum_type.add(name="Strategy")
    se.value.add(name="LIMIT_GROUPS", number=0)
    se.value.add(name="LIMIT_VECTORS", number=2)
    field(gi, "strategy", 6, F.TYPE_ENUM, REQ, ".gulon.GroupedIndex.Strategy")
    field(gi, "limit", 7, F.TYPE_INT32, REQ)
    ix = msg(fd, "Index")
    ix.oneof_decl.add(name="implementation")
    field(ix, "sorted", 1, F.TYPE_MESSAGE, OPT, ".gulon.SortedIndex").oneof_index = 0
    field(ix, "grouped", 2, F.TYPE_MESSAGE, OPT, ".gulon.GroupedIndex").oneof_index = 0
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("gulon.Index"))


@pytest.fixture(scope="module")
def IndexMsg():
    return _schema()


def _example(kind, rng, n=37, D=7, M=3, K=256, P=4):
    dims = [3, 2, 2]
    quantizers, at = [], 0
    for d in dims[:M]:
        quantizers.append({"start_index": at, "dimension": d,
                           "centroids": rng.standard_normal((K, d)).astype(np.float32)})
        at += d
    vi = {"product_quantizer": {"num_clusters": K, "quantizers": quantizers},
          "data": {"code_width": 8, "length": n,
                   "encodings": [rng.integers(0, 256, n).astype(np.uint8) for _ in range(M)]}}
    words = sorted("wé%04d" % i for i in range(n))
    ix = {"kind": kind, "words": words, "vector_index": vi, "metric": S.COSINE}
    if kind == "grouped":
        ix["centroids"] = rng.standard_normal((P, D)).astype(np.float32)
        ix["offsets"] = np.array([0, 9, 30], np.int32)[:P - 1]
        ix["strategy"] = S.LIMIT_VECTORS
        ix["limit"] = 1234567
    return ix


def _fill(msg, ix):
    m = msg.sorted if ix["kind"] == "sorted" else msg.grouped
    (m.sorted_words if ix["kind"] == "sorted" else m.grouped_words).extend(ix["words"])
    m.metric = ix["metric"]
    pq = m.vector_index.product_quantizer
    pq.num_clusters = ix["vector_index"]["product_quantizer"]["num_clusters"]
    for q in ix["vector_index"]["product_quantizer"]["quantizers"]:
        qq = pq.quantizers.add(start_index=q["start_index"], dimension=q["dimension"])
        for c in q["centroids"]:
            qq.centroids.add().values.extend(float(v) for v in c)
    d = m.vector_index.data
    d.code_width = ix["vector_index"]["data"]["code_width"]
    d.length = ix["vector_index"]["data"]["length"]
    d.encodings.extend(p.tobytes() for p in ix["vector_index"]["data"]["encodings"])
    if ix["kind"] == "grouped":
        for c in ix["centroids"]:
            m.centroids.add().values.extend(float(v) for v in c)
        m.offsets.extend(int(o) for o in ix["offsets"])
        m.strategy = ix["strategy"]
        m.limit = ix["limit"]


def _same(a, b):
    assert a["kind"] == b["kind"] and a["words"] == b["words"] and a["metric"] == b["metric"]
    pa, pb_ = a["vector_index"]["product_quantizer"], b["vector_index"]["product_quantizer"]
    assert pa["num_clusters"] == pb_["num_clusters"] and len(pa["quantizers"]) == len(pb_["quantizers"])
    for x, y in zip(pa["quantizers"], pb_["quantizers"]):
        assert x["start_index"] == y["start_index"] and x["dimension"] == y["dimension"]
        assert np.array_equal(np.asarray(x["centroids"]).view(np.uint32), np.asarray(y["centroids"]).view(np.uint32))
    da, db = a["vector_index"]["data"], b["vector_index"]["data"]
    assert da["code_width"] == db["code_width"] and da["length"] == db["length"]
    assert all(np.array_equal(x, y) for x, y in zip(da["encodings"], db["encodings"]))
    if a["kind"] == "grouped":
        assert np.array_equal(a["centroids"].view(np.uint32), b["centroids"].view(np.uint32))
        assert np.array_equal(a["offsets"], b["offsets"])
        assert a["strategy"] == b["strategy"] and a["limit"] == b["limit"]


@pytest.mark.parametrize("kind", ["sorted", "grouped"])
def test_bytes_equal_protobuf_serialization(IndexMsg, kind):
    """scalapb and google.protobuf both write fields in field-number order with unpacked proto2
    repeated scalars: the codec's bytes equal the library's for the same message."""
    ix = _example(kind, np.random.default_rng(1))
    msg = IndexMsg()
    _fill(msg, ix)
    assert S.encode_index(ix) == msg.SerializeToString()


@pytest.mark.parametrize("kind", ["sorted", "grouped"])
def test_decode_what_protobuf_wrote(IndexMsg, kind):
    ix = _example(kind, np.random.default_rng(2))
    msg = IndexMsg()
    _fill(msg, ix)
    _same(S.decode_index(msg.SerializeToString()), ix)


@pytest.mark.parametrize("kind", ["sorted", "grouped"])
def test_protobuf_parses_what_the_codec_wrote(IndexMsg, kind):
    ix = _example(kind, np.random.default_rng(3))
    msg = IndexMsg()
    msg.ParseFromString(S.encode_index(ix))
    ref = IndexMsg()
    _fill(ref, ix)
    assert msg == ref and msg.WhichOneof("implementation") == kind


def test_negative_int32_and_packed_offsets():
    ix = _example("grouped", np.random.default_rng(4))
    ix["limit"] = -5                        # int32 negatives are 10-byte varints
    out = S.decode_index(S.encode_index(ix))
    assert out["limit"] == -5
    # a proto3-style writer may pack `repeated int32 offsets`; parsers must accept both forms
    body = S.encode_index(ix)
    packed = S._f_bytes(5, b"".join(S._varint(int(o)) for o in ix["offsets"]))
    unpacked = b"".join(S._f_varint(5, int(o)) for o in ix["offsets"])
    r = S._Reader(body)
    f, w = r.field()
    inner = bytes(r.bytes_())
    assert f == 2 and unpacked in inner
    out2 = S.decode_index(S._f_bytes(2, inner.replace(unpacked, packed)))
    assert np.array_equal(out2["offsets"], ix["offsets"])


def test_empty_index_and_unknown_fields():
    ix = {"kind": "sorted", "words": [], "metric": S.L2,
          "vector_index": {"product_quantizer": {"num_clusters": 256, "quantizers": []},
                           "data": {"code_width": 8, "length": 0, "encodings": []}}}
    raw = S.encode_index(ix)
    _same(S.decode_index(raw), ix)
    # unknown fields are skipped (forward compatibility)
    _same(S.decode_index(S._f_varint(9, 7) + raw + S._f_bytes(12, b"xyz")), ix)


def test_errors_follow_the_reference():
    """Index.fromProtobuf, G/Index.scala:175-207: empty oneof, unrecognised strategy / metric."""
    with pytest.raises(ValueError, match="missing index implementation"):
        S.decode_index(b"")
    ix = _example("grouped", np.random.default_rng(5))
    ix["strategy"] = 1
    with pytest.raises(ValueError, match="LIMIT_GROUPS or LIMIT_VECTORS"):
        S.decode_index(S.encode_index(ix))
    ix = _example("sorted", np.random.default_rng(6))
    ix["metric"] = 3
    with pytest.raises(ValueError, match="metric"):
        S.decode_index(S.encode_index(ix))
    with pytest.raises(ValueError):
        S.decode_index(S.encode_index(_example("sorted", np.random.default_rng(7)))[:-3])
