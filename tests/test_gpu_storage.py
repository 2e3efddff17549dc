"""GPU: indexes written in the reference's on-disk format (P/index.proto) load straight into HBM and
answer queries bit-identically to the index they were written from (Index.read / toProtobuf /
fromProtobuf, G/Index.scala:147-207)."""
import numpy as np
import pytest

from test_gpu_grouped import build, clustered

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


@pytest.mark.parametrize("normalized", [False, True])
def test_sorted_index_round_trip(g, oracle, tmp_path, normalized):
    from gulon_b200 import storage
    rng = np.random.default_rng(11)
    n, D, M, K = 5000, 22, 4, 256                       # 22 = 6 + 6 + 5 + 5: ragged windows
    X = clustered(rng, n, D)
    if normalized:
        X = g.normalize(X)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(K, M, 3))
    enc = pq.encode(g.Matrix(X))
    words = ["w%05d" % i for i in range(n)]
    ix = storage.SortedIndex(words, g.PQIndex(pq, enc), normalized)
    path = tmp_path / "sorted.index"
    storage.write(ix, path)
    back = storage.read(path)
    assert isinstance(back, storage.SortedIndex) and back.words == words and back.normalized == normalized
    assert np.array_equal(back.vector_index.product_quantizer.codebook().view(np.uint32),
                          pq.codebook().view(np.uint32))
    assert np.array_equal(back.vector_index.data.codes, enc.codes)
    Q = clustered(rng, 33, D)
    a, b = ix.batch_query(10, Q), back.batch_query(10, Q)
    assert np.array_equal(a.keys, b.keys) and np.array_equal(a.values.view(np.uint32), b.values.view(np.uint32))
    # and both equal the oracle on the codes that were stored
    qn = oracle.normalize(Q) if normalized else Q
    wi, wd, ws = oracle.pq_query(qn, pq.codebook(), enc.codes, 10)
    assert np.array_equal(b.keys, wi) and np.array_equal(b.values.view(np.uint32), wd.view(np.uint32))
    # lookup by word = decode of the stored row (SortedIndex#lookup, G/Index.scala:318-322)
    assert np.array_equal(back.lookup("w00017"), ix.vector_index.decode(17))
    assert back.lookup("absent") is None
    # the bytes are a fixed point
    assert storage.to_protobuf(back) == path.read_bytes()


@pytest.mark.parametrize("strategy", [("groups", 3), ("vectors", 1500)])
def test_grouped_index_round_trip(g, oracle, tmp_path, strategy):
    from gulon_b200 import storage
    rng = np.random.default_rng(12)
    X, coarse, ks, gv, res, pq, ix = build(g, oracle, rng, keys=True)
    ix.strategy = g.LimitGroups(strategy[1]) if strategy[0] == "groups" else g.LimitVectors(strategy[1])
    raw = storage.to_protobuf(ix)
    back = storage.from_protobuf(raw)
    assert back.strategy == ix.strategy and back.normalized == ix.normalized
    assert back.grouped.keys == gv.keys
    assert np.array_equal(back.grouped.offsets, gv.offsets)
    assert np.array_equal(back.grouped.centroids.view(np.uint32), gv.centroids.view(np.uint32))
    Q = np.concatenate((X[rng.integers(0, X.shape[0], 10)] + 0.01, clustered(rng, 9, X.shape[1])))
    a, b = ix.batch_query(10, Q.astype(np.float32)), back.batch_query(10, Q.astype(np.float32))
    assert np.array_equal(a.size, b.size) and np.array_equal(a.keys, b.keys)
    assert np.array_equal(a.values.view(np.uint32), b.values.view(np.uint32))
    assert np.array_equal(back.lookup(123).view(np.uint32), ix.lookup(123).view(np.uint32))
    assert storage.to_protobuf(back) == raw


@pytest.mark.parametrize("K", [1, 4, 16, 100])
def test_small_codebooks_use_the_packed_coders(g, oracle, tmp_path, K):
    """numClusters <= 16 selects Coder0 / Coder2 / Coder4 (G/ProductQuantizer.scala:11-16,
    G/Coder.scala:35-45): the packed planes are what the file holds; the device works on one byte per id."""
    from gulon_b200 import storage
    from gulon_b200.coder import factory_for, max_width
    rng = np.random.default_rng(100 + K)
    n, D, M = 3001, 20, 4
    X = clustered(rng, n, D)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(K, M, 3))
    enc = pq.encode(g.Matrix(X))
    width = factory_for(max_width(K)).width
    assert enc.coder.width == width == {1: 0, 4: 2, 16: 4, 100: 8}[K]
    want_codes = oracle.pq_encode(X, pq.codebook(), tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(enc.codes, want_codes)
    assert all(len(p) == (n * width + 7) // 8 for p in enc.unwrapped_encodings)
    ix = storage.SortedIndex(["w%05d" % i for i in range(n)], g.PQIndex(pq, enc), False)
    raw = storage.to_protobuf(ix)
    d = storage.decode_index(raw)
    assert d["vector_index"]["data"]["code_width"] == width
    assert all(len(p) == (n * width + 7) // 8 for p in d["vector_index"]["data"]["encodings"])
    back = storage.from_protobuf(raw)
    assert back.vector_index.data == enc
    Q = clustered(rng, 9, D)
    a, b = ix.batch_query(7, Q), back.batch_query(7, Q)
    wi, wd, ws = oracle.pq_query(Q, pq.codebook(), enc.codes, 7)
    for r in (a, b):
        assert np.array_equal(r.keys, wi) and np.array_equal(r.values.view(np.uint32), wd.view(np.uint32))
    assert storage.to_protobuf(back) == raw


def test_sorted_index_builder_and_word_queries(g, oracle):
    """Index.sorted (G/Index.scala:107-113), queryByWord (:44-45) and Index.Result (:62-94)."""
    from gulon_b200 import storage
    rng = np.random.default_rng(13)
    n, D, M = 3000, 16, 4
    X = clustered(rng, n, D)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(256, M, 3))
    words = ["w%05d" % i for i in range(n)]
    ix = storage.SortedIndex.build(words, X, pq)
    assert ix.size == n and ix.dimension == D
    assert np.array_equal(ix.vector_index.data.codes, oracle.pq_encode(X, pq.codebook(), tie_mode=oracle.TIE_LOWEST))
    with pytest.raises(ValueError, match="sorted"):
        storage.SortedIndex.build(words[::-1], X, pq)
    with pytest.raises(ValueError, match="one word per row"):
        storage.SortedIndex.build(words[:-1], X, pq)
    r = ix.query_by_word(5, "w00123")
    # the decoded row is at ADC distance 0 from its own code; the lowest id among identical codes leads
    assert r[1][0] == 0.0 and r[0][0] <= 123
    assert np.array_equal(ix.vector_index.data.codes[:, r[0][0]], ix.vector_index.data.codes[:, 123])
    assert ix.query_by_word(5, "nope") is None
    Q = clustered(rng, 4, D)
    res = ix.results(3, Q)
    wi, wd, ws = oracle.pq_query(Q, pq.codebook(), ix.vector_index.data.codes, 3)
    assert [[w for w, _ in row] for row in res] == [[words[i] for i in row] for row in wi.tolist()]
    assert np.array_equal(np.array([[d for _, d in row] for row in res], np.float32).view(np.uint32), wd.view(np.uint32))
