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
