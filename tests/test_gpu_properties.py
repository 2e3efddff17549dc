"""GPU: the reference's own ScalaCheck properties (T/KMeansSpec.scala, T/ProductQuantizerSpec.scala,
T/IndexSpec.scala) through the library's public API -- the same statements tests/test_oracle_properties.py
makes about the oracle, here about the CUDA path (ties broken by lowest index instead of java.util.Random)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from test_oracle_properties import gen_vectors, gen_pq

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def objective(g, oracle, X, km):
    a = km.assign(g.Vectors(g.Matrix(X)))
    return float(np.sum([oracle.distance_sq(X[i], km.centroids[a[i]]) for i in range(len(X))], dtype=np.float64))


@settings(max_examples=25, deadline=None, derandomize=True)
@given(gen_vectors())
def test_compute_clusters_converges(g, gv):
    X, cents = gv
    km, info = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X)), g.KMeansConfig(len(cents), 100, seed=0),
                                         return_info=True)
    assert info["converged"] and km.k == len(cents)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(gen_vectors())
def test_iterate_progresses_towards_minimum(g, oracle, gv):
    X, cents = gv
    v = g.Vectors(g.Matrix(X))
    km = g.KMeans.init(len(cents), v)
    prev = objective(g, oracle, X, km)
    for iters in (1, 3, 7):
        km = km.iterate(v, iters)
        cur = objective(g, oracle, X, km)
        assert cur <= prev * (1 + 1e-5) + 1e-6
        prev = cur


@settings(max_examples=20, deadline=None, derandomize=True)
@given(gen_vectors())
def test_does_not_get_stuck_when_clusters_are_not_distinct(g, oracle, gv):
    X, cents = gv
    v = g.Vectors(g.Matrix(X))
    a0 = np.zeros(len(X), np.int32)
    k0 = g.KMeans.from_assignment(len(cents), X.shape[1], v, a0)
    k1 = k0.iterate(v, 1)
    if not np.array_equal(a0, k1.assign(v)):
        assert objective(g, oracle, X, k0) >= objective(g, oracle, X, k1) * (1 - 1e-6)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(gen_pq(), st.integers(1, 20))
def test_decode_encode_is_idempotent_and_decode_selects_centroids(g, pq, n):
    D, M, K, cb, rng = pq
    q = g.ProductQuantizer.from_codebook(cb, D)
    X = rng.uniform(-6, 6, (n, D)).astype(f32)
    d1 = q.decode(q.encode(X))
    d2 = q.decode(q.encode(d1))
    assert np.allclose(d1.data, d2.data, rtol=1e-3, atol=1e-3)
    all_codes = g.EncodedMatrix.from_planes(q.coder_factory(K), np.tile(np.arange(K, dtype=np.uint8), (M, 1)))
    dec = q.decode(all_codes).data
    for m, qz in enumerate(q.quantizers):
        assert np.array_equal(dec[:, qz.from_:qz.from_ + qz.dimension], cb[m, :, :qz.dimension])


@settings(max_examples=25, deadline=None, derandomize=True)
@given(gen_pq(), st.integers(1, 12))
def test_encode_selects_closest_encoding(g, oracle, pq, n_rand):
    D, M, K, cb, rng = pq
    q = g.ProductQuantizer.from_codebook(cb, D)
    p = rng.uniform(-6, 6, (1, D)).astype(f32)
    p0 = q.decode(q.encode(p)).data[0]
    d = np.sqrt(oracle.distance_sq(p[0], p0))
    rand = g.EncodedMatrix.from_planes(q.coder_factory(n_rand), rng.integers(0, K, (M, n_rand)).astype(np.uint8))
    for r in q.decode(rand).data:
        assert d <= np.sqrt(oracle.distance_sq(p[0], r)) * (1 + 1e-4) + 1e-4


@settings(max_examples=25, deadline=None, derandomize=True)
@given(gen_pq(), st.integers(5, 60), st.integers(1, 8))
def test_sorted_index_queries_encoded_nearest_neighbours(g, pq, n, k):
    from gulon_b200.storage import SortedIndex
    D, M, K, cb, rng = pq
    q = g.ProductQuantizer.from_codebook(cb, D)
    X = rng.uniform(-6, 6, (n, D)).astype(f32)
    ix = SortedIndex.build(["w%03d" % i for i in range(n)], X, q)
    dec = q.decode(ix.vector_index.data).data
    query = rng.uniform(-6, 6, (1, D)).astype(f32)
    k = min(k, n)
    ids, ds = ix.query(k, query[0])
    ex = g.exact_nearest_neighbours(g.Matrix(dec), query, k + 1 if k < n else k)
    ed = ex.values[0]
    if k < n and ed[k] - ed[k - 1] <= 1e-4 * max(1.0, ed[k]):
        return
    assert sorted(ids.tolist()) == sorted(ex.keys[0, :k].tolist())
    assert np.allclose(ds, ed[:k], rtol=1e-4, atol=1e-4)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(gen_pq(), st.integers(1, 50))
def test_query_by_word_finds_word(g, pq, n):
    from gulon_b200.storage import SortedIndex
    D, M, K, cb, rng = pq
    q = g.ProductQuantizer.from_codebook(cb, D)
    X = rng.uniform(-6, 6, (n, D)).astype(f32)
    words = ["w%03d" % i for i in range(n)]
    ix = SortedIndex.build(words, X, q)
    dec = q.decode(ix.vector_index.data).data
    _, counts = np.unique(dec, axis=0, return_counts=True)
    k = int(counts.max()) + 1
    i = int(rng.integers(0, n))
    ids, ds = ix.query_by_word(k, words[i])
    assert i in ids.tolist() and ds[0] == 0.0
