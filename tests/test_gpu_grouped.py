"""GPU parity of the grouped (IVF) index: WordVectors#grouped / Grouped#residuals
(G/WordVectors.scala:24-58,118-138), Index.grouped (G/Index.scala:133-145) and GroupedIndex#query with
both search-space strategies (G/Index.scala:266-299) against the oracle's host restatement.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def clustered(rng, n, d, centres=12, scale=3.0, noise=0.4):
    c = rng.normal(size=(centres, d)).astype(np.float32) * scale
    x = c[rng.integers(0, centres, n)] + rng.normal(size=(n, d)).astype(np.float32) * noise
    return np.ascontiguousarray(x, np.float32)


def build(g, oracle, rng, n=6000, D=24, P=12, M=4, K=64, keys=False, normalized=False):
    X = clustered(rng, n, D, centres=P)
    if normalized:
        X = g.normalize(X)
    coarse = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X)), g.KMeansConfig(P, 6, seed=1))
    ks = ["w%05d" % int(v) for v in rng.permutation(n)] if keys else None
    gv = g.GroupedVectors.group(g.Matrix(X), coarse, keys=ks)
    res = gv.residuals_dev()
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(res), g.ProductQuantizerConfig(K, M, 4))
    ix = g.GroupedIndex.build(gv, pq, normalized=normalized, strategy=g.LimitGroups(3))
    return X, coarse, ks, gv, res, pq, ix


def test_grouping_matches_reference_literally(g, oracle):
    rng = np.random.default_rng(1)
    X, coarse, ks, gv, res, pq, ix = build(g, oracle, rng, keys=True)
    a = oracle.assign(X, 0, X.shape[1], coarse.centroids, batch=25000, tie_mode=oracle.TIE_LOWEST)
    order, cents, offsets = oracle.grouped_build(X, a, coarse.centroids, keys=ks)
    assert np.array_equal(gv.order, order)
    assert np.array_equal(gv.offsets, offsets)
    assert np.array_equal(gv.centroids.view(np.uint32), cents.view(np.uint32))
    assert gv.centroids.shape[0] == len(gv.offsets) + 1
    # keys ascend inside every group (the stable double sort)
    for i in range(gv.centroids.shape[0]):
        f, u = gv.bounds(i)
        assert gv.keys[f:u] == sorted(gv.keys[f:u])
    # the reference's seed quirk: unless row 0 is in the lowest cluster, group 0 is empty
    if a[0] != a.min():
        assert gv.bounds(0) == (0, 0) and offsets[0] == 0
    want = oracle.grouped_residuals(X, order, cents, offsets)
    assert np.array_equal(res.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert [gv.cluster_of(i) for i in (0, 17, gv.size - 1)] == \
        [int(np.searchsorted(offsets, i, side="right")) for i in (0, 17, gv.size - 1)]


@pytest.mark.parametrize("strategy", [("groups", 1), ("groups", 3), ("groups", 50), ("vectors", 1),
                                      ("vectors", 1500), ("vectors", 10 ** 9), ("groups", 0)])
@pytest.mark.parametrize("normalized", [False, True])
@pytest.mark.parametrize("work_list", [True, False])
def test_grouped_query_matches_oracle(g, oracle, strategy, normalized, work_list):
    """work_list: gulon_grouped_query_dev, one launch for every (query, partition) pair of the batch;
    otherwise one ranged PQIndex#batchQuery per probed partition.  Both equal the oracle."""
    rng = np.random.default_rng(7)
    X, coarse, ks, gv, res, pq, ix = build(g, oracle, rng, normalized=normalized)
    ix.work_list = work_list
    ix.strategy = g.LimitGroups(strategy[1]) if strategy[0] == "groups" else g.LimitVectors(strategy[1])
    Q = np.concatenate((X[rng.integers(0, X.shape[0], 20)] + 0.01,
                        clustered(rng, 21, X.shape[1]))).astype(np.float32)
    k = 10
    got = ix.batch_query(k, Q)
    codes = oracle.pq_encode(res.cpu().numpy(), pq.codebook(), tie_mode=oracle.TIE_LOWEST)
    wi, wd, ws = oracle.grouped_query(Q, gv.centroids, gv.offsets, gv.size, pq.codebook(), codes, k,
                                      strategy, normalized=normalized)
    assert np.array_equal(got.size, ws)
    assert np.array_equal(got.values.view(np.uint32), wd.view(np.uint32))
    assert np.array_equal(got.keys, wi)
    # single-query form and the mapping back to the rows the index was built from
    one = ix.query(k, Q[3])
    assert np.array_equal(one[0], wi[3, :ws[3]])
    rows = ix.original_rows(got.keys)
    assert np.array_equal(rows[got.keys >= 0], gv.order[got.keys[got.keys >= 0]])


def test_grouped_lookup_and_query_by_position(g, oracle):
    rng = np.random.default_rng(11)
    X, coarse, ks, gv, res, pq, ix = build(g, oracle, rng, n=3000, K=256)
    ix.strategy = g.LimitGroups(4)
    codes = oracle.pq_encode(res.cpu().numpy(), pq.codebook(), tie_mode=oracle.TIE_LOWEST)
    dec = oracle.pq_decode(codes, pq.codebook(), X.shape[1])
    for pos in (0, 1234, gv.size - 1):
        want = gv.centroids[gv.cluster_of(pos)] + dec[pos]          # MathUtils.add, G/Index.scala:247-254
        got = ix.lookup(pos)
        assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
        keys, dist = ix.query_by_position(3, pos)                    # queryByWord finds the word
        assert pos in keys


def test_grouped_index_large_batch_recall_vs_full_scan(g, oracle):
    # 200k rows, 64 partitions: probing more groups can only improve the distances, and probing all
    # groups reproduces the per-partition exhaustive answer
    rng = np.random.default_rng(13)
    n, D, P = 200000, 32, 64
    X = clustered(rng, n, D, centres=P, noise=0.8)
    coarse = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X[:50000])), g.KMeansConfig(P, 5, seed=2))
    gv = g.GroupedVectors.group(g.Matrix(X), coarse)
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(gv.residuals_dev()[:50000].contiguous()),
                                  g.ProductQuantizerConfig(256, 8, 4))
    ix = g.GroupedIndex.build(gv, pq, strategy=g.LimitGroups(2))
    Q = X[rng.integers(0, n, 500)] + rng.normal(size=(500, D)).astype(np.float32) * 0.05
    r2 = ix.batch_query(10, Q)
    ix.strategy = g.LimitGroups(8)
    r8 = ix.batch_query(10, Q)
    assert np.all(r8.values[:, 0] <= r2.values[:, 0])
    assert np.all(r8.size == 10)
    # partitions of ~3000 rows (several sort rounds per pair) and k beyond one warp's list: the work-list
    # kernel against the ranged scans, bit for bit
    for k in (10, 200):
        a = ix.batch_query(k, Q[:64])
        ix.work_list = False
        try:
            b = ix.batch_query(k, Q[:64])
        finally:
            ix.work_list = True
        assert np.array_equal(a.keys, b.keys) and np.array_equal(a.values.view(np.uint32), b.values.view(np.uint32))
        assert np.array_equal(a.size, b.size)
    # a query that is a stored row finds its own position among the nearest
    pos = rng.integers(0, n, 50)
    own = ix.batch_query(5, gv.matrix_dev[pos].cpu().numpy())
    hit = [(p in own.keys[i]) or np.isclose(own.values[i, 0], own.values[i, :].min()) for i, p in enumerate(pos)]
    assert all(hit)
