"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for PQ codes, assignments, neighbour ids and -- because every kernel replays the
reference's fp32 operation order with round-to-nearest and no FMA -- also for distances, LUTs and
running-mean centroids.  The sum/count centroid update (the shardable mode) is compared within
1e-5 relative, the tolerance the north star states for floating point.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-5


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


# ---- KMeans.assign ----------------------------------------------------------------------------
@pytest.mark.parametrize("n,D,frm,dim,K", [
    (5000, 20, 3, 10, 256), (777, 8, 0, 8, 256), (1300, 16, 0, 16, 100), (513, 5, 2, 1, 7),
    (2000, 40, 5, 17, 33), (600, 300, 0, 300, 50), (1, 10, 0, 10, 256), (4096, 12, 1, 9, 1),
    (3000, 64, 7, 13, 600),
])
def test_assign_matches_oracle(g, oracle, n, D, frm, dim, K):
    rng = np.random.default_rng(n + 31 * dim + K)
    X = clustered(rng, n, D)
    Cm = rng.normal(size=(K, dim)).astype(np.float32) * 2
    got = g.KMeans(dim, Cm).assign(g.Vectors(g.Matrix(X), frm, frm + dim))
    want = oracle.assign(X, frm, dim, Cm, batch=0, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(got, want)
    got2 = g.KMeans(dim, Cm).par_assign(g.Vectors(g.Matrix(X), frm, frm + dim))
    assert np.array_equal(got2, want)


def test_assign_duplicate_centroids_lowest_index(g, oracle):
    # duplicate centroids (KMeans.init samples with replacement) => exact ties => lowest index
    rng = np.random.default_rng(5)
    X = clustered(rng, 3000, 10)
    Cm = rng.normal(size=(64, 10)).astype(np.float32)
    Cm[40:] = Cm[:24]
    got = g.KMeans(10, Cm).assign(g.Vectors(g.Matrix(X)))
    stats = np.zeros(2, np.int64)
    want = oracle.assign(X, 0, 10, Cm, tie_mode=oracle.TIE_LOWEST, stats=stats)
    assert stats[0] > 0  # the oracle saw tie events
    assert np.array_equal(got, want)
    assert got.max() < 40


def test_assign_empty(g):
    out = g.KMeans(4, np.zeros((3, 4), np.float32)).assign(g.Vectors(g.Matrix(np.zeros((0, 4), np.float32))))
    assert out.shape == (0,)


# ---- KMeans.init / fromAssignment ---------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 7, 29])
def test_init_matches_oracle(g, oracle, seed):
    rng = np.random.default_rng(seed)
    X = clustered(rng, 1000, 12)
    km, rows = g.KMeans.init(25, g.Vectors(g.Matrix(X), 2, 9), seed, return_rows=True)
    wc, wr = oracle.kmeans_init(X, 2, 7, 25, seed)
    assert np.array_equal(rows, wr)
    assert np.array_equal(km.centroids, wc)


@pytest.mark.parametrize("n,D,frm,dim,K", [(20000, 20, 4, 10, 256), (999, 9, 0, 9, 5),
                                           (30000, 40, 3, 33, 16), (5, 3, 0, 3, 8)])
def test_update_running_mean_bit_exact(g, oracle, n, D, frm, dim, K):
    rng = np.random.default_rng(n + K)
    X = clustered(rng, n, D)
    a = rng.integers(0, K, n).astype(np.int32)
    if K > 3:
        a[a == 2] = 1  # leave cluster 2 empty: it must stay all-zero
    km, cnt = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X), frm, frm + dim), a,
                                       g.UPDATE_RUNNING_MEAN, return_counts=True)
    wc, wcnt = oracle.from_assignment(X, frm, dim, a, K, return_counts=True)
    assert np.array_equal(cnt, wcnt)
    assert np.array_equal(km.centroids.view(np.uint32), wc.view(np.uint32))


@pytest.mark.parametrize("n,D,frm,dim,K", [(20000, 20, 4, 10, 256), (999, 9, 0, 9, 5),
                                           (30000, 40, 3, 33, 16)])
def test_update_sum_mode_close(g, oracle, n, D, frm, dim, K):
    rng = np.random.default_rng(n + K + 1)
    X = clustered(rng, n, D)
    a = rng.integers(0, K, n).astype(np.int32)
    km, cnt = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X), frm, frm + dim), a,
                                       g.UPDATE_SUM, return_counts=True)
    wc, wcnt = oracle.from_assignment(X, frm, dim, a, K, return_counts=True)
    assert np.array_equal(cnt, wcnt)
    ref = X[:, frm:frm + dim].astype(np.float64)
    for k in range(K):
        if wcnt[k]:
            exact = ref[a == k].mean(axis=0)
            scale = np.abs(ref[a == k]).mean() + 1e-30
            assert np.all(np.abs(km.centroids[k] - exact) <= REL * scale * 4)
        else:
            assert not km.centroids[k].any()


# ---- KMeans.computeClusters / ProductQuantizer.apply ------------------------------------------
@pytest.mark.parametrize("n,dim,K,iters", [(4000, 6, 16, 30), (2500, 10, 256, 5), (300, 3, 8, 0)])
def test_compute_clusters_bit_exact(g, oracle, n, dim, K, iters):
    rng = np.random.default_rng(n)
    X = clustered(rng, n, dim + 4, centres=9)
    reports = []
    km, info = g.KMeans.compute_clusters(
        g.Vectors(g.Matrix(X), 2, 2 + dim),
        g.KMeansConfig(K, iters, seed=3, report=reports.append), return_info=True)
    w = oracle.compute_clusters(X, 2, dim, K, iters, seed=3, tie_mode=oracle.TIE_LOWEST)
    assert info["updates"] == w["updates"]
    assert info["converged"] == w["converged"]
    assert np.array_equal(km.centroids.view(np.uint32), w["centroids"].view(np.uint32))
    # ProgressReport stream: the initial report, then one per loop pass (G/KMeans.scala:141-151)
    assert len(reports) == w["updates"] + 1
    assert reports[0].num_iterations == 0 and not reports[0].converged
    for r, wr in zip(reports[1:], w["report"]):
        assert r.num_iterations == int(wr[0])
        assert r.converged == bool(wr[3])
        assert r.step_mean == pytest.approx(float(wr[1]), rel=1e-6, abs=1e-12)
        assert r.step_stddev == pytest.approx(float(wr[2]), rel=1e-5, abs=1e-9)


def test_compute_clusters_sum_mode_objective(g, oracle):
    rng = np.random.default_rng(11)
    X = clustered(rng, 6000, 8, centres=10)
    km = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X)),
                                   g.KMeansConfig(10, 50, update_mode=g.UPDATE_SUM))
    w = oracle.compute_clusters(X, 0, 8, 10, 50, tie_mode=oracle.TIE_LOWEST)
    a = oracle.assign(X, 0, 8, km.centroids, tie_mode=oracle.TIE_LOWEST)
    obj = oracle.objective(X, 0, 8, km.centroids, a)
    wobj = oracle.objective(X, 0, 8, w["centroids"], w["assignments"])
    assert obj == pytest.approx(wobj, rel=1e-3)


@pytest.mark.parametrize("n,D,M,K,iters", [(3000, 20, 4, 32, 8), (1500, 11, 3, 16, 4)])
def test_pq_train_bit_exact(g, oracle, n, D, M, K, iters):
    rng = np.random.default_rng(D)
    X = clustered(rng, n, D)
    seen = []
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(K, M, iters, report=seen.append))
    wcb, wnu, wconv = oracle.pq_train(X, M, K, iters, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(pq.codebook().view(np.uint32), wcb.view(np.uint32))
    assert seen and seen[-1].total_iterations == M * iters
    frm, dim, _ = g.subvector_windows(D, M)
    assert [q.from_ for q in pq.quantizers] == list(frm)
    assert [q.dimension for q in pq.quantizers] == list(dim)


# ---- ProductQuantizer.encode / decode -----------------------------------------------------------
def random_codebook(rng, X, M, K):
    D = X.shape[1]
    dmax = -(-D // M)
    ideal = dmax
    full = M - (ideal * M - D)
    cb = np.zeros((M, K, dmax), np.float32)
    f = 0
    for m in range(M):
        d = ideal if m < full else ideal - 1
        rows = rng.integers(0, X.shape[0], K)
        cb[m, :, :d] = X[rows, f:f + d] + rng.normal(size=(K, d)).astype(np.float32) * 0.05
        f += d
    return cb


@pytest.mark.parametrize("n,D,M,K", [(20000, 100, 10, 256), (3001, 37, 5, 256), (1000, 128, 16, 256),
                                     (517, 30, 30, 17), (1, 20, 2, 256), (2048, 64, 4, 256)])
def test_encode_matches_oracle(g, oracle, n, D, M, K):
    rng = np.random.default_rng(n + D)
    X = clustered(rng, n, D)
    cb = random_codebook(rng, X, M, K)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(g.Matrix(X))
    want = oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST)
    assert enc.codes.dtype == np.uint8 and enc.codes.shape == (M, n)
    assert np.array_equal(enc.codes, want)
    # decode(encode) returns the chosen centroids exactly (T/ProductQuantizerSpec.scala:15-45)
    dec = pq.decode(enc).data
    assert np.array_equal(dec, oracle.pq_decode(want, cb, D))
    # re-encoding the reconstruction agrees with the oracle too (exact idempotence is not an fp32
    # property of the off - 2*dot score; T/ProductQuantizerSpec.scala:15-26 allows 1e-3)
    assert np.array_equal(pq.encode(g.Matrix(dec)).codes,
                          oracle.pq_encode(dec, cb, tie_mode=oracle.TIE_LOWEST))


def test_encode_chunked_host_path(g, oracle):
    rng = np.random.default_rng(77)
    X = clustered(rng, 10007, 24)
    cb = random_codebook(rng, X, 3, 256)
    pq = g.ProductQuantizer.from_codebook(cb, 24)
    g.set_option("encode_chunk_rows", 1000)
    try:
        got = pq.encode(X).codes
    finally:
        g.set_option("encode_chunk_rows", 1 << 18)
    assert np.array_equal(got, oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST))


def test_encode_empty(g):
    pq = g.ProductQuantizer.from_codebook(np.zeros((2, 4, 3), np.float32), 6)
    assert pq.encode(np.zeros((0, 6), np.float32)).codes.shape == (2, 0)


def test_encode_dev_matches_host(g):
    import torch
    rng = np.random.default_rng(3)
    X = clustered(rng, 5000, 40)
    cb = random_codebook(rng, X, 4, 256)
    pq = g.ProductQuantizer.from_codebook(cb, 40)
    codes = pq.encode_dev(torch.from_numpy(X).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(codes[:, :5000].cpu().numpy(), pq.encode(X).codes)


# ---- Index.prepareQuery ---------------------------------------------------------------------------
@pytest.mark.parametrize("nq,D,M,K", [(13, 100, 10, 256), (1, 37, 5, 100), (4100, 16, 2, 256)])
def test_prepare_query_bit_exact(g, oracle, nq, D, M, K):
    rng = np.random.default_rng(nq)
    X = clustered(rng, 2000, D)
    cb = random_codebook(rng, X, M, K)
    Q = clustered(rng, nq, D)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    got = g.prepare_query(pq, Q)
    want = oracle.prepare_query(Q, cb)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- PQIndex.batchQuery ---------------------------------------------------------------------------
def build_index(g, rng, n, D, M, K=256, codes=None):
    X = clustered(rng, max(n, 300), D)
    cb = random_codebook(rng, X, M, K)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    if codes is None:
        codes = rng.integers(0, K, (M, n)).astype(np.uint8)
    return pq, cb, codes, g.PQIndex(pq, g.EncodedMatrix(g.Coder8(n), codes))


IMPLS = ["simple", "fused", "pruned8", "pruned16", "pruned8w1", "pruned8w2", "pruned16w2"]


def impl_id(g, name):
    return {"simple": (g.SCAN_SIMPLE, 0, 0), "fused": (g.SCAN_FUSED, 0, 0),
            "pruned": (g.SCAN_PRUNED, 0, 0),
            "pruned8": (g.SCAN_PRUNED, 8, 4), "pruned16": (g.SCAN_PRUNED, 16, 4),
            "pruned8w1": (g.SCAN_PRUNED, 8, 1), "pruned8w2": (g.SCAN_PRUNED, 8, 2),
            "pruned16w2": (g.SCAN_PRUNED, 16, 2)}[name]


def check_query(g, oracle, ix, cb, codes, Q, k, frm, until, impl, boot_rows=4096, lb=0, stage_div=32):
    impl, bits, words = impl
    if bits == 8 and codes.shape[0] > 127:
        pytest.skip("8-bit lower-bound fields need M <= 127")
    g.set_option("scan_impl", impl)
    g.set_option("pruned_bits", bits)
    g.set_option("pruned_words", words)
    g.set_option("boot_rows", boot_rows)       # small boot so that test-sized ranges reach the
    g.set_option("pruned_lb_quantizers", lb)
    g.set_option("pruned_stage_div", stage_div)
    try:                                       # pruned kernel proper
        got = ix.batch_query(k, Q, frm, until)
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
        g.set_option("pruned_bits", 0)
        g.set_option("pruned_words", 0)
        g.set_option("boot_rows", 0)
        g.set_option("pruned_lb_quantizers", 0)
        g.set_option("pruned_stage_div", 32)
    ids, ds, sz = oracle.pq_query(Q, cb, codes, k, frm, until, topk_mode=oracle.TOPK_CANONICAL)
    assert np.array_equal(got.size, sz)
    for q in range(Q.shape[0]):
        n = sz[q]
        assert np.array_equal(got.values[q, :n].view(np.uint32), ds[q, :n].view(np.uint32)), q
        assert np.array_equal(got.keys[q, :n], ids[q, :n]), q
        assert np.all(got.keys[q, n:] == -1) and np.all(np.isinf(got.values[q, n:]))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("n,D,M,nq,k,frm,until", [
    (100000, 100, 10, 9, 10, 0, None),      # c1 shape, reduced N
    (60000, 128, 16, 4, 10, 5, 59990),      # c4 shape; ragged range
    (40000, 300, 30, 6, 10, 17, 39001),     # c2 shape
    (20000, 37, 5, 1, 1, 0, None),          # ragged windows, k = 1, single query
    (30000, 24, 3, 5, 128, 100, 29000),     # k at the fused limit
    (5000, 16, 2, 3, 10, 0, None),          # shorter than one scan item
    (9, 16, 2, 3, 10, 0, None),             # k > N
    (70000, 20, 2, 150, 10, 0, None),       # more query groups than one wave of splits
])
def test_query_matches_oracle(g, oracle, impl, n, D, M, nq, k, frm, until):
    rng = np.random.default_rng(n + nq)
    pq, cb, codes, ix = build_index(g, rng, n, D, M)
    Q = clustered(rng, nq, D)
    until = n if until is None else until
    check_query(g, oracle, ix, cb, codes, Q, k, frm, until, impl_id(g, impl))


@pytest.mark.parametrize("impl", IMPLS)
def test_query_heavy_ties(g, oracle, impl):
    # few distinct codes => many equal distances => (distance, id) order decides
    rng = np.random.default_rng(2)
    n, D, M = 50000, 16, 2
    codes = rng.integers(0, 3, (M, n)).astype(np.uint8)
    pq, cb, codes, ix = build_index(g, rng, n, D, M, codes=codes)
    Q = clustered(rng, 5, D)
    check_query(g, oracle, ix, cb, codes, Q, 25, 3, n - 1, impl_id(g, impl))


@pytest.mark.parametrize("impl", ["pruned8", "pruned16", "pruned", "pruned8w1", "pruned8w2",
                                  "pruned16w2"])
@pytest.mark.parametrize("n,D,M,nq,k,boot", [
    (300000, 100, 10, 21, 10, 65536),     # default boot, several tiles
    (200000, 300, 30, 8, 10, 8192),       # c2 shape
    (150000, 128, 16, 3, 100, 4096),      # c4 shape, large k
    (120000, 1000, 100, 5, 10, 4096),     # c5 shape: M = 100 sub-quantizers
    (100000, 16, 2, 9, 1, 16),            # boot shorter than one vector load
])
def test_query_pruned_matches_oracle_clustered_codes(g, oracle, impl, n, D, M, nq, k, boot):
    """Codes produced by encoding clustered data (realistic distance distribution: the lower bound
    prunes almost everything) instead of uniform random codes."""
    rng = np.random.default_rng(n + M)
    X = clustered(rng, n, D, centres=40)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    Q = clustered(rng, nq, D, centres=40)
    Q[0] = X[n // 2]                       # a query that coincides with a database row
    check_query(g, oracle, ix, cb, enc.codes, Q, k, 0, n, impl_id(g, impl), boot_rows=boot)


@pytest.mark.parametrize("impl", ["pruned", "pruned8", "pruned16", "pruned8w1"])
@pytest.mark.parametrize("lb,stage_div", [(1, 0), (1, 2), (3, 2), (7, 3), (1000, 2), (0, 2)])
@pytest.mark.parametrize("n,D,M,nq,k,boot", [
    (200000, 300, 30, 19, 10, 8192),      # c2 shape
    (170000, 1000, 100, 5, 10, 4096),     # c5 shape: 8-bit fields only through the subset
    (150000, 37, 5, 40, 100, 4096),       # ragged windows, large k, several tiles
])
def test_query_pruned_subset_bound_and_stages(g, oracle, impl, lb, stage_div, n, D, M, nq, k, boot):
    """The lower bound over a SUBSET of the quantizers (pruned_lb_quantizers) and the two-stage scan
    (first range / stage_div rows with the full bound, tables re-quantised against the merged lists)
    change only how much is pruned, never the result."""
    if impl in ("pruned8", "pruned8w1") and min(lb, M) > 127:
        pytest.skip("8-bit lower-bound fields hold at most 127 quantizers")
    rng = np.random.default_rng(n + M + lb)
    X = clustered(rng, n, D, centres=40)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    Q = clustered(rng, nq, D, centres=40)
    Q[0] = X[n // 3]
    check_query(g, oracle, ix, cb, enc.codes, Q, k, 0, n, impl_id(g, impl), boot_rows=boot, lb=lb,
                stage_div=stage_div)
    if lb:
        assert g._native.counter("pscan_lb_quantizers") == min(lb, M)


def test_query_pruned_feedback_keeps_results(g, oracle):
    """Auto mode adapts the subset size from the survivor rate of earlier launches on the same index;
    every launch still returns the oracle's answer."""
    rng = np.random.default_rng(77)
    n, D, M = 160000, 120, 12
    X = clustered(rng, n, D, centres=40)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    seen = set()
    for it in range(6):
        Q = clustered(rng, 33, D, centres=40)
        check_query(g, oracle, ix, cb, enc.codes, Q, 10, 0, n, impl_id(g, "pruned"), boot_rows=4096)
        seen.add(g._native.counter("pscan_lb_quantizers"))
    assert all(1 <= v <= M for v in seen)


@pytest.mark.parametrize("impl", ["fused", "pruned8", "pruned16", "pruned8w1", "pruned8w2"])
def test_query_descending_distances_overflow_path(g, oracle, impl):
    # rows ordered by DEcreasing distance: every row beats the running k-th best, the worst case
    # for the fused kernel's candidate buffer
    rng = np.random.default_rng(4)
    n, D, M = 40000, 8, 1
    pq, cb, _, _ = build_index(g, rng, 10, D, M)
    Q = clustered(rng, 4, D)
    lut = oracle.prepare_query(Q, cb)[0, 0]
    order = np.argsort(-lut, kind="stable").astype(np.uint8)     # codes by decreasing distance
    codes = order[(np.arange(n) * 256 // n)].reshape(1, n)
    ix = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(n), codes))
    check_query(g, oracle, ix, cb, codes, Q, 10, 0, n, impl_id(g, impl))


def test_query_cosine_normalises_queries(g, oracle):
    rng = np.random.default_rng(8)
    pq, cb, codes, ix = build_index(g, rng, 3000, 20, 4)
    Q = clustered(rng, 7, 20)
    got = ix.batch_query(5, Q, normalize=True)
    Qn = oracle.normalize(Q)
    assert np.array_equal(g.normalize(Q).view(np.uint32), Qn.view(np.uint32))
    ids, ds, sz = oracle.pq_query(Qn, cb, codes, 5)
    assert np.array_equal(got.keys, ids) and np.array_equal(got.values, ds)


def test_query_range_errors_and_empty(g):
    rng = np.random.default_rng(9)
    pq, cb, codes, ix = build_index(g, rng, 1000, 12, 3)
    Q = clustered(rng, 2, 12)
    with pytest.raises(ValueError):
        ix.batch_query(3, Q, 10, 5)           # require(from <= until)
    with pytest.raises(ValueError):
        ix.batch_query(3, Q, 0, 1001)         # require(until <= length)
    with pytest.raises(ValueError):
        ix.batch_query(3, Q, -1, 10)
    r = ix.batch_query(3, Q, 7, 7)
    assert np.all(r.size == 0) and np.all(r.keys == -1)
    r0 = ix.batch_query(0, Q)
    assert r0.keys.shape == (2, 0)


def test_query_split_ranges_merge_to_whole(g):
    # TopKHeap#merge property (T/TopKHeapSpec.scala:33-52): per-range results merged == whole scan
    import torch
    import ctypes as C
    from gulon_b200 import _native as N
    rng = np.random.default_rng(10)
    n = 50000
    pq, cb, codes, ix = build_index(g, rng, n, 32, 4)
    Q = clustered(rng, 11, 32)
    whole = ix.batch_query(10, Q)
    cuts = [0, 7000, 7001, 30000, n]
    parts = [ix.batch_query(10, Q, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    ids = torch.from_numpy(np.stack([p.keys for p in parts])).cuda()
    ds = torch.from_numpy(np.stack([p.values for p in parts])).cuda()
    oi = torch.empty((11, 10), dtype=torch.int32, device="cuda")
    od = torch.empty((11, 10), dtype=torch.float32, device="cuda")
    oz = torch.empty((11,), dtype=torch.int32, device="cuda")
    N.check(N.lib().gulon_topk_merge_dev(ids.data_ptr(), ds.data_ptr(), len(parts), 11, 10,
                                         oi.data_ptr(), od.data_ptr(), oz.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(oi.cpu().numpy(), whole.keys)
    assert np.array_equal(od.cpu().numpy(), whole.values)
    assert np.all(oz.cpu().numpy() == 10)


def test_query_dev_matches_host(g):
    import torch
    rng = np.random.default_rng(12)
    pq, cb, codes, ix = build_index(g, rng, 40000, 24, 3)
    Q = clustered(rng, 33, 24)
    host = ix.batch_query(10, Q, id_offset=1000)
    ids, ds, sz = ix.batch_query_dev(10, torch.from_numpy(Q).cuda(), id_offset=1000)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy(), host.keys)
    assert np.array_equal(ds.cpu().numpy(), host.values)
    assert host.keys.min() >= 1000


# ---- exact kNN / re-rank ----------------------------------------------------------------------------
@pytest.mark.parametrize("n,D,nq,k,frm,until", [(5000, 30, 7, 10, 0, None), (3000, 300, 3, 5, 11, 2999),
                                                (40, 8, 2, 50, 0, None)])
def test_exact_topk_matches_oracle(g, oracle, n, D, nq, k, frm, until):
    rng = np.random.default_rng(n)
    X = clustered(rng, n, D)
    Q = clustered(rng, nq, D)
    until = n if until is None else until
    got = g.exact_nearest_neighbours(g.Matrix(X), Q, k, frm, until)
    ids, ds, sz = oracle.exact_nn(X, Q, k, frm, until, topk_mode=oracle.TOPK_CANONICAL)
    assert np.array_equal(got.size, sz)
    for q in range(nq):
        m = sz[q]
        assert np.array_equal(got.values[q, :m].view(np.uint32), ds[q, :m].view(np.uint32))
        assert np.array_equal(got.keys[q, :m], ids[q, :m])


def test_rerank_matches_exact_subset(g, oracle):
    from gulon_b200.index import rerank
    rng = np.random.default_rng(21)
    X = clustered(rng, 4000, 100)
    Q = clustered(rng, 6, 100)
    cand = np.stack([rng.choice(4000, 200, replace=False) for _ in range(6)]).astype(np.int32)
    cand[:, ::17] = -1
    got = rerank(g.Matrix(X), Q, cand, 10)
    for q in range(6):
        c = cand[q][cand[q] >= 0]
        d = np.array([oracle.distance_sq(X[i], Q[q]) for i in c], np.float32)
        o = np.lexsort((c, d))[:10]
        assert np.array_equal(got.keys[q], c[o])
        assert np.array_equal(got.values[q].view(np.uint32), d[o].view(np.uint32))


# ---- full-size properties (no oracle: the domain's own invariants) ---------------------------------
def test_large_scan_split_invariance(g):
    """1M x m=16 codes, fused scan: scanning [0,N) equals merging two disjoint ranges."""
    rng = np.random.default_rng(31)
    n = 1_000_000
    pq, cb, codes, ix = build_index(g, rng, n, 128, 16)
    Q = clustered(rng, 40, 128)
    whole = ix.batch_query(10, Q)
    a = ix.batch_query(10, Q, 0, 400_003)
    b = ix.batch_query(10, Q, 400_003, n)
    for q in range(40):
        ids = np.concatenate([a.keys[q], b.keys[q]])
        ds = np.concatenate([a.values[q], b.values[q]])
        o = np.lexsort((ids, ds))[:10]
        assert np.array_equal(ids[o], whole.keys[q])
        assert np.array_equal(ds[o], whole.values[q])
    g.set_option("scan_impl", g.SCAN_SIMPLE)
    try:
        simple = ix.batch_query(10, Q[:8])
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
    assert np.array_equal(simple.keys, whole.keys[:8])
    assert np.array_equal(simple.values, whole.values[:8])


def test_self_query_finds_own_code(g):
    """A decoded database row queried against the index has ADC distance 0 to itself
    (T/IndexSpec.scala:62-73 analogue)."""
    rng = np.random.default_rng(41)
    n, D, M = 200_000, 40, 8
    X = clustered(rng, n, D, centres=200)
    pq = g.ProductQuantizer.train(g.Matrix(X[:20000]), g.ProductQuantizerConfig(256, M, 3))
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    rows = rng.integers(0, n, 20)
    Q = pq.decode(g.EncodedMatrix(g.Coder8(20), enc.codes[:, rows])).data
    r = ix.batch_query(1, Q)
    assert np.all(r.values[:, 0] == 0.0)
    for i, row in enumerate(rows):
        assert np.array_equal(enc.codes[:, r.keys[i, 0]], enc.codes[:, row])
        assert r.keys[i, 0] <= row  # lowest id among identical codes


# ---- sharded k-means: two ranks emulated by two host threads on one GPU ------------------------------
def test_sharded_kmeans_two_ranks_one_gpu(g, oracle):
    """gulon_kmeans_train with gulon_comm_t hooks: rows sharded over 2 ranks, one all-reduce of
    sums/counts (+ one of the changed-assignment count) per Lloyd iteration.  The ranks are two host
    threads sharing the GPU; their hooks exchange through a host barrier (no kernel ever waits on
    another rank)."""
    import ctypes as C
    import threading
    import torch
    from gulon_b200 import _native as N
    from gulon_b200.sharded import _view

    rng = np.random.default_rng(17)
    n, D, K = 20000, 12, 16
    X = clustered(rng, n, D, centres=16, scale=4.0, noise=0.3)
    world = 2
    bounds = [(0, 9984), (9984, n)]
    bar = threading.Barrier(world)
    slots = [None] * world
    dev = torch.device("cuda", 0)
    calls = [0] * world

    def make_comm(rank):
        def allreduce(buf, cnt, typestr):
            try:
                t = _view(buf, cnt, typestr, dev)
                torch.cuda.synchronize()
                slots[rank] = t
                bar.wait()
                total = slots[0].clone()
                for r in range(1, world):
                    total += slots[r]
                torch.cuda.synchronize()
                bar.wait()
                t.copy_(total)
                torch.cuda.synchronize()
                bar.wait()
                calls[rank] += 1
                return 0
            except Exception:
                bar.abort()
                return 1
        f32 = N.Comm.ALLREDUCE_F32(lambda u, b, c, s: allreduce(b, c, "<f4"))
        i32 = N.Comm.ALLREDUCE_I32(lambda u, b, c, s: allreduce(b, c, "<i4"))
        ag = N.Comm.ALLGATHER(lambda u, a, b, c, s: 1)
        return N.Comm(rank, world, f32, i32, ag, None), (f32, i32, ag)

    out = [None] * world
    errs = []

    def run(rank):
        try:
            lo, hi = bounds[rank]
            comm, keep = make_comm(rank)
            km, info = g.KMeans.compute_clusters(
                g.Vectors(g.Matrix(X[lo:hi])), g.KMeansConfig(K, 40, seed=5, update_mode=g.UPDATE_SUM),
                comm=comm, n_total=n, row_offset=lo, return_info=True)
            out[rank] = (km.centroids.copy(), info)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            bar.abort()

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(120)
    assert not errs, errs
    (c0, i0), (c1, i1) = out
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32))     # every rank ends identical
    assert i0 == i1 and calls[0] == calls[1] > 0
    # same problem on one rank, same update rule
    km, info = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X)),
                                         g.KMeansConfig(K, 40, seed=5, update_mode=g.UPDATE_SUM),
                                         return_info=True)
    a_sh = oracle.assign(X, 0, D, c0, tie_mode=oracle.TIE_LOWEST)
    a_1 = oracle.assign(X, 0, D, km.centroids, tie_mode=oracle.TIE_LOWEST)
    o_sh = oracle.objective(X, 0, D, c0, a_sh)
    o_1 = oracle.objective(X, 0, D, km.centroids, a_1)
    assert o_sh == pytest.approx(o_1, rel=1e-4)
    assert info["converged"] == i0["converged"]
    if np.array_equal(a_sh, a_1):
        assert np.allclose(c0, km.centroids, rtol=1e-5, atol=1e-5)
