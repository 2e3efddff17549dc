"""GPU parity of the tensor-core assignment / encode path (gulon_b200/csrc/tcassign.cuh).

The tcgen05 contraction only FILTERS centroids; the winners are re-evaluated in the reference's
literal fp32 arithmetic (G/KMeans.scala:24-55,70-98), so assignments and PQ codes must be
bit-identical to the oracle and to the exact CUDA-core kernel -- including exact ties (lowest index),
near ties one ulp apart, duplicate centroids, non-finite inputs and the every-chunk-is-a-candidate case.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXACT, TENSOR = 1, 2


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


@pytest.fixture()
def tensor(g):
    """Force the tensor path and count what it did."""
    from gulon_b200 import _native
    g.set_option("assign_impl", TENSOR)
    g.set_option("profile", 1)
    yield _native
    g.set_option("profile", 0)
    g.set_option("assign_impl", 0)


def clustered(rng, n, d, centres=12, scale=3.0, noise=0.4):
    c = rng.normal(size=(centres, d)).astype(np.float32) * scale
    x = c[rng.integers(0, centres, n)] + rng.normal(size=(n, d)).astype(np.float32) * noise
    return np.ascontiguousarray(x, np.float32)


def assign_both(g, X, frm, dim, Cm):
    v = g.Vectors(g.Matrix(X), frm, frm + dim)
    g.set_option("assign_impl", TENSOR)
    t = g.KMeans(dim, Cm).assign(v)
    g.set_option("assign_impl", EXACT)
    e = g.KMeans(dim, Cm).assign(v)
    g.set_option("assign_impl", TENSOR)
    return t, e


@pytest.mark.parametrize("n,D,frm,dim,K", [
    (5000, 20, 3, 10, 256), (777, 8, 0, 8, 256), (1300, 16, 1, 15, 100), (513, 5, 2, 1, 7),
    (1, 10, 0, 10, 256), (4096, 12, 1, 9, 1), (127, 3, 0, 3, 2), (129, 14, 0, 14, 255),
    (20000, 7, 0, 7, 256), (8193, 30, 10, 10, 256), (300, 2, 0, 2, 9), (999, 4, 0, 4, 64),
    (2500, 5, 0, 5, 256), (2500, 6, 0, 6, 31), (2500, 11, 0, 11, 200), (2500, 12, 0, 12, 256),
    (2500, 13, 0, 13, 256),
])
def test_tensor_assign_matches_oracle(g, oracle, tensor, n, D, frm, dim, K):
    rng = np.random.default_rng(n + 31 * dim + K)
    X = clustered(rng, n, D)
    Cm = rng.normal(size=(K, dim)).astype(np.float32) * 2
    t, e = assign_both(g, X, frm, dim, Cm)
    want = oracle.assign(X, frm, dim, Cm, batch=0, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(e, want)
    assert np.array_equal(t, want)
    assert tensor.counter("assign_tc_rows") >= n  # the tensor kernel really ran


@pytest.mark.parametrize("n,D,frm,dim", [(300_000, 12, 1, 10), (150_001, 40, 8, 8), (200_000, 33, 3, 15)])
def test_tensor_assign_single_window_many_blocks(g, oracle, tensor, n, D, frm, dim):
    # one window per group: every sweep group skips row blocks, the raw-tile ring is recycled fastest
    rng = np.random.default_rng(n)
    X = clustered(rng, n, D, centres=200)
    Cm = X[rng.integers(0, n, 256), frm:frm + dim].copy()
    t, e = assign_both(g, X, frm, dim, Cm)
    assert np.array_equal(t, e)
    sl = slice(n // 2, n // 2 + 5000)
    want = oracle.assign(np.ascontiguousarray(X[sl]), frm, dim, Cm, batch=0, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(t[sl], want)


def test_tensor_assign_centroids_from_data(g, oracle, tensor):
    # centroids sampled from the rows (KMeans.init): a row equals its own centroid, near ties abound
    rng = np.random.default_rng(11)
    X = clustered(rng, 30000, 10, centres=300, noise=0.05)
    Cm = X[rng.integers(0, X.shape[0], 256)].copy()
    t, e = assign_both(g, X, 0, 10, Cm)
    want = oracle.assign(X, 0, 10, Cm, batch=0, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(t, want) and np.array_equal(e, want)
    pairs, rows = tensor.counter("assign_tc_pairs"), tensor.counter("assign_tc_rows")
    assert rows == 30000 and pairs >= rows  # at least the winner's chunk per row
    assert pairs < 8 * rows                 # and the filter does discard most chunks


def test_tensor_assign_duplicate_centroids_lowest_index(g, oracle, tensor):
    rng = np.random.default_rng(5)
    X = clustered(rng, 3000, 10)
    Cm = rng.normal(size=(256, 10)).astype(np.float32)
    Cm[128:] = Cm[:128]          # every centroid twice, 16 chunks apart
    stats = np.zeros(2, np.int64)
    want = oracle.assign(X, 0, 10, Cm, tie_mode=oracle.TIE_LOWEST, stats=stats)
    assert stats[0] > 0
    t, e = assign_both(g, X, 0, 10, Cm)
    assert np.array_equal(t, want) and np.array_equal(e, want)
    assert t.max() < 128


def test_tensor_assign_one_ulp_apart(g, oracle, tensor):
    # centroid pairs that differ in the last bit of one coordinate: the approximate scores cannot
    # tell them apart, the exact recheck must
    rng = np.random.default_rng(6)
    X = clustered(rng, 6000, 8)
    Cm = rng.normal(size=(256, 8)).astype(np.float32) * 2
    Cm[1::2] = Cm[0::2]
    Cm[1::2, 3] = np.nextafter(Cm[0::2, 3], np.float32(np.inf))
    t, e = assign_both(g, X, 0, 8, Cm)
    want = oracle.assign(X, 0, 8, Cm, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(t, want) and np.array_equal(e, want)
    assert (want % 2 == 1).any() and (want % 2 == 0).any()


def test_tensor_assign_all_equal_centroids_every_chunk(g, oracle, tensor):
    # all centroids identical: every chunk of every row is a candidate
    rng = np.random.default_rng(7)
    X = clustered(rng, 1000, 10)
    Cm = np.tile(rng.normal(size=(1, 10)).astype(np.float32), (256, 1))
    t, e = assign_both(g, X, 0, 10, Cm)
    assert not t.any() and not e.any()
    assert tensor.counter("assign_tc_pairs") == 32 * 1000  # every chunk of every row was rechecked
    # all-zero centroids (two or more empty clusters, G/KMeans.scala:198-226)
    t, e = assign_both(g, X, 0, 10, np.zeros((256, 10), np.float32))
    assert not t.any() and not e.any()


def test_tensor_assign_non_finite_and_huge_rows(g, oracle, tensor):
    rng = np.random.default_rng(8)
    X = clustered(rng, 2000, 6)
    X[3, 2] = np.nan
    X[10, 0] = np.inf
    X[11, 5] = -np.inf
    X[500] = 3e19
    X[501] = -1e30
    X[502, 1] = 1e38
    X[900:910] = 0.0
    Cm = rng.normal(size=(256, 6)).astype(np.float32) * 2
    t, e = assign_both(g, X, 0, 6, Cm)
    want = oracle.assign(X, 0, 6, Cm, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(e, want)
    assert np.array_equal(t, want)
    # huge / non-finite centroids
    Cm2 = Cm.copy()
    Cm2[7] = 1e20
    Cm2[9, 1] = np.inf
    Cm2[200, 0] = np.nan
    t, e = assign_both(g, X, 0, 6, Cm2)
    want = oracle.assign(X, 0, 6, Cm2, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(e, want)
    assert np.array_equal(t, want)


def test_tensor_assign_tiny_and_mixed_scales(g, oracle, tensor):
    rng = np.random.default_rng(9)
    X = clustered(rng, 5000, 10) * np.float32(1e-20)
    Cm = (rng.normal(size=(256, 10)) * 1e-20).astype(np.float32)
    t, e = assign_both(g, X, 0, 10, Cm)
    want = oracle.assign(X, 0, 10, Cm, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(t, want) and np.array_equal(e, want)
    # per-column scales 1e-3 .. 1e3
    sc = np.float32(10.0) ** rng.integers(-3, 4, 10).astype(np.float32)
    X = clustered(rng, 5000, 10) * sc
    Cm = X[rng.integers(0, 5000, 256)] + (rng.normal(size=(256, 10)) * 1e-3).astype(np.float32) * sc
    t, e = assign_both(g, X, 0, 10, Cm.astype(np.float32))
    want = oracle.assign(X, 0, 10, Cm.astype(np.float32), tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(t, want) and np.array_equal(e, want)


def random_codebook(rng, X, M, K):
    D = X.shape[1]
    dmax = -(-D // M)
    full = M - (dmax * M - D)
    cb = np.zeros((M, K, dmax), np.float32)
    f = 0
    for m in range(M):
        d = dmax if m < full else dmax - 1
        rows = rng.integers(0, X.shape[0], K)
        cb[m, :, :d] = X[rows, f:f + d] + rng.normal(size=(K, d)).astype(np.float32) * 0.05
        f += d
    return cb


@pytest.mark.parametrize("n,D,M,K", [(20000, 100, 10, 256), (3001, 37, 5, 256), (1000, 128, 16, 256),
                                     (517, 30, 30, 17), (1, 20, 2, 256), (9000, 300, 30, 256),
                                     (4100, 43, 4, 256), (2048, 1000, 100, 256)])
def test_tensor_encode_matches_oracle(g, oracle, tensor, n, D, M, K):
    rng = np.random.default_rng(n + D)
    X = clustered(rng, n, D)
    cb = random_codebook(rng, X, M, K)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    got = pq.encode(g.Matrix(X)).codes
    want = oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST)
    assert got.dtype == np.uint8 and got.shape == (M, n)
    assert np.array_equal(got, want)
    assert tensor.counter("assign_tc_rows") >= n * M


def test_tensor_encode_large_matches_exact_kernel(g, oracle, tensor):
    # 400k x 300 (c2's shape, 4 % of its rows): tensor path vs the exact CUDA-core kernel on all
    # rows, vs the oracle on a slice
    import torch
    rng = np.random.default_rng(123)
    n, D, M = 400_000, 300, 30
    cen = rng.normal(size=(2048, D)).astype(np.float32)
    X = cen[rng.integers(0, 2048, n)] + rng.normal(size=(n, D)).astype(np.float32) * 0.5
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    dX = torch.from_numpy(X).cuda()
    t = pq.encode_dev(dX)[:, :n].cpu().numpy()
    g.set_option("assign_impl", EXACT)
    e = pq.encode_dev(dX)[:, :n].cpu().numpy()
    assert np.array_equal(t, e)
    sl = slice(123_000, 131_000)
    assert np.array_equal(t[:, sl], oracle.pq_encode(X[sl], cb, tie_mode=oracle.TIE_LOWEST))
    pairs, rows = tensor.counter("assign_tc_pairs"), tensor.counter("assign_tc_rows")
    assert rows == n * M
    print("candidate chunks per (row, window): %.3f" % (pairs / rows))


def test_tensor_kmeans_training_bit_exact(g, oracle, tensor):
    # Lloyd iterations through the tensor-core assignment: same centroids, bit for bit
    rng = np.random.default_rng(21)
    X = clustered(rng, 9000, 10, centres=40)
    km, info = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X)), g.KMeansConfig(64, 6, seed=3),
                                         return_info=True)
    w = oracle.compute_clusters(X, 0, 10, 64, 6, seed=3, tie_mode=oracle.TIE_LOWEST)
    assert info["updates"] == w["updates"] and info["converged"] == w["converged"]
    assert np.array_equal(km.centroids.view(np.uint32), w["centroids"].view(np.uint32))
    assert tensor.counter("assign_tc_rows") >= 9000
