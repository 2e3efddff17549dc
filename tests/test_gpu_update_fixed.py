"""GPU tests of the shardable centroid update (GULON_UPDATE_SUM) as an exact fixed-point sum
(gulon_b200/csrc/kupdate.cuh).

Reference: KMeans.fromAssignment, G/KMeans.scala:198-226 (running mean; the sum/count form is the
shardable restatement).  The fixed-point sum is an integer computation, so the tests pin
it bit for bit against a numpy int64 model, check that it is reproducible, and that training over
two ranks gives exactly the single-rank centroids.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIX_BITS = 28


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


def integer_model(X, frm, dim, a, K):
    """The kernel's arithmetic restated with numpy integers."""
    W = X[:, frm:frm + dim]
    fin = np.isfinite(W)
    amax = np.abs(W[fin]).max() if fin.any() else 0.0
    e = int(np.frexp(np.float32(amax))[1]) if amax > 0 else 0   # 2^e > amax >= 2^(e-1)
    s = min(100, max(-100, FIX_BITS - e))
    scale = np.float32(2.0) ** np.float32(s)
    v = np.rint(W * scale).astype(np.int64)                      # x * 2^S is exact in fp32
    out = np.zeros((K, dim), np.float32)
    cnt = np.bincount(a, minlength=K).astype(np.int32)
    for k in range(K):
        if cnt[k]:
            tot = v[a == k].sum(axis=0)
            out[k] = (tot.astype(np.float64) / np.float64(scale) / np.float64(cnt[k])).astype(np.float32)
    return out, cnt


@pytest.mark.parametrize("n,D,frm,dim,K", [(20000, 20, 4, 10, 256), (999, 9, 0, 9, 5), (70000, 32, 3, 15, 200),
                                           (5000, 8, 0, 8, 256), (1, 12, 2, 10, 3), (300000, 12, 1, 10, 256),
                                           (4097, 40, 7, 16, 64), (3000, 5, 2, 1, 7)])
def test_fixed_sum_is_the_integer_model(g, n, D, frm, dim, K):
    rng = np.random.default_rng(n + K)
    X = clustered(rng, n, D)
    a = rng.integers(0, K, n).astype(np.int32)
    km, cnt = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X), frm, frm + dim), a,
                                       g.UPDATE_SUM, return_counts=True)
    want, wcnt = integer_model(X, frm, dim, a, K)
    assert np.array_equal(cnt, wcnt)
    assert np.array_equal(km.centroids.view(np.uint32), want.view(np.uint32))
    # and it is the mean to well within the north star's 1e-5
    ref = X[:, frm:frm + dim].astype(np.float64)
    for k in np.flatnonzero(wcnt)[:50]:
        exact = ref[a == k].mean(axis=0)
        assert np.all(np.abs(km.centroids[k] - exact) <= 1e-6 * (np.abs(ref).max() + 1e-30))


def test_fixed_sum_skewed_clusters_and_scales(g):
    # one huge cluster (many carries out of the low word), tiny and large magnitudes, negative values
    rng = np.random.default_rng(3)
    n, dim, K = 400000, 10, 256
    X = (rng.normal(size=(n, dim)) * 1000.0 - 700.0).astype(np.float32)
    X[::7] *= np.float32(1e-6)
    a = np.zeros(n, np.int32)
    a[::1000] = rng.integers(1, K, len(a[::1000]))
    km, cnt = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X)), a, g.UPDATE_SUM, return_counts=True)
    want, wcnt = integer_model(X, 0, dim, a, K)
    assert np.array_equal(cnt, wcnt)
    assert np.array_equal(km.centroids.view(np.uint32), want.view(np.uint32))


def test_fixed_sum_reproducible_and_order_independent(g):
    rng = np.random.default_rng(5)
    n, D, K = 150000, 30, 256
    X = clustered(rng, n, D)
    a = rng.integers(0, K, n).astype(np.int32)
    v = g.Vectors(g.Matrix(X), 10, 20)
    c1 = g.KMeans.from_assignment(K, 10, v, a, g.UPDATE_SUM).centroids
    c2 = g.KMeans.from_assignment(K, 10, v, a, g.UPDATE_SUM).centroids
    assert np.array_equal(c1.view(np.uint32), c2.view(np.uint32))
    # a row permutation changes every partial order of addition, not the result
    p = rng.permutation(n)
    c3 = g.KMeans.from_assignment(K, 10, g.Vectors(g.Matrix(X[p]), 10, 20), a[p], g.UPDATE_SUM).centroids
    assert np.array_equal(c1.view(np.uint32), c3.view(np.uint32))


def test_fixed_sum_non_finite_rows_poison_only_their_cluster(g):
    rng = np.random.default_rng(6)
    n, dim, K = 5000, 6, 8
    X = clustered(rng, n, dim)
    a = rng.integers(0, K, n).astype(np.int32)
    X[17, 2] = np.nan
    X[99, 0] = np.inf
    km = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X)), a, g.UPDATE_SUM)
    poisoned = {int(a[17]), int(a[99])}
    Xc = X.copy()
    Xc[17, 2] = 0
    Xc[99, 0] = 0
    want, _ = integer_model(Xc, 0, dim, a, K)
    for k in range(K):
        if k in poisoned:
            assert np.isnan(km.centroids[k]).all()
        else:
            assert np.array_equal(km.centroids[k].view(np.uint32), want[k].view(np.uint32))


def test_fp32_sum_path_still_available(g, oracle):
    rng = np.random.default_rng(8)
    n, dim, K = 20000, 10, 64
    X = clustered(rng, n, dim)
    a = rng.integers(0, K, n).astype(np.int32)
    g.set_option("update_fixed", 0)
    try:
        km = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X)), a, g.UPDATE_SUM)
    finally:
        g.set_option("update_fixed", 1)
    want, _ = integer_model(X, 0, dim, a, K)
    assert np.allclose(km.centroids, want, rtol=1e-5, atol=1e-5)


def test_sharded_training_is_bit_identical_to_one_rank(g):
    """gulon_kmeans_train over 2 and 3 ranks (host threads sharing the GPU, hooks exchanging through
    a host barrier) with the int64 / max hooks: the centroids equal the single-rank run bit for bit."""
    import threading
    import torch
    from gulon_b200 import _native as N
    from gulon_b200.sharded import _view

    rng = np.random.default_rng(17)
    n, D, K = 30000, 12, 32
    X = clustered(rng, n, D, centres=32, scale=4.0, noise=0.6)
    cfg = dict(seed=5, update_mode=g.UPDATE_SUM)
    one, info1 = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X), 1, 11), g.KMeansConfig(K, 25, **cfg),
                                           return_info=True)
    dev = torch.device("cuda", 0)

    for world, cuts in ((2, [0, 14992, n]), (3, [0, 7008, 21000, n])):
        bar = threading.Barrier(world)
        slots = [None] * world
        calls = [dict(i64=0, mx=0) for _ in range(world)]

        def make_comm(rank):
            def allreduce(buf, cnt, typestr, op, key=None):
                try:
                    t = _view(buf, cnt, typestr, dev)
                    torch.cuda.synchronize()
                    slots[rank] = t
                    bar.wait()
                    total = slots[0].clone()
                    for r in range(1, world):
                        total = torch.maximum(total, slots[r]) if op == "max" else total + slots[r]
                    torch.cuda.synchronize()
                    bar.wait()
                    t.copy_(total)
                    torch.cuda.synchronize()
                    bar.wait()
                    if key:
                        calls[rank][key] += 1
                    return 0
                except Exception:
                    bar.abort()
                    return 1
            cbs = (N.Comm.ALLREDUCE_F32(lambda u, b, c, s: allreduce(b, c, "<f4", "sum")),
                   N.Comm.ALLREDUCE_I32(lambda u, b, c, s: allreduce(b, c, "<i4", "sum")),
                   N.Comm.ALLGATHER(lambda u, a, b, c, s: 1),
                   N.Comm.ALLREDUCE_I64(lambda u, b, c, s: allreduce(b, c, "<i8", "sum", "i64")),
                   N.Comm.ALLREDUCE_MAX_F32(lambda u, b, c, s: allreduce(b, c, "<f4", "max", "mx")))
            return N.Comm(rank, world, cbs[0], cbs[1], cbs[2], None, cbs[3], cbs[4]), cbs

        out = [None] * world
        errs = []

        def run(rank):
            try:
                lo, hi = cuts[rank], cuts[rank + 1]
                comm, keep = make_comm(rank)
                km, info = g.KMeans.compute_clusters(
                    g.Vectors(g.Matrix(X[lo:hi]), 1, 11), g.KMeansConfig(K, 25, **cfg),
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
        for r in range(world):
            assert np.array_equal(out[r][0].view(np.uint32), one.centroids.view(np.uint32)), (world, r)
            assert out[r][1] == info1
            assert calls[r]["i64"] > 0 and calls[r]["mx"] == 1


# ---- fused Lloyd pass (assignment + sums + changed count in one read of the matrix) -------------------
@pytest.mark.parametrize("n,D,M,K,iters", [
    (50000, 40, 4, 256, 6),        # width 10, the BASELINE width
    (30000, 37, 5, 100, 5),        # ragged windows: widths 8 and 7, K < 256
    (4097, 16, 2, 256, 3),         # width 8, a 1-row last tile
    (200, 30, 3, 16, 4),           # fewer rows than one tile
    (20000, 45, 3, 256, 4),        # width 15: the largest the tensor path takes
    (70000, 300, 30, 256, 3),      # c2 / c3 shape: 30 windows, 15 units of 2 per row range
    (9000, 128, 16, 256, 25),      # c4 shape, runs to convergence or 25 iterations
])
def test_fused_lloyd_pass_equals_three_pass_training(g, n, D, M, K, iters):
    """ProductQuantizer.apply in sum mode: the fused pass (tca::tc_assign_kernel<.., true>) against the
    round-1 sequence assign -> update_fixed -> count_diff.  Same integer sums, same assignments, so the
    centroids, the iteration counts and the convergence flags are identical."""
    rng = np.random.default_rng(n + D)
    X = clustered(rng, n, D, centres=40)
    X[7] = X[3]
    pts = g.Matrix(X).device()
    out = {}
    for fused in (0, 1):
        reports = {}
        cfg = g.ProductQuantizerConfig(K, M, iters, update_mode=g.UPDATE_SUM,
                                       report=lambda r: reports.update(
                                           {i: (k.num_iterations, k.converged) for i, k in enumerate(r.kmeans_reports)}))
        g.set_option("train_fused", fused)
        try:
            launches0 = g.kernel_launches()
            out[fused] = (g.ProductQuantizer.train(pts, cfg).codebook(), dict(reports), g.kernel_launches() - launches0)
        finally:
            g.set_option("train_fused", 1)
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert out[0][1] == out[1][1]
    assert out[1][2] < out[0][2]          # fewer launches: no update / compare kernels


def test_fused_single_window_with_offset_and_duplicates(g, oracle):
    """KMeans.computeClusters on a column window that starts off the 16-byte grid, with duplicate rows
    (exact score ties -> lowest index) -- fused == three-pass, and the assignments of the result equal the
    oracle's for the returned centroids."""
    rng = np.random.default_rng(4)
    n, D = 12000, 23
    X = clustered(rng, n, D, centres=6)
    X[100:200] = X[0]
    out = []
    for fused in (0, 1):
        g.set_option("train_fused", fused)
        try:
            km, info = g.KMeans.compute_clusters(g.Vectors(g.Matrix(X), 5, 16),
                                                 g.KMeansConfig(64, 12, seed=3, update_mode=g.UPDATE_SUM),
                                                 return_info=True)
        finally:
            g.set_option("train_fused", 1)
        out.append((km.centroids.copy(), info))
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32)) and out[0][1] == out[1][1]


def test_fixed_sum_precision_is_relative_to_the_window_maximum(g):
    """the fixed-point sum keeps 2^-28 of the window's largest |x|, not of each value: with one
    outlier of 1e3 among coordinates of ~1e-3 the centroid error is bounded by 2^-29 * 1e3 per coordinate
    (absolute), which is what the header documents; without the outlier it is ~1e-9 relative."""
    rng = np.random.default_rng(2)
    n, dim, K = 20000, 8, 4
    X = (rng.normal(size=(n, dim)) * 1e-3).astype(np.float32)
    a = rng.integers(0, K, n).astype(np.int32)
    exact = np.stack([X[a == k].astype(np.float64).mean(axis=0) for k in range(K)])
    km = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(X)), a, g.UPDATE_SUM)
    # unit = 2^(e - 28) with 2^e > max|x| >= 2^(e-1): half a unit is at most 2^-28 max|x|
    assert np.abs(km.centroids - exact).max() <= 2.0 ** -28 * np.abs(X).max() + 1e-12
    Xo = X.copy()
    Xo[0, 0] = 1e3
    exact_o = np.stack([Xo[a == k].astype(np.float64).mean(axis=0) for k in range(K)])
    km_o = g.KMeans.from_assignment(K, dim, g.Vectors(g.Matrix(Xo)), a, g.UPDATE_SUM)
    err = np.abs(km_o.centroids - exact_o)
    assert err.max() <= 2.0 ** -28 * 1e3 + 0.25 * 2.0 ** -23              # the documented bound (+ the fp32 result rounding)
    assert err.max() > 1e-9                                                # ... and it is really that coarse
