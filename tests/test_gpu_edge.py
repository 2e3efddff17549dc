"""GPU parity on the reference's edge cases through EVERY scan implementation:
non-finite and huge values in queries and codebooks, k beyond the in-kernel list size (the reference's
recall harness defaults reach k = 1000, G/Tests.scala:53), k > N.

Contract for non-finite distances: the reference heap has no defined behaviour for
NaN -- `root > v` is false both ways (G/TopKHeap.scala:69-79), so what it keeps depends on the row
order; +inf is ordinary (`inf > inf` is false: a later row at +inf never displaces an earlier one).  The
canonical rule here is a TOTAL order: every NaN ranks after +inf, ties (including NaN with NaN) by id.
Under that rule the oracle's canonical mode and the GPU agree on ids; a NaN's bits are not compared
(x86 propagates payloads, the GPU returns the canonical quiet NaN)."""
import numpy as np
import pytest

from test_gpu_parity import IMPLS, build_index, clustered, impl_id, random_codebook

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def run_query(g, ix, Q, k, frm, until, impl, boot_rows=4096):
    impl, bits, words = impl
    g.set_option("scan_impl", impl)
    g.set_option("pruned_bits", bits)
    g.set_option("pruned_words", words)
    g.set_option("boot_rows", boot_rows)
    try:
        return ix.batch_query(k, Q, frm, until)
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
        g.set_option("pruned_bits", 0)
        g.set_option("pruned_words", 0)
        g.set_option("boot_rows", 0)


def assert_same(got, ids, ds, sz):
    """ids equal; distances bit-equal where finite or infinite, NaN where the oracle has NaN."""
    assert np.array_equal(got.size, sz)
    assert np.array_equal(got.keys, ids)
    nan = np.isnan(ds)
    assert np.array_equal(np.isnan(got.values), nan)
    assert np.array_equal(got.values[~nan].view(np.uint32), ds[~nan].view(np.uint32))


@pytest.mark.parametrize("impl", IMPLS)
def test_nonfinite_queries(g, oracle, impl):
    """NaN / +-inf / 1e38 coordinates in QUERIES: a NaN coordinate makes every table entry of its
    quantizer NaN (all distances NaN: the first k rows by id); inf or 1e38 make them +inf (squares
    overflow); the other queries of the same tile must be unaffected."""
    rng = np.random.default_rng(90)
    n, D, M, k = 70_000, 24, 4, 10
    X = clustered(rng, n, D, centres=30)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    Q = clustered(rng, 21, D, centres=30)
    Q[1, 3] = np.nan
    Q[2, 0] = np.inf
    Q[3, 7] = -np.inf
    Q[4, 5] = 1e38
    Q[5, :] = np.nan
    Q[6, 2] = -1e38
    Q[7, 1] = 3e19                     # squares to 9e38 > FLT_MAX: +inf through one quantizer only
    Q[20, 23] = np.nan                 # last query of a ragged tile
    got = run_query(g, ix, Q, k, 0, n, impl_id(g, impl))
    ids, ds, sz = oracle.pq_query(Q, cb, enc.codes, k, topk_mode=oracle.TOPK_CANONICAL)
    assert np.all(np.isnan(ds[1])) and np.all(np.isinf(ds[2])) and np.array_equal(ids[5], np.arange(k))
    assert_same(got, ids, ds, sz)


@pytest.mark.parametrize("impl", IMPLS)
def test_nonfinite_codebook_entries(g, oracle, impl):
    """NaN / inf / huge centroids: only rows whose code selects such a centroid get a NaN / +inf
    distance; they rank last (NaN after +inf), ties by id."""
    rng = np.random.default_rng(91)
    n, D, M, k = 66_000, 16, 4, 12
    X = clustered(rng, n, D, centres=20)
    cb = random_codebook(rng, X, M, 256)
    cb[0, 7, 1] = np.nan
    cb[1, 200, 0] = np.inf
    cb[2, 13, 2] = 1e38
    cb[3, 99, 3] = -np.inf
    codes = rng.integers(0, 256, (M, n)).astype(np.uint8)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    ix = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(n), codes))
    Q = clustered(rng, 9, D, centres=20)
    got = run_query(g, ix, Q, k, 0, n, impl_id(g, impl))
    ids, ds, sz = oracle.pq_query(Q, cb, codes, k, topk_mode=oracle.TOPK_CANONICAL)
    assert_same(got, ids, ds, sz)
    # an index in which EVERY row is non-finite for some queries: order is NaN last, by id
    codes2 = codes.copy()
    codes2[0, ::2] = 7                  # NaN through quantizer 0 on even rows
    codes2[1, 1::2] = 200               # +inf through quantizer 1 on odd rows
    ix2 = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(n), codes2))
    got2 = run_query(g, ix2, Q, k, 0, n, impl_id(g, impl))
    ids2, ds2, sz2 = oracle.pq_query(Q, cb, codes2, k, topk_mode=oracle.TOPK_CANONICAL)
    assert np.all(np.isinf(ds2)) and np.all(ids2 % 2 == 1)          # +inf rows come before NaN rows
    assert_same(got2, ids2, ds2, sz2)


@pytest.mark.parametrize("k", [129, 500, 1000, 2049, 5000])
def test_k_beyond_the_kernel_lists(g, oracle, k):
    """k > 128 (the in-kernel list size) through the automatic path; the reference's harness defaults
    reach 1000 (G/Tests.scala:53).  k > 2048 takes the full-sort selection."""
    rng = np.random.default_rng(k)
    n, D, M = 300_000, 32, 4
    X = clustered(rng, n, D, centres=50)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    Q = clustered(rng, 5, D, centres=50)
    got = ix.batch_query(k, Q, 7, n - 3)
    ids, ds, sz = oracle.pq_query(Q, cb, enc.codes, k, 7, n - 3, topk_mode=oracle.TOPK_LITERAL)
    # the literal heap returns the same distance multiset; canonicalise equal-distance groups by id
    for q in range(len(Q)):
        order = np.lexsort((ids[q], ds[q]))
        assert np.array_equal(got.values[q].view(np.uint32), ds[q][order].view(np.uint32))
        cut = ds[q][order][-1]
        inside = got.values[q] < cut                                 # rows tied at the cut may differ by id
        assert np.array_equal(got.keys[q][inside], ids[q][order][inside])
    assert np.all(got.size == k)


def test_k_larger_than_n_at_300k_rows(g, oracle):
    """k > N: every row comes back, (distance, id) ascending, the tail slots empty (TopKHeap(k) with
    k > N keeps everything, G/TopKHeap.scala:69-79)."""
    rng = np.random.default_rng(12)
    n, D, M, k = 300_000, 16, 2, 300_017
    X = clustered(rng, n, D, centres=50)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    ix = g.PQIndex(pq, enc)
    Q = clustered(rng, 2, D, centres=50)
    got = ix.batch_query(k, Q)
    lut = oracle.prepare_query(Q, cb)
    for q in range(2):
        d = np.zeros(n, np.float32)
        for m in range(M):
            d = d + lut[q, m][enc.codes[m]]                         # fp32, quantizer order: PQIndex.distances
        order = np.lexsort((np.arange(n), d))
        assert got.size[q] == n
        assert np.array_equal(got.keys[q, :n], order.astype(np.int32))
        assert np.array_equal(got.values[q, :n].view(np.uint32), d[order].view(np.uint32))
        assert np.all(got.keys[q, n:] == -1) and np.all(np.isinf(got.values[q, n:]))


def test_index_rejects_centroid_ids_beyond_k(g):
    """a code >= K must not reach a table read (the reference throws ArrayIndexOutOfBounds on
    the first lookup, G/Index.scala:401-406); device-side validation when the index is created."""
    import torch
    rng = np.random.default_rng(1)
    D, M, K, n = 8, 2, 100, 5000
    cb = rng.normal(size=(M, K, 4)).astype(np.float32)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    good = rng.integers(0, K, (M, n)).astype(np.uint8)
    g.PQIndex(pq, g.EncodedMatrix.from_planes(g.Coder8(n), good))
    bad = good.copy()
    bad[1, 4321] = K
    with pytest.raises(ValueError):
        g.PQIndex(pq, g.EncodedMatrix.from_planes(g.Coder8(n), bad))
    stride = (n + 15) // 16 * 16
    t = torch.zeros((M, stride), dtype=torch.uint8, device="cuda")
    t[:, :n] = torch.from_numpy(bad).cuda()
    with pytest.raises(ValueError):
        g.PQIndex.from_device_codes(pq, t, n)


@pytest.mark.parametrize("impl", ["fused", "pruned8", "pruned16", "pruned8w2"])
@pytest.mark.parametrize("k", [129, 300, 1000])
def test_k_chunked_passes_through_the_fast_kernels(g, oracle, impl, k):
    """k > 128 through the fused / pruned kernels in passes of <= 128 (each pass admits only keys greater
    than the last key of the previous one), heavy ties included: equal to the oracle's canonical top-k."""
    rng = np.random.default_rng(k)
    n, D, M = 280_000, 24, 3
    pq, cb, _, _ = build_index(g, rng, 10, D, M)
    codes = rng.integers(0, 256, (M, n)).astype(np.uint8)
    codes[:, 1000:1400] = codes[:, 999:1000]          # 400 rows with one code: a tie group larger than a pass
    ix = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(n), codes))
    Q = clustered(rng, 6, D)
    Q[0] = pq.decode(g.EncodedMatrix(g.Coder8(1), codes[:, 999:1000])).data[0]   # the tie group is the nearest
    got = run_query(g, ix, Q, k, 3, n - 5, impl_id(g, impl))
    ids, ds, sz = oracle.pq_query(Q, cb, codes, k, 3, n - 5, topk_mode=oracle.TOPK_CANONICAL)
    assert_same(got, ids, ds, sz)
