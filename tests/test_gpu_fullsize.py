"""GPU: BASELINE.json's full shapes (c1 1M x 100, c2 10M x 300, a c4 shard 12.5M x 128, c5 1M x 1000).

Against the ORACLE at full size (`check_against_oracle`): PQIndex.batchQuery of 8 sampled queries over
the whole index (ids and distance bits), ProductQuantizer#encode of a 50 000-row slice taken from the
middle of the data set (codes byte-equal; the rows are regenerated on host cores by the CPU twin of
the data generator, oracle/synth.c), and the trained codebook of the index itself (bit-equal to the
oracle's ProductQuantizer.apply on the same training rows).

And through size-independent properties over thousands of queries: pruned scan == exact fused scan ==
any-k selection scan, range split + merge == whole scan (TopKHeap#merge, T/TopKHeapSpec.scala:33-52),
a decoded row finds its own code at distance 0 (T/IndexSpec.scala:62-73 analogue), encode is
idempotent on decoded rows and independent of the kernel (tensor-core filter vs exact CUDA cores),
(distance, id) order of every result.  Data are generated on the device (gulon_b200.synth)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def build(g, rows, D, M, centres=4096, nonneg=False, train_rows=131072, iters=4):
    import torch
    from gulon_b200 import _native as N
    from gulon_b200.synth import Mixture
    dev = torch.device("cuda", 0)
    mix = Mixture(D, device=dev, centres=centres, nonneg=nonneg, span=40.0 if nonneg else 1.0)
    xt = mix.rows(0, train_rows)
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(256, M, iters))
    del xt
    stride = (rows + 15) // 16 * 16
    codes = torch.zeros((M, stride), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    CH = 1 << 20
    for r0 in range(0, rows, CH):
        n = min(CH, rows - r0)
        x = mix.rows(r0, r0 + n)
        N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST,
                                            codes.data_ptr() + r0, stride, st))
        torch.cuda.synchronize()
        del x
    return mix, pq, codes, g.PQIndex.from_device_codes(pq, codes, rows)


def check_against_oracle(g, oracle, mix, pq, codes, rows, D, Q, ids, ds, k=10, train_rows=None, iters=None):
    """The oracle over the FULL index on a query sample, over a mid-data-set slice for encode, and
    (when the training protocol is given) for the codebook itself."""
    T = oracle.host_cores()
    cb = pq.codebook()
    hc = codes[:, :rows].cpu().numpy()
    nq = Q.shape[0]
    pick = np.unique(np.linspace(0, nq - 1, 8).astype(np.int64))
    qs = Q[pick].cpu().numpy()
    oi, od, _ = oracle.pq_query(qs, cb, hc, k, topk_mode=oracle.TOPK_CANONICAL, nthreads=T)
    assert np.array_equal(oi, ids[pick])
    assert np.array_equal(od.view(np.uint32), ds[pick].view(np.uint32))
    twin = oracle.SynthMixture(D, **mix.kw)
    lo = (rows // 2) | 1                                    # an odd offset: unaligned chunk starts
    x = twin.rows(lo, lo + 50_000, nthreads=T)
    assert np.array_equal(oracle.pq_encode(x, cb, tie_mode=oracle.TIE_LOWEST, nthreads=T), hc[:, lo:lo + 50_000])
    if train_rows:
        xt = twin.rows(0, train_rows, nthreads=T)
        want, _, _ = oracle.pq_train(xt, cb.shape[0], 256, iters, tie_mode=oracle.TIE_LOWEST, nthreads=T)
        assert np.array_equal(want.view(np.uint32), cb.view(np.uint32))


def query(g, ix, k, Q, impl=None, frm=0, until=None, **opts):
    if impl is not None:
        g.set_option("scan_impl", impl)
    for name, v in opts.items():
        g.set_option(name, v)
    try:
        r = ix.batch_query_dev(k, Q, frm, until)
        return r[0].cpu().numpy(), r[1].cpu().numpy()
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
        for name in opts:
            g.set_option(name, 0 if name != "pruned_stage_div" else 32)


def check_scan_properties(g, mix, pq, codes, ix, rows, k=10, nq=2400, n_exact=96):
    import torch
    Q = mix.rows(0, nq, stream_seed=1)
    ids, ds = query(g, ix, k, Q)                                   # automatic: pruned, feedback, two stages
    assert ids.shape == (nq, k) and ids.min() >= 0 and ids.max() < rows
    # ascending (distance, id) and no duplicates
    assert np.all(np.diff(ds, axis=1) >= 0)
    tie = np.diff(ds, axis=1) == 0
    assert np.all(np.diff(ids, axis=1)[tie] > 0)
    assert all(len(set(r)) == k for r in ids[:200])
    # exact kernel and the any-k selection path agree bit for bit on a sample
    fi, fd = query(g, ix, k, Q[:n_exact], impl=g.SCAN_FUSED)
    assert np.array_equal(fi, ids[:n_exact]) and np.array_equal(fd.view(np.uint32), ds[:n_exact].view(np.uint32))
    si, sd = query(g, ix, k, Q[:8], impl=g.SCAN_SIMPLE)
    assert np.array_equal(si, ids[:8]) and np.array_equal(sd.view(np.uint32), ds[:8].view(np.uint32))
    # the knobs change the work, not the answer
    for opts in ({"pruned_lb_quantizers": 1000}, {"pruned_lb_quantizers": max(2, len(pq.quantizers) // 3)},
                 {"pruned_stage_div": 0}, {"pruned_rowcodes": 0}):
        if "pruned_rowcodes" in opts:
            g.set_option("pruned_rowcodes", 0)
            try:
                oi, od = query(g, ix, k, Q[:512])
            finally:
                g.set_option("pruned_rowcodes", 1)
        else:
            oi, od = query(g, ix, k, Q[:512], impl=g.SCAN_PRUNED, **opts)
        assert np.array_equal(oi, ids[:512]) and np.array_equal(od.view(np.uint32), ds[:512].view(np.uint32)), opts
    # split + merge == whole (ragged cut)
    cut = rows // 3 + 5
    ai, ad = query(g, ix, k, Q[:512], until=cut)
    bi, bd = query(g, ix, k, Q[:512], frm=cut)
    mi = np.concatenate([ai, bi], axis=1)
    md = np.concatenate([ad, bd], axis=1)
    for q in range(512):
        o = np.lexsort((mi[q], md[q]))[:k]
        assert np.array_equal(mi[q][o], ids[q]) and np.array_equal(md[q][o], ds[q])
    # a decoded row is at distance 0 from its own code; the lowest id among identical codes wins
    pick = torch.from_numpy(np.random.default_rng(3).integers(0, rows, 64)).cuda()
    own = codes[:, pick].cpu().numpy()
    dec = pq.decode(g.EncodedMatrix.from_planes(g.Coder8(64), own)).data
    zi, zd = query(g, ix, 1, torch.from_numpy(dec).cuda())
    assert np.all(zd[:, 0] == 0.0)
    assert np.all(zi[:, 0] <= pick.cpu().numpy())
    assert np.array_equal(codes[:, torch.from_numpy(zi[:, 0].astype(np.int64)).cuda()].cpu().numpy(), own)
    return Q, ids, ds


def check_encode_properties(g, mix, pq, codes, rows, D):
    import torch
    from gulon_b200 import _native as N
    n = min(rows, 1 << 20)
    x = mix.rows(rows - n, rows)
    st = torch.cuda.current_stream().cuda_stream
    stride = (n + 15) // 16 * 16
    M = len(pq.quantizers)
    out = {}
    for name, impl in (("tc", 2), ("exact", 1)):
        g.set_option("assign_impl", impl)
        try:
            c = torch.zeros((M, stride), dtype=torch.uint8, device=x.device)
            N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST, c.data_ptr(), stride, st))
            torch.cuda.synchronize()
            out[name] = c[:, :n]
        finally:
            g.set_option("assign_impl", 0)
    assert torch.equal(out["tc"], out["exact"])
    assert torch.equal(out["tc"], codes[:, rows - n:rows])          # what the index holds
    # idempotence: encode(decode(codes)) == codes on rows whose codes are unambiguous -- decoded rows sit
    # exactly on their centroids, so the nearest centroid is the code unless another centroid coincides
    sub = out["tc"][:, :4096].cpu().numpy()
    dec = pq.decode(g.EncodedMatrix.from_planes(g.Coder8(4096), sub))
    again = pq.encode(dec).codes
    cb = pq.codebook()
    for m in range(M):
        diff = np.flatnonzero(again[m] != sub[m])
        for i in diff:                       # only (numerically) coincident centroids may differ
            a, b = cb[m, again[m, i]].astype(np.float64), cb[m, sub[m, i]].astype(np.float64)
            assert ((a - b) ** 2).sum() <= 1e-5 * ((a ** 2).sum() + (b ** 2).sum()), (m, i)


def test_c1_shape_1m_x_100_iid(g, oracle):
    """configs[0]: iid unit-variance rows (no cluster structure: the lower bound prunes least here)."""
    rows, D, M = 1_000_000, 100, 10
    mix, pq, codes, ix = build(g, rows, D, M, centres=0, train_rows=32768, iters=2)
    Q, ids, ds = check_scan_properties(g, mix, pq, codes, ix, rows, nq=1200, n_exact=64)
    check_against_oracle(g, oracle, mix, pq, codes, rows, D, Q, ids, ds, train_rows=32768, iters=2)


def test_c2_shape_10m_x_300(g, oracle):
    rows, D, M = 10_000_000, 300, 30
    mix, pq, codes, ix = build(g, rows, D, M)
    Q, ids, ds = check_scan_properties(g, mix, pq, codes, ix, rows)
    check_against_oracle(g, oracle, mix, pq, codes, rows, D, Q, ids, ds)
    check_encode_properties(g, mix, pq, codes, rows, D)


def test_c5_shape_1m_x_1000_with_rerank(g, oracle):
    import torch
    from gulon_b200.index import rerank
    rows, D, M = 1_000_000, 1000, 100
    mix, pq, codes, ix = build(g, rows, D, M, train_rows=65536, iters=3)
    Q, ids, ds = check_scan_properties(g, mix, pq, codes, ix, rows, nq=1200, n_exact=64)
    check_against_oracle(g, oracle, mix, pq, codes, rows, D, Q, ids, ds)
    check_encode_properties(g, mix, pq, codes, rows, D)
    # re-rank (c5): exact distances of 100 PQ candidates; the re-ranked top-10 are the 10 smallest exact
    # distances among the candidates, ascending, and the exact nearest neighbour of a database row is itself
    X = mix.rows(0, rows)
    pts = g.DevicePoints.from_torch(X)
    qs = Q[:32].cpu().numpy()
    cand, _ = query(g, ix, 100, Q[:32])
    rr = rerank(pts, qs, cand.astype(np.int32), 10)
    for q in range(32):
        # the oracle's exactNearestNeighbours over the candidate rows (MathUtils.distanceSq, sequential fp32)
        c = np.sort(cand[q])
        xh = X[torch.from_numpy(c.astype(np.int64)).cuda()].cpu().numpy()
        ei, ed, _ = oracle.exact_nn(xh, qs[q:q + 1], 10, topk_mode=oracle.TOPK_CANONICAL)
        assert np.array_equal(c[ei[0]], rr.keys[q])
        assert np.array_equal(ed[0].view(np.uint32), rr.values[q].view(np.uint32))
    self_rows = np.array([5, 77_777, rows - 1])
    nn = g.exact_nearest_neighbours(pts, X[torch.from_numpy(self_rows).cuda()].cpu().numpy(), 1)
    assert np.array_equal(nn.keys[:, 0], self_rows) and np.all(nn.values[:, 0] == 0.0)


def test_c4_shard_shape_12m5_x_128(g, oracle):
    rows, D, M = 12_500_000, 128, 16
    mix, pq, codes, ix = build(g, rows, D, M, centres=16384, nonneg=True)
    Q, ids, ds = check_scan_properties(g, mix, pq, codes, ix, rows, nq=1200, n_exact=64)
    check_against_oracle(g, oracle, mix, pq, codes, rows, D, Q, ids, ds)
