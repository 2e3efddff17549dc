"""gulon_pq_rerank_query[_dev] / gulon_rerank_dev: PQ top-R candidates + exact fp32 re-rank, single
GPU and row-sharded (ranks = host threads on one GPU, hooks over host barriers), against the
oracle: PQIndex.batchQuery for the candidates (G/Index.scala:414-440), exactNearestNeighbours /
distanceSq restricted to them (G/Index.scala:209-229, G/MathUtils.scala:85-95)."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_sharded import ThreadGroup, clustered, run_ranks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.skip("no CUDA device")
    return g


def oracle_rerank(o, X, Q, cb, codes, k, R):
    ci, _, _ = o.pq_query(Q, cb, codes, R, topk_mode=o.TOPK_CANONICAL)
    wi = np.full((len(Q), k), -1, np.int32)
    wd = np.full((len(Q), k), np.inf, np.float32)
    for q in range(len(Q)):
        cand = np.sort(ci[q][ci[q] >= 0])
        ei, ed, _ = o.exact_nn(X[cand], Q[q:q + 1], k, topk_mode=o.TOPK_CANONICAL)
        m = ei[0] >= 0
        wi[q, :m.sum()] = cand[ei[0][m]]
        wd[q, :m.sum()] = ed[0][m]
    return wi, wd


@pytest.mark.parametrize("n,D,M,k,R,world", [(50_000, 40, 8, 10, 100, 1), (50_000, 40, 8, 10, 100, 2),
                                             (30_000, 37, 5, 3, 128, 3), (400, 16, 4, 10, 64, 2),
                                             (300_000, 24, 6, 10, 50, 2)])
def test_rerank_query_vs_oracle(g, oracle, n, D, M, k, R, world):
    import torch
    from gulon_b200 import _native as N
    from gulon_b200.sharded import shard_bounds
    rng = np.random.default_rng(n + D + world)
    X = clustered(rng, n, D)
    X[5] = X[3]                                   # duplicate rows: equal exact distances, id order decides
    Q = clustered(rng, 23, D)
    pq = g.ProductQuantizer.train(g.Matrix(X[:8000]), g.ProductQuantizerConfig(256 if n > 1000 else 16, M, 3))
    enc = pq.encode(X)
    cb = pq.codebook()
    wi, wd = oracle_rerank(oracle, X, Q, cb, enc.codes, k, R)
    dev = torch.device("cuda", 0)
    bounds = shard_bounds(n, world)
    grp = ThreadGroup(world, dev)
    Qd = torch.from_numpy(Q).to(dev)

    def rank_fn(rank):
        lo, hi = bounds[rank]
        ix = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(hi - lo), np.ascontiguousarray(enc.codes[:, lo:hi])))
        pts = g.Matrix(np.ascontiguousarray(X[lo:hi])).device()
        rc, keep = grp.comm(rank) if world > 1 else (None, None)
        ids = torch.empty((len(Q), k), dtype=torch.int32, device=dev)
        ds = torch.empty((len(Q), k), dtype=torch.float32, device=dev)
        sz = torch.empty((len(Q),), dtype=torch.int32, device=dev)
        N.check(N.lib().gulon_pq_rerank_query_dev(ix.handle, pts.handle, C.byref(rc) if rc is not None else None,
                                                  Qd.data_ptr(), len(Q), D, k, R, 0, lo, ids.data_ptr(),
                                                  ds.data_ptr(), sz.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        hi_, hd, hs = np.empty((len(Q), k), np.int32), np.empty((len(Q), k), np.float32), np.empty(len(Q), np.int32)
        N.check(N.lib().gulon_pq_rerank_query(ix.handle, pts.handle, C.byref(rc) if rc is not None else None,
                                              Q.ctypes.data, len(Q), D, k, R, 0, lo, hi_.ctypes.data,
                                              hd.ctypes.data, hs.ctypes.data))
        return ids.cpu().numpy(), ds.cpu().numpy(), sz.cpu().numpy(), hi_, hd, hs

    out, errs = run_ranks(world, rank_fn)
    assert not errs, errs
    for ids, ds, sz, hi_, hd, hs in out:
        assert np.array_equal(ids, wi)
        assert np.array_equal(ds.view(np.uint32), wd.view(np.uint32))
        assert np.array_equal(sz, (wi >= 0).sum(1))
        assert np.array_equal(hi_, wi) and np.array_equal(hd.view(np.uint32), wd.view(np.uint32))


def test_rerank_dev_matches_host_rerank_and_oracle(g, oracle):
    """gulon_rerank_dev (fused kernel) == gulon_rerank (host form) == oracle distanceSq + top-k; odd D
    (scalar tail), unaligned leading dimension, -1 and out-of-shard candidates."""
    import torch
    from gulon_b200 import _native as N
    from gulon_b200.index import rerank
    rng = np.random.default_rng(5)
    n, D, R, k, nq = 5000, 301, 77, 10, 40
    X = clustered(rng, n, D)
    Q = clustered(rng, nq, D)
    cand = np.stack([rng.permutation(n)[:R] for _ in range(nq)]).astype(np.int32)
    cand[:, 5] = -1
    cand[3, :] = -1
    want = rerank(g.Matrix(X), Q, cand, k)
    for q in (0, 3, 17):
        c = np.sort(cand[q][cand[q] >= 0])
        if len(c) == 0:
            assert want.size[q] == 0 and np.all(want.keys[q] == -1)
            continue
        ei, ed, _ = oracle.exact_nn(X[c], Q[q:q + 1], k, topk_mode=oracle.TOPK_CANONICAL)
        assert np.array_equal(c[ei[0]], want.keys[q])
        assert np.array_equal(ed[0].view(np.uint32), want.values[q].view(np.uint32))
    dev = torch.device("cuda", 0)
    pts = g.Matrix(X).device()
    # the same rows seen as a shard that starts at global row 1000: candidates shift, some fall outside
    lo = 1000
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
    sz = torch.empty((nq,), dtype=torch.int32, device=dev)
    gc = torch.from_numpy(np.where(cand >= 0, cand + lo, -1).astype(np.int32)).to(dev)
    Qd = torch.from_numpy(Q).to(dev)
    N.check(N.lib().gulon_rerank_dev(pts.handle, Qd.data_ptr(), nq, D, gc.data_ptr(), R, k, lo, ids.data_ptr(),
                                     ds.data_ptr(), sz.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got_i = ids.cpu().numpy()
    assert np.array_equal(np.where(got_i >= 0, got_i - lo, -1), want.keys)
    assert np.array_equal(ds.cpu().numpy().view(np.uint32), want.values.view(np.uint32))
    assert np.array_equal(sz.cpu().numpy(), want.size)
