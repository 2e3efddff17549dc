"""gulon_pq_query_sharded[_dev] -- the multi-GPU PQIndex#batchQuery behind the C ABI -- driven through
its gulon_comm_t hooks by host THREADS that share one GPU (each thread is a rank with its own index
handle over its own row shard; the hooks exchange through host barriers).  The answer must equal the
oracle's PQIndex.batchQuery over the concatenated shards (G/Index.scala:414-440 + TopKHeap#merge,
G/TopKHeap.scala:44-53), ids and distance bits.  The NCCL wiring of the same entry point is
exercised by bench.py at world > 1 (also oracle-checked there)."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.skip("no CUDA device")
    return g


def clustered(rng, n, D, centres=64, scale=2.0, noise=0.4):
    c = rng.normal(size=(centres, D)).astype(np.float32) * scale
    return (c[rng.integers(0, centres, n)] + noise * rng.normal(size=(n, D))).astype(np.float32)


class ThreadGroup:
    """An all-gather among `size` threads: slots + a barrier, device buffers viewed through torch."""

    def __init__(self, size, dev):
        self.size, self.dev = size, dev
        self.bar = threading.Barrier(size)
        self.slots = [None] * size
        self.calls = 0
        self.fail = False

    def comm(self, rank):
        import torch
        from gulon_b200 import _native as N
        from gulon_b200.sharded import _view

        def allgather(_u, send, recv, nbytes, _stream):
            if self.fail:
                return 1
            try:
                s = _view(send, nbytes, "|u1", self.dev)
                r = _view(recv, nbytes * self.size, "|u1", self.dev)
                torch.cuda.synchronize()
                self.slots[rank] = s
                self.bar.wait()
                r.copy_(torch.cat([self.slots[i] for i in range(self.size)]))
                torch.cuda.synchronize()
                self.bar.wait()
                if rank == 0:
                    self.calls += 1
                return 0
            except Exception:
                self.bar.abort()
                return 1

        bad = lambda *a: 1
        cbs = (N.Comm.ALLREDUCE_F32(bad), N.Comm.ALLREDUCE_I32(bad), N.Comm.ALLGATHER(allgather))
        return N.Comm(rank, self.size, cbs[0], cbs[1], cbs[2], None), cbs


def run_ranks(world, fn):
    out, errs = [None] * world, []

    def body(rank):
        try:
            out[rank] = fn(rank)
        except Exception as e:  # pragma: no cover
            errs.append((rank, e))

    ths = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(300)
    return out, errs


@pytest.mark.parametrize("R,Cg,nq,k", [(2, 1, 37, 10), (3, 1, 5, 7), (2, 2, 41, 10), (1, 2, 9, 3)])
def test_sharded_query_threads_vs_oracle(g, oracle, R, Cg, nq, k):
    import torch
    from gulon_b200 import _native as N
    from gulon_b200.sharded import shard_bounds
    rng = np.random.default_rng(100 + 10 * R + Cg)
    n, D, M = 60_000, 24, 6
    X = clustered(rng, n, D)
    Q = clustered(rng, nq, D)
    pq = g.ProductQuantizer.train(g.Matrix(X[:8000]), g.ProductQuantizerConfig(256, M, 3))
    enc = pq.encode(X)
    cb = pq.codebook()
    want_i, want_d, _ = oracle.pq_query(Q, cb, enc.codes, k, topk_mode=oracle.TOPK_CANONICAL)
    dev = torch.device("cuda", 0)
    world = R * Cg
    bounds = shard_bounds(n, R)
    row_groups = [ThreadGroup(R, dev) for _ in range(Cg)]
    col_groups = [ThreadGroup(Cg, dev) for _ in range(R)]
    Qd = torch.from_numpy(Q).to(dev)

    def rank_fn(rank):
        shard, group = rank % R, rank // R
        lo, hi = bounds[shard]
        ix = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(hi - lo), np.ascontiguousarray(enc.codes[:, lo:hi])))
        rc, keep1 = row_groups[group].comm(shard) if R > 1 else (None, None)
        qc, keep2 = col_groups[shard].comm(group) if Cg > 1 else (None, None)
        ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
        sz = torch.empty((nq,), dtype=torch.int32, device=dev)
        N.check(N.lib().gulon_pq_query_sharded_dev(
            ix.handle, C.byref(rc) if rc is not None else None, C.byref(qc) if qc is not None else None,
            Qd.data_ptr(), nq, D, k, 0, lo, ids.data_ptr(), ds.data_ptr(), sz.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        # the host form: host queries in, host answers out
        hi_, hd, hs = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32), np.empty(nq, np.int32)
        N.check(N.lib().gulon_pq_query_sharded(
            ix.handle, C.byref(rc) if rc is not None else None, C.byref(qc) if qc is not None else None,
            Q.ctypes.data, nq, D, k, 0, lo, hi_.ctypes.data, hd.ctypes.data, hs.ctypes.data))
        return ids.cpu().numpy(), ds.cpu().numpy(), sz.cpu().numpy(), hi_, hd, hs

    out, errs = run_ranks(world, rank_fn)
    assert not errs, errs
    for ids, ds, sz, hi_, hd, hs in out:          # every rank holds the whole, identical answer
        assert np.array_equal(ids, want_i)
        assert np.array_equal(ds.view(np.uint32), want_d.view(np.uint32))
        assert np.all(sz == min(k, n))
        assert np.array_equal(hi_, want_i) and np.array_equal(hd.view(np.uint32), want_d.view(np.uint32))
        assert np.array_equal(hs, sz)
    if R > 1:
        assert row_groups[0].calls == 2               # ONE exchange per query batch (dev + host call)


def test_sharded_query_short_shard_and_k_larger_than_shard(g, oracle):
    """A shard with fewer than k rows contributes what it has; empty slots are id -1 / +inf."""
    import torch
    from gulon_b200 import _native as N
    rng = np.random.default_rng(7)
    n, D, M, k, nq = 40, 8, 2, 16, 6
    X = clustered(rng, n, D, centres=4)
    Q = clustered(rng, nq, D, centres=4)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(16, M, 2))
    enc = pq.encode(X)
    want_i, want_d, _ = oracle.pq_query(Q, pq.codebook(), enc.codes, k, topk_mode=oracle.TOPK_CANONICAL)
    dev = torch.device("cuda", 0)
    bounds = [(0, 32), (32, 40)]
    grp = ThreadGroup(2, dev)

    def rank_fn(rank):
        lo, hi = bounds[rank]
        ix = g.PQIndex(pq, g.EncodedMatrix(g.Coder8(hi - lo), np.ascontiguousarray(enc.codes[:, lo:hi])))
        rc, keep = grp.comm(rank)
        ids, ds, sz = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32), np.empty(nq, np.int32)
        N.check(N.lib().gulon_pq_query_sharded(ix.handle, C.byref(rc), None, Q.ctypes.data, nq, D, k, 0, lo,
                                               ids.ctypes.data, ds.ctypes.data, sz.ctypes.data))
        return ids, ds, sz

    out, errs = run_ranks(2, rank_fn)
    assert not errs, errs
    for ids, ds, sz in out:
        assert np.array_equal(ids, want_i)
        assert np.array_equal(ds.view(np.uint32), want_d.view(np.uint32))
        assert np.all(sz == 16)


def test_sharded_query_hook_failure_is_ecomm(g):
    import torch
    from gulon_b200 import _native as N
    rng = np.random.default_rng(3)
    X = clustered(rng, 2000, 8, centres=4)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(16, 2, 2))
    ix = g.PQIndex(pq, pq.encode(X))
    grp = ThreadGroup(2, torch.device("cuda", 0))
    grp.fail = True
    rc, keep = grp.comm(0)
    ids, ds = np.empty((3, 4), np.int32), np.empty((3, 4), np.float32)
    with pytest.raises(N.GulonError) as e:
        N.check(N.lib().gulon_pq_query_sharded(ix.handle, C.byref(rc), None, X.ctypes.data, 3, 8, 4, 0, 0,
                                               ids.ctypes.data, ds.ctypes.data, None))
    assert e.value.code == N.ECOMM
    with pytest.raises(ValueError):   # require(...): global ids must stay Int
        N.check(N.lib().gulon_pq_query_sharded(ix.handle, None, None, X.ctypes.data, 3, 8, 4, 0, 2**31 - 100,
                                               ids.ctypes.data, ds.ctypes.data, None))


def test_two_devices_one_process(g, oracle):
    """the > 48 KB shared-memory opt-in is per device.  One process, one thread per GPU
    (gulon_set_device), each with its own handles: train, encode, pruned + fused scan, merge."""
    import torch
    from gulon_b200 import _native as N
    if g.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    devs = (C.c_int32 * 2)(0, 1)
    N.check(N.lib().gulon_init(devs, 2))
    rng = np.random.default_rng(11)
    n, D, M, k = 300_000, 32, 8, 10
    X = clustered(rng, n, D)
    Q = clustered(rng, 20, D)

    def rank_fn(rank):
        N.check(N.lib().gulon_set_device(rank))
        # sum mode: tc_assign + update_fixed (both need the opt-in); then the running mean kernels
        pq = g.ProductQuantizer.train(g.Matrix(X[:20000]),
                                      g.ProductQuantizerConfig(256, M, 3, update_mode=g.UPDATE_SUM))
        g.ProductQuantizer.train(g.Matrix(X[:5000]), g.ProductQuantizerConfig(256, M, 1))
        enc = pq.encode(X)
        ix = g.PQIndex(pq, enc)
        r = ix.batch_query(k, Q)
        return pq.codebook(), enc.codes, r.keys, r.values

    out, errs = run_ranks(2, rank_fn)
    N.check(N.lib().gulon_set_device(0))
    assert not errs, errs
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    cb, codes, keys, vals = out[1]
    wi, wd, _ = oracle.pq_query(Q, cb, codes, k, topk_mode=oracle.TOPK_CANONICAL)
    assert np.array_equal(keys, wi) and np.array_equal(vals.view(np.uint32), wd.view(np.uint32))
