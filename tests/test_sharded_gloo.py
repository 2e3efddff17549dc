"""world_size-2 gloo tests of the multi-GPU host logic (run on CPU).

The device ops are replaced by the oracle here (tests may use it); what is under test is the
sharding plan, the exchange hooks handed to the library (gulon_comm_t) and the all-gather + merge
plumbing of ShardedPQIndex."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import ctypes as C
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gulon_b200.sharded import ShardedPQIndex, TorchComm, merge_topk_host, shard_bounds, shard_plan
        from oracle import oracle as o

        # --- exchange hooks, called exactly as the library calls them (raw pointers) ---
        comm = TorchComm(device="cpu")
        f = np.arange(10, dtype=np.float32) * (rank + 1)
        assert comm.struct.allreduce_sum_f32(None, f.ctypes.data, 10, None) == 0
        tri = world * (world + 1) // 2
        assert np.array_equal(f, np.arange(10, dtype=np.float32) * tri)
        i = np.full(5, rank + 1, np.int32)
        assert comm.struct.allreduce_sum_i32(None, i.ctypes.data, 5, None) == 0
        assert np.all(i == tri)
        send = np.full(8, rank + 7, np.uint8)
        recv = np.zeros(8 * world, np.uint8)
        assert comm.struct.allgather(None, send.ctypes.data, recv.ctypes.data, 8, None) == 0
        assert all(np.all(recv[8 * r:8 * r + 8] == 7 + r) for r in range(world))
        assert comm.struct.rank == rank and comm.struct.world == world

        # --- sharded scan: identical data on every rank, each scans its own shard ---
        rng = np.random.default_rng(0)
        n, D, M, K, k = 5003, 12, 3, 256, 10
        cb = rng.normal(size=(M, K, 4)).astype(np.float32)
        codes = rng.integers(0, 8, (M, n)).astype(np.uint8)   # few codes => ties across shards
        Q = rng.normal(size=(9, D)).astype(np.float32)
        wi, wd, wz = o.pq_query(Q, cb, codes, k)

        class OracleOps:
            def __init__(self, lo, hi):
                self.lo, self.hi = lo, hi

            def local_query(self, kk, queries, id_offset):
                ids, ds, sz = o.pq_query(queries.numpy(), cb, codes[:, self.lo:self.hi], kk)
                ids = np.where(ids >= 0, ids + id_offset, -1).astype(np.int32)
                return torch.from_numpy(ids), torch.from_numpy(ds), torch.from_numpy(sz)

            def merge(self, ids_all, ds_all, kk):
                return tuple(torch.from_numpy(a) for a in
                             merge_topk_host(ids_all.numpy(), ds_all.numpy(), kk))

        # every (row shards x query groups) factorisation of the world: plain row sharding (world, 1),
        # replicated planes with the query batch split (1, world), and the mixed plans in between
        plans = [(r, world // r) for r in range(1, world + 1) if world % r == 0]
        for R, Cq in plans:
            lo, hi = shard_bounds(n, R)[rank % R]
            sh = ShardedPQIndex(None, lo, ops=OracleOps(lo, hi), plan=(R, Cq))
            assert (sh.row_shard, sh.query_group) == (rank % R, rank // R)
            ids, ds, sz = sh.batch_query(k, torch.from_numpy(Q))      # 9 queries: ragged slices
            assert np.array_equal(ids.numpy(), wi), (R, Cq)
            assert np.array_equal(ds.numpy(), wd), (R, Cq)
            assert np.array_equal(sz.numpy(), wz), (R, Cq)
        assert shard_plan(10_000_000, 8) == (1, 8) and shard_plan(100_000_000, 8) == (8, 1)
        assert shard_plan(40_000_000, 8) == (4, 2) and shard_plan(5, 3) == (1, 3)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, "fail: " + traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from gulon_b200.sharded import shard_bounds
    for n, w in [(100_000_000, 8), (5003, 2), (10, 4), (0, 2), (16, 1), (17, 8)]:
        b = shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        for (lo, hi), (lo2, _) in zip(b[:-1], b[1:]):
            assert hi == lo2 and lo <= hi
        assert all(lo % 16 == 0 or lo == n for lo, _ in b)


def test_merge_topk_host_is_lexicographic():
    from gulon_b200.sharded import merge_topk_host
    ids = np.array([[[5, 9, -1]], [[2, 7, 8]]], np.int32)
    ds = np.array([[[1.0, 2.0, np.inf]], [[1.0, 2.0, 3.0]]], np.float32)
    i, d, s = merge_topk_host(ids, ds, 4)
    assert i.tolist() == [[2, 5, 7, 9]] and d.tolist() == [[1.0, 1.0, 2.0, 2.0]] and s.tolist() == [4]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_gloo_exchange_and_sharded_scan(oracle, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    res = sorted(out.get(timeout=10) for _ in range(world))
    assert res == [(r, "ok") for r in range(world)], res
    assert all(p.exitcode == 0 for p in procs)
