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
        from gulon_b200.sharded import ShardedPQIndex, TorchComm, merge_topk_host, shard_bounds
        from oracle import oracle as o

        # --- exchange hooks, called exactly as the library calls them (raw pointers) ---
        comm = TorchComm(device="cpu")
        f = np.arange(10, dtype=np.float32) * (rank + 1)
        assert comm.struct.allreduce_sum_f32(None, f.ctypes.data, 10, None) == 0
        assert np.array_equal(f, np.arange(10, dtype=np.float32) * 3)
        i = np.full(5, rank + 1, np.int32)
        assert comm.struct.allreduce_sum_i32(None, i.ctypes.data, 5, None) == 0
        assert np.all(i == 3)
        send = np.full(8, rank + 7, np.uint8)
        recv = np.zeros(16, np.uint8)
        assert comm.struct.allgather(None, send.ctypes.data, recv.ctypes.data, 8, None) == 0
        assert np.all(recv[:8] == 7) and np.all(recv[8:] == 8)
        assert comm.struct.rank == rank and comm.struct.world == world

        # --- sharded scan: identical data on every rank, each scans its own shard ---
        rng = np.random.default_rng(0)
        n, D, M, K, k = 5003, 12, 3, 256, 10
        cb = rng.normal(size=(M, K, 4)).astype(np.float32)
        codes = rng.integers(0, 8, (M, n)).astype(np.uint8)   # few codes => ties across shards
        Q = rng.normal(size=(9, D)).astype(np.float32)
        lo, hi = shard_bounds(n, world)[rank]

        class OracleOps:
            def local_query(self, kk, queries, id_offset):
                ids, ds, sz = o.pq_query(queries.numpy(), cb, codes[:, lo:hi], kk)
                ids = np.where(ids >= 0, ids + id_offset, -1).astype(np.int32)
                return torch.from_numpy(ids), torch.from_numpy(ds), torch.from_numpy(sz)

            def merge(self, ids_all, ds_all, kk):
                return tuple(torch.from_numpy(a) for a in
                             merge_topk_host(ids_all.numpy(), ds_all.numpy(), kk))

        sh = ShardedPQIndex(None, lo, ops=OracleOps())
        ids, ds, sz = sh.batch_query(k, torch.from_numpy(Q))
        wi, wd, wz = o.pq_query(Q, cb, codes, k)
        assert np.array_equal(ids.numpy(), wi)
        assert np.array_equal(ds.numpy(), wd)
        assert np.array_equal(sz.numpy(), wz)
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
def test_two_rank_gloo_exchange_and_sharded_scan(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    res = sorted(out.get(timeout=10) for _ in range(2))
    assert res == [(0, "ok"), (1, "ok")], res
    assert all(p.exitcode == 0 for p in procs)
