"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The reference is a single JVM; what shards here follows its own structure:
  * scan + top-k: database rows are independent and the reference already merges per-range heaps
    (TopKHeap#merge, G/TopKHeap.scala:44-53; GroupedIndex.query, G/Index.scala:273-281).  Each
    rank owns a contiguous row shard of every code plane, emits [Q][k] (distance, global id) and one
    all-gather + (distance, id) merge yields the same answer on every rank;
  * encode: rows independent, no exchange;
  * k-means: rows sharded; one all-reduce of per-cluster sums and counts per Lloyd iteration
    (done inside the library through the gulon_comm_t hooks built by `TorchComm`).
"""
import ctypes as C

import numpy as np

from . import _native as N


def shard_bounds(n_total, world, align=16):
    """Contiguous row shards [lo, hi) per rank; interior boundaries are multiples of `align` rows
    so that every shard's code planes start 16-byte aligned."""
    if world < 1 or n_total < 0:
        raise ValueError("need world >= 1 and n_total >= 0")
    per = -(-n_total // world)
    per = -(-per // align) * align
    return [(min(r * per, n_total), min((r + 1) * per, n_total)) for r in range(world)]


class _RawBuffer:
    """Zero-copy view of a raw pointer for torch.as_tensor / numpy."""

    def __init__(self, ptr, n, typestr, cuda):
        iface = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                 "strides": None}
        if cuda:
            self.__cuda_array_interface__ = iface
        else:
            self.__array_interface__ = iface


def _view(ptr, n, typestr, device):
    import torch
    if device.type == "cuda":
        return torch.as_tensor(_RawBuffer(ptr, n, typestr, True), device=device)
    return torch.from_numpy(np.asarray(_RawBuffer(ptr, n, typestr, False)))


class TorchComm:
    """gulon_comm_t whose hooks call torch.distributed collectives on the buffers the library
    hands over (device pointers with NCCL; host pointers with gloo in the CPU tests)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        self.device = torch.device(device)
        self.calls = {"allreduce_f32": 0, "allreduce_i32": 0, "allgather": 0}

        def ar_f32(_u, buf, n, stream):
            return self._allreduce(buf, n, "<f4", "allreduce_f32", stream)

        def ar_i32(_u, buf, n, stream):
            return self._allreduce(buf, n, "<i4", "allreduce_i32", stream)

        def ag(_u, send, recv, nbytes, stream):
            try:
                with self._on(stream):
                    s = _view(send, nbytes, "|u1", self.device)
                    r = _view(recv, nbytes * self.world, "|u1", self.device)
                    dist.all_gather_into_tensor(r, s, group=self.group)
                self.calls["allgather"] += 1
                self.bytes["allgather"] += int(nbytes)
                return 0
            except Exception:  # surfaced by the library as GULON_ECOMM "hook failed"
                return 1

        def ar_i64(_u, buf, n, stream):
            return self._allreduce(buf, n, "<i8", "allreduce_i64", stream)

        def ar_max(_u, buf, n, stream):
            return self._allreduce(buf, n, "<f4", "allreduce_max_f32", stream, op=dist.ReduceOp.MAX)

        self.calls["allreduce_i64"] = 0
        self.calls["allreduce_max_f32"] = 0
        self.bytes = {k: 0 for k in self.calls}      # payload per rank, summed over the calls
        self._cbs = (N.Comm.ALLREDUCE_F32(ar_f32), N.Comm.ALLREDUCE_I32(ar_i32), N.Comm.ALLGATHER(ag),
                     N.Comm.ALLREDUCE_I64(ar_i64), N.Comm.ALLREDUCE_MAX_F32(ar_max))
        self.struct = N.Comm(self.rank, self.world, self._cbs[0], self._cbs[1], self._cbs[2], None,
                             self._cbs[3], self._cbs[4])

    def _on(self, stream):
        """The hooks must enqueue on the library's stream: make it torch's current stream for the call
        (NCCL orders its own stream against the current one).  Host buffers (gloo) need nothing."""
        import contextlib
        import torch
        if self.device.type != "cuda":
            return contextlib.nullcontext()
        cur = torch.cuda.current_stream(self.device)
        if int(stream or 0) == cur.cuda_stream:
            return contextlib.nullcontext()
        return torch.cuda.stream(torch.cuda.ExternalStream(int(stream or 0), device=self.device))

    def _allreduce(self, buf, n, typestr, name, stream=None, op=None):
        try:
            if n > 0:
                with self._on(stream):
                    t = _view(buf, n, typestr, self.device)
                    self.dist.all_reduce(t, op=op if op is not None else self.dist.ReduceOp.SUM,
                                         group=self.group)
            self.calls[name] += 1
            self.bytes[name] += int(n) * int(typestr[-1])
            return 0
        except Exception:
            return 1


def merge_topk_host(ids, dists, k):
    """(distance, id) merge of [S][Q][k] result sets on the host -- used only to cross-check the
    device merge in tests and by the gloo tests' stand-in ops."""
    S, Q, kk = ids.shape
    out_i = np.full((Q, k), -1, np.int32)
    out_d = np.full((Q, k), np.inf, np.float32)
    out_s = np.zeros(Q, np.int32)
    for q in range(Q):
        i = ids[:, q, :].reshape(-1)
        d = dists[:, q, :].reshape(-1)
        keep = i >= 0
        i, d = i[keep], d[keep]
        o = np.lexsort((i, d))[:k]
        out_i[q, :len(o)] = i[o]
        out_d[q, :len(o)] = d[o]
        out_s[q] = len(o)
    return out_i, out_d, out_s


class NativeOps:
    """The device ops of the sharded scan: the CUDA library (the only product implementation).
    `ShardedPQIndex` drives it through ONE C entry point, gulon_pq_query_sharded[_dev] (local scan ->
    allgather hook -> (distance, id) merge -> allgather of the query slices), so a JVM host gets the
    same multi-GPU path without Python; `local_query` / `merge` remain for callers that compose the
    steps themselves."""

    def __init__(self, index):
        import torch
        self.index = index
        self.device = torch.device("cuda", torch.cuda.current_device())

    def sharded_query(self, k, queries, row_comm, query_comm, row_offset, normalize=False):
        """queries: CUDA float32 [Q][D] tensor -> (ids, dists, sizes) CUDA tensors; or a host array /
        CPU tensor -> numpy arrays (only this rank's query slice crosses PCIe)."""
        import torch
        rc = C.byref(row_comm.struct) if row_comm is not None else None
        qc = C.byref(query_comm.struct) if query_comm is not None else None
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            q = queries
            nq = q.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int32, device=q.device)
            ds = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            sz = torch.empty((nq,), dtype=torch.int32, device=q.device)
            ld = q.stride(0) if nq > 1 else max(q.shape[1], 1)
            N.check(N.lib().gulon_pq_query_sharded_dev(
                self.index.handle, rc, qc, q.data_ptr(), nq, ld, k, int(bool(normalize)), row_offset,
                ids.data_ptr(), ds.data_ptr(), sz.data_ptr(), torch.cuda.current_stream(q.device).cuda_stream))
            return ids, ds, sz
        q = queries.numpy() if isinstance(queries, torch.Tensor) else np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, k), np.int32)
        ds = np.empty((nq, k), np.float32)
        sz = np.empty((nq,), np.int32)
        N.check(N.lib().gulon_pq_query_sharded(
            self.index.handle, rc, qc, q.ctypes.data, nq, q.strides[0] // 4 if nq > 1 else q.shape[1], k,
            int(bool(normalize)), row_offset, ids.ctypes.data, ds.ctypes.data, sz.ctypes.data))
        return ids, ds, sz

    def local_query(self, k, queries, id_offset):
        return self.index.batch_query_dev(k, queries, id_offset=id_offset)

    def merge(self, ids_all, dists_all, k):
        import torch
        S, Q, _ = ids_all.shape
        oi = torch.empty((Q, k), dtype=torch.int32, device=ids_all.device)
        od = torch.empty((Q, k), dtype=torch.float32, device=ids_all.device)
        oz = torch.empty((Q,), dtype=torch.int32, device=ids_all.device)
        N.check(N.lib().gulon_topk_merge_dev(ids_all.data_ptr(), dists_all.data_ptr(), S, Q, k,
                                             oi.data_ptr(), od.data_ptr(), oz.data_ptr(),
                                             torch.cuda.current_stream(ids_all.device).cuda_stream))
        return oi, od, oz


def shard_plan(n_rows, world, min_rows_per_shard=8_000_000):
    """(R, C) with R * C == world: R row shards of the code planes x C query groups.

    The pruned scan is most efficient on long row ranges (its thresholds tighten with the rows a CTA has
    seen, and the per-launch table builds amortise over the range), and code planes are small (M bytes
    per row: 300 MB at 10M x m=30), so an index is split into as many row shards as keep
    `min_rows_per_shard` rows each and the remaining factor of the world size splits the QUERY batch:
    10M rows on 8 GPUs -> 1 shard x 8 query groups (every rank holds the planes, scans 1/8 of the queries);
    100M rows on 8 GPUs -> 8 row shards x 1; 40M rows on 8 GPUs -> 4 x 2."""
    if world < 1:
        raise ValueError("need world >= 1")
    R = 1
    for r in range(1, world + 1):
        if world % r == 0 and n_rows // r >= min_rows_per_shard:
            R = r
    return R, world // R


class ShardedPQIndex:
    """A PQIndex spread over the ranks of the default process group by a (R, C) plan (`shard_plan`):
    rank = query_group * R + row_shard.  The ranks of a query group hold the R contiguous row shards of
    the code planes; each scans its shard for the group's slice of the query batch, one all-gather +
    (distance, id) merge inside the group gives the slice's answer, and one all-gather across the groups
    assembles the batch: identical results on every rank.  plan = (world, 1) is plain row sharding.

    `ops` provides `local_query(k, queries, id_offset)` -> (ids, dists, sizes) tensors and
    `merge(ids_all [S][Q][k], dists_all, k)`; the default is the CUDA library.
    """

    def __init__(self, local_index, row_offset, group=None, ops=None, plan=None):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.R, self.C = plan if plan is not None else (self.world, 1)
        if self.R * self.C != self.world:
            raise ValueError("plan %r does not factor the world size %d" % ((self.R, self.C), self.world))
        self.row_shard, self.query_group = self.rank % self.R, self.rank // self.R
        self.row_group = group      # ranks holding the shards of one index copy
        self.col_group = None       # ranks holding the same shard, one per query group
        if self.C > 1:
            if group is not None:
                raise ValueError("a (R, C) plan with C > 1 spans the default process group")
            # every rank creates every group, in the same order
            rows = [dist.new_group(ranks=[c * self.R + i for i in range(self.R)]) for c in range(self.C)]
            cols = [dist.new_group(ranks=[c * self.R + i for c in range(self.C)]) for i in range(self.R)]
            self.row_group = rows[self.query_group]
            self.col_group = cols[self.row_shard]
        self.row_offset = int(row_offset)
        self.ops = ops if ops is not None else NativeOps(local_index)
        self.row_comm = self.col_comm = None
        if isinstance(self.ops, NativeOps) and self.world > 1:
            # exchange hooks of the C entry point, one gulon_comm_t per process group
            if self.R > 1:
                self.row_comm = TorchComm(group=self.row_group, device=self.ops.device)
            if self.C > 1:
                self.col_comm = TorchComm(group=self.col_group, device=self.ops.device)

    def query_slice(self, Q):
        """rows [lo, hi) of a Q-query batch this rank's query group answers, and the padded slice length."""
        per = -(-Q // self.C)
        lo = min(self.query_group * per, Q)
        return lo, min(lo + per, Q), per

    def batch_query(self, k, queries):
        """queries: the same [Q][D] tensor on every rank (a host tensor is sliced before it is copied to
        the device).  Returns merged (ids, dists, sizes) for the whole batch on every rank."""
        import torch
        if isinstance(self.ops, NativeOps):
            return self.ops.sharded_query(k, queries, self.row_comm, self.col_comm, self.row_offset)
        # the same orchestration spelled out over stand-in ops (CPU tests of the host logic)
        Q = queries.shape[0]
        lo, hi, per = self.query_slice(Q)
        mine = queries[lo:hi]
        if not mine.is_cuda and hasattr(self.ops, "device"):
            mine = mine.to(self.ops.device, non_blocking=True)
        if mine.shape[0] < per:      # equal slice lengths for the collectives
            pad = torch.zeros((per - mine.shape[0], queries.shape[1]), dtype=mine.dtype, device=mine.device)
            mine = torch.cat((mine, pad))
        ids, ds, sz = self.ops.local_query(k, mine, self.row_offset)
        if self.R > 1:
            # rank-major concatenation along dim 0 == [R][per][k]
            ids_all = torch.empty((self.R * per, k), dtype=ids.dtype, device=ids.device)
            ds_all = torch.empty((self.R * per, k), dtype=ds.dtype, device=ds.device)
            self.dist.all_gather_into_tensor(ids_all, ids.contiguous(), group=self.row_group)
            self.dist.all_gather_into_tensor(ds_all, ds.contiguous(), group=self.row_group)
            ids, ds, sz = self.ops.merge(ids_all.view(self.R, per, k), ds_all.view(self.R, per, k), k)
        if self.C > 1:
            out = []
            for t in (ids, ds, sz):
                t = t.contiguous()
                full = torch.empty((self.C * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                self.dist.all_gather_into_tensor(full, t, group=self.col_group)
                out.append(full[:Q])
            ids, ds, sz = out
        return ids, ds, sz
