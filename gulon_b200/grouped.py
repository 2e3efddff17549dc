"""WordVectors.Grouped / Index.GroupedIndex: the coarse-partitioned (IVF) index over the C ABI.

Reference: WordVectors#grouped and Grouped#residuals (G/WordVectors.scala:24-58,118-138),
Index.grouped (G/Index.scala:133-145), GroupedIndex#query / searchSpace / lookup
(G/Index.scala:232-299).  Host logic only (the sort by (cluster, key), the partition table, the
search-space strategies, batching of (query, partition) pairs); every arithmetic step runs in the
library's kernels: KMeans#parAssign (tensor-core assignment), MathUtils.subtract per row
(`gulon_subtract_rows_dev`), ProductQuantizer#encode, the coarse probe (`gulon_exact_topk`),
the ranged ADC scans (`gulon_pq_query_dev` on [from, until)) and the (distance, id) merge of the
per-partition heaps (`gulon_topk_merge_dev`).  torch is plumbing (device buffers, gathers).

Literal quirk kept: WordVectors#grouped seeds its run detection with `assignments(0)` -- the
assignment of ROW 0, not of the first sorted row (G/WordVectors.scala:39) -- so unless row 0 belongs to
the lowest non-empty cluster the partition table starts with an EMPTY group whose centroid duplicates
a later group's.  Indexes built here have the same partition table as the reference's.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .index import PQIndex, TopK, exact_nearest_neighbours
from .kmeans import KMeans
from .quantizer import ProductQuantizer
from .vectors import DevicePoints, Matrix, Vectors, normalize


@dataclass(frozen=True)
class LimitGroups:
    """GroupedIndex.Strategy.LimitGroups(count), G/Index.scala:305."""
    count: int


@dataclass(frozen=True)
class LimitVectors:
    """GroupedIndex.Strategy.LimitVectors(count), G/Index.scala:306."""
    count: int


def _subtract_rows(x, src, cent, group, out):
    """out[i] = x[src[i]] - cent[group[i]] on the device (torch tensors; src / group may be None)."""
    import torch
    n, D = out.shape
    st = torch.cuda.current_stream(out.device).cuda_stream
    N.check(N.lib().gulon_subtract_rows_dev(
        x.data_ptr(), x.stride(0) if x.shape[0] > 1 else max(D, 1),
        src.data_ptr() if src is not None else None,
        cent.data_ptr() if cent is not None else None,
        (cent.stride(0) if cent.shape[0] > 1 else max(D, 1)) if cent is not None else 0,
        group.data_ptr() if group is not None else None, n, D, out.data_ptr(),
        out.stride(0) if n > 1 else max(D, 1), st))
    return out


class GroupedVectors:
    """WordVectors.Grouped(keys, toMatrix, centroids, offsets), G/WordVectors.scala:99-139.

    order[i] = original row of grouped position i; matrix = the rows in grouped order (device
    tensor); centroids [P][D]; offsets [P-1] (start of groups 1..P-1)."""

    def __init__(self, order, matrix_dev, centroids, offsets, keys=None):
        self.order = order
        self.matrix_dev = matrix_dev
        self.centroids = np.ascontiguousarray(centroids, np.float32)
        self.offsets = np.ascontiguousarray(offsets, np.int32)
        self.keys = keys
        if self.centroids.shape[0] != self.offsets.shape[0] + 1 and self.size > 0:
            raise AssertionError("%d != %d + 1" % (self.centroids.shape[0], self.offsets.shape[0]))

    @property
    def size(self):
        return int(self.order.shape[0])

    @property
    def dimension(self):
        return int(self.centroids.shape[1])

    def bounds(self, i):
        """GroupedIndex#getBounds, G/Index.scala:260-264."""
        start = 0 if i == 0 else int(self.offsets[i - 1])
        end = self.size if i == len(self.offsets) else int(self.offsets[i])
        return start, end

    def cluster_of(self, i):
        """Grouped#clusterOf, G/WordVectors.scala:111-114 (binary search over the offsets)."""
        return int(np.searchsorted(self.offsets, i, side="right"))

    def group_of_positions(self):
        return np.searchsorted(self.offsets, np.arange(self.size), side="right").astype(np.int32)

    @staticmethod
    def group(matrix, clustering: KMeans, keys=None, device=None):
        """WordVectors#grouped(clustering), G/WordVectors.scala:24-58."""
        import torch
        if clustering.k <= 0:
            raise ValueError("must have at least 1 cluster")
        m = matrix if isinstance(matrix, Matrix) else Matrix(matrix)
        a = clustering.par_assign(Vectors(m))
        n = m.rows
        # Array.range(0, size).sortBy(word(_)).sortBy(assignments(_)): two stable sorts
        if keys is None:
            by_key = np.arange(n)
        else:
            if len(keys) != n:
                raise ValueError("one key per row expected")
            # sortBy(word(_)): java.lang.String#compareTo order = UTF-16 code units
            k16 = [w.encode("utf-16-be", "surrogatepass") if isinstance(w, str) else w for w in keys]
            by_key = np.asarray(sorted(range(n), key=lambda i: k16[i]), dtype=np.int64)
        order = by_key[np.argsort(a[by_key], kind="stable")].astype(np.int64)
        if n > 0:
            a_sorted = a[order]
            change = np.flatnonzero(a_sorted[1:] != a_sorted[:-1]) + 1
            prev0 = int(a[0])                       # literal: assignments(0), G/WordVectors.scala:39
            if int(a_sorted[0]) != prev0:
                change = np.concatenate(([0], change))
            cents = np.concatenate((clustering.centroids[[prev0]], clustering.centroids[a_sorted[change]]))
            offsets = change.astype(np.int32)
        else:
            cents = np.zeros((0, m.cols), np.float32)
            offsets = np.zeros(0, np.int32)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        x = torch.from_numpy(m.data).to(dev)
        g = torch.empty((n, m.cols), dtype=torch.float32, device=dev)
        if n > 0:
            _subtract_rows(x, torch.from_numpy(order).to(dev), None, None, g)   # gather only
        gk = [keys[j] for j in order] if keys is not None else None
        return GroupedVectors(order, g, cents, offsets, gk)

    def residuals_dev(self):
        """Grouped#residuals, G/WordVectors.scala:118-138: row minus the centroid of its group."""
        import torch
        dev = self.matrix_dev.device
        out = torch.empty_like(self.matrix_dev)
        if self.size > 0:
            _subtract_rows(self.matrix_dev, None, torch.from_numpy(self.centroids).to(dev),
                           torch.from_numpy(self.group_of_positions()).to(dev), out)
        return out


class GroupedIndex:
    """Index.GroupedIndex(keyIndex, vectorIndex, metric, clustering, strategy), G/Index.scala:229-299.

    Results are grouped positions (`TopK.keys`); `original_rows(keys)` maps them to the rows of the
    matrix the index was built from, `words(keys)` to the keys if any were given."""

    def __init__(self, grouped: GroupedVectors, vector_index: PQIndex, normalized: bool, strategy):
        self.grouped = grouped
        self.vector_index = vector_index
        self.normalized = bool(normalized)
        self.strategy = strategy
        self._cent_points = None
        self.work_list = True     # one launch per batch (gulon_grouped_query_dev); False: one ranged scan per partition

    # -- Index.grouped, G/Index.scala:133-145 --------------------------------------------------------
    @staticmethod
    def build(grouped: GroupedVectors, residuals_quantizer: ProductQuantizer, normalized=False,
              strategy=LimitGroups(8)):
        codes = residuals_quantizer.encode_dev(grouped.residuals_dev())
        ix = PQIndex.from_device_codes(residuals_quantizer, codes, grouped.size)
        return GroupedIndex(grouped, ix, normalized, strategy)

    @property
    def dimension(self):
        return self.vector_index.dimension

    @property
    def size(self):
        return self.grouped.size

    def original_rows(self, keys):
        k = np.asarray(keys)
        out = np.full(k.shape, -1, np.int64)
        ok = k >= 0
        out[ok] = self.grouped.order[k[ok]]
        return out

    def words(self, keys):
        if self.grouped.keys is None:
            raise ValueError("the index was built without keys")
        return [[self.grouped.keys[i] for i in row if i >= 0] for row in np.atleast_2d(keys)]

    # -- GroupedIndex#lookup, G/Index.scala:247-254 ------------------------------------------------
    def lookup(self, position):
        part = self.grouped.cluster_of(position)
        base = self.grouped.centroids[part]
        if self.vector_index.data is not None:
            codes = self.vector_index.data(position)
        else:
            codes = self.vector_index._keepalive[:, position].cpu().numpy()
        residual = self.vector_index.product_quantizer.decode(codes)
        return (base + residual).astype(np.float32)      # MathUtils.add, fp32

    def query_by_position(self, k, position):
        """Index#queryByWord, G/Index.scala:44-45."""
        return self.query(k, self.lookup(position))

    # -- GroupedIndex#searchSpace, G/Index.scala:283-299 ----------------------------------------------
    def search_space(self, queries):
        """-> list of int arrays: the partitions each (already normalised) query probes, nearest first."""
        P = self.grouped.centroids.shape[0]
        if P == 0 or queries.shape[0] == 0:
            return [np.zeros(0, np.int32) for _ in range(queries.shape[0])]
        if self._cent_points is None:
            self._cent_points = DevicePoints.from_host(self.grouped.centroids)
        if isinstance(self.strategy, LimitGroups):
            m = min(max(int(self.strategy.count), 0), P)
            nn = exact_nearest_neighbours(self._cent_points, queries, m)
            return [nn.keys[q, :nn.size[q]] for q in range(queries.shape[0])]
        order = exact_nearest_neighbours(self._cent_points, queries, P)
        sizes = np.diff(np.concatenate(([0], self.grouped.offsets, [self.size]))).astype(np.int64)
        out = []
        for q in range(queries.shape[0]):
            o = order.keys[q, :order.size[q]]
            cum = np.cumsum(sizes[o])
            # probe until the cumulative size reaches the limit (the group that crosses it included)
            stop = int(np.searchsorted(cum, int(self.strategy.count), side="left")) + 1
            out.append(o[:min(stop, len(o))] if int(self.strategy.count) > 0 else o[:0])
        return out

    # -- GroupedIndex#query / batchQuery, G/Index.scala:255-281 ---------------------------------------
    def query(self, k, query):
        r = self.batch_query(k, np.asarray(query, np.float32).reshape(1, -1))
        return r[0]

    def batch_query(self, k, vectors, batch=16384):
        import torch
        q = vectors.data if isinstance(vectors, Matrix) else np.ascontiguousarray(vectors, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dimension:
            raise ValueError("expected [Q][%d] queries" % self.dimension)
        if self.normalized:
            q = normalize(q)
        Q = q.shape[0]
        ids = np.full((Q, k), -1, np.int32)
        ds = np.full((Q, k), np.inf, np.float32)
        sz = np.zeros(Q, np.int32)
        if self.grouped.matrix_dev is not None:
            dev = self.grouped.matrix_dev.device
        elif self.vector_index._keepalive is not None:
            dev = self.vector_index._keepalive.device      # an index loaded from disk has no raw rows
        else:
            dev = torch.device("cuda", torch.cuda.current_device())
        cent_dev = torch.from_numpy(self.grouped.centroids).to(dev)
        for q0 in range(0, Q, batch):
            qb = q[q0:q0 + batch]
            r = self._query_batch(k, qb, cent_dev)
            ids[q0:q0 + len(qb)], ds[q0:q0 + len(qb)], sz[q0:q0 + len(qb)] = r
        return TopK(ids, ds, sz)

    def _query_batch(self, k, q, cent_dev):
        import torch
        dev = cent_dev.device
        nq = q.shape[0]
        probes = self.search_space(q)
        S = max((len(p) for p in probes), default=0)
        out_ids = np.full((nq, k), -1, np.int32)
        out_ds = np.full((nq, k), np.inf, np.float32)
        out_sz = np.zeros(nq, np.int32)
        if S == 0 or k == 0 or nq == 0:
            return out_ids, out_ds, out_sz
        if self.work_list and 1 <= k <= 1024 and self.vector_index.product_quantizer.num_clusters <= 256:
            return self._query_batch_work_list(k, q, cent_dev, probes, S)
        return self._query_batch_ranged(k, q, cent_dev, probes, S)

    def _query_batch_work_list(self, k, q, cent_dev, probes, S):
        """ONE launch for every probed (query, partition) pair: gulon_grouped_query_dev."""
        import torch
        dev = cent_dev.device
        nq = q.shape[0]
        P = self.grouped.centroids.shape[0]
        pair_q = np.concatenate([np.full(len(p), i, np.int32) for i, p in enumerate(probes)])
        pair_part = np.concatenate(probes).astype(np.int32)
        pair_slot = np.concatenate([np.arange(len(p), dtype=np.int32) for p in probes])
        bounds = np.concatenate(([0], self.grouped.offsets, [self.size])).astype(np.int32)
        # pairs of one partition next to each other: its code rows stay in L2 while they are scanned
        o = np.argsort(pair_part, kind="stable")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_pq, d_pp, d_ps, d_b = t(pair_q[o]), t(pair_part[o]), t(pair_slot[o]), t(bounds)
        qd = t(q)
        ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
        sz = torch.empty((nq,), dtype=torch.int32, device=dev)
        N.check(N.lib().gulon_grouped_query_dev(
            self.vector_index.handle, qd.data_ptr(), nq, q.shape[1], cent_dev.data_ptr(), P, d_b.data_ptr(),
            d_pq.data_ptr(), d_pp.data_ptr(), d_ps.data_ptr(), len(pair_q), S, k, ids.data_ptr(), ds.data_ptr(),
            sz.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize(dev)
        return ids.cpu().numpy(), ds.cpu().numpy(), sz.cpu().numpy()

    def _query_batch_ranged(self, k, q, cent_dev, probes, S):
        """The same through one ranged PQIndex#batchQuery per probed partition (any k, K > 256; the
        cross-check of the work-list kernel)."""
        import torch
        dev = cent_dev.device
        nq = q.shape[0]
        out_ids = np.full((nq, k), -1, np.int32)
        out_ds = np.full((nq, k), np.inf, np.float32)
        out_sz = np.zeros(nq, np.int32)
        # (query, partition, probe rank) triples, grouped by partition
        qi = np.concatenate([np.full(len(p), i, np.int64) for i, p in enumerate(probes)])
        part = np.concatenate(probes).astype(np.int32)
        rank = np.concatenate([np.arange(len(p), dtype=np.int64) for p in probes])
        o = np.argsort(part, kind="stable")
        qi, part, rank = qi[o], part[o], rank[o]
        n_pairs = len(qi)
        qd = torch.from_numpy(q).to(dev)
        res = torch.empty((n_pairs, self.dimension), dtype=torch.float32, device=dev)
        _subtract_rows(qd, torch.from_numpy(qi).to(dev), cent_dev, torch.from_numpy(part).to(dev), res)
        # one slot per probe rank: the per-partition heaps land at [rank][query]
        slot_ids = torch.full((S, nq, k), -1, dtype=torch.int32, device=dev)
        slot_ds = torch.full((S, nq, k), float("inf"), dtype=torch.float32, device=dev)
        starts = np.flatnonzero(np.concatenate(([True], part[1:] != part[:-1])))
        ends = np.concatenate((starts[1:], [n_pairs]))
        dest = torch.from_numpy(rank * nq + qi).to(dev)
        flat_ids = slot_ids.view(S * nq, k)
        flat_ds = slot_ds.view(S * nq, k)
        for s, e in zip(starts, ends):
            frm, until = self.grouped.bounds(int(part[s]))
            if until <= frm:
                continue                                   # empty group: an empty heap
            r_ids, r_ds, _ = self.vector_index.batch_query_dev(k, res[s:e], frm, until)
            flat_ids.index_copy_(0, dest[s:e], r_ids)
            flat_ds.index_copy_(0, dest[s:e], r_ds)
        m_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        m_ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
        m_sz = torch.empty((nq,), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().gulon_topk_merge_dev(slot_ids.data_ptr(), slot_ds.data_ptr(), S, nq, k,
                                             m_ids.data_ptr(), m_ds.data_ptr(), m_sz.data_ptr(), st))
        torch.cuda.synchronize(dev)
        return m_ids.cpu().numpy(), m_ds.cpu().numpy(), m_sz.cpu().numpy()
