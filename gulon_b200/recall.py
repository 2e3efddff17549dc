"""The reference's recall harness, G/Tests.scala:11-121, with the brute-force parts on the device:
ground truth = Index.exactNearestNeighbours over the raw vectors (`gulon_exact_topk`), scoring = the
exact distance of every returned key (`gulon_rerank`), recall@k = share of the first k returned keys
whose exact distance is within the k-th true distance (eps = 0) or within (1 + eps) of it.

SummaryStats (G/MathUtils.scala:5-60) is mirrored in fp32 so that means / variances print like the
reference's.  Sampling uses java.util.Random(seed).nextInt(size) like Tests.sample (G/Tests.scala:76-87).
"""
from dataclasses import dataclass

import numpy as np

from .index import exact_nearest_neighbours, rerank
from .vectors import DevicePoints, Matrix

DEFAULT_KS = (1, 2, 3, 5, 10, 25, 50, 100, 500, 1000)      # Tests.defaultKs, G/Tests.scala:54

f32 = np.float32


class JavaRandom:
    """java.util.Random: the 48-bit LCG behind scala.util.Random (nextInt(bound) per its Javadoc)."""

    def __init__(self, seed):
        self.s = (int(seed) ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def _next(self, bits):
        self.s = (self.s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.s >> (48 - bits)
        return v - (1 << 32) if v >= 1 << 31 else v               # (int) cast: low 32 bits, signed

    def next_int(self, bound):
        if bound <= 0:
            raise ValueError("bound must be positive")
        if bound & (bound - 1) == 0:
            return (bound * self._next(31)) >> 31
        while True:
            bits = self._next(31)
            val = bits % bound
            if bits - val + (bound - 1) < (1 << 31):
                return val


@dataclass(frozen=True)
class SummaryStats:
    """SummaryStats(count, mean, s), G/MathUtils.scala:5-24 (fp32)."""
    count: int = 0
    mean: float = 0.0
    s: float = 0.0

    @property
    def variance(self):
        return float(f32(self.s) / f32(self.count)) if self.count else float("nan")

    @property
    def std_dev(self):
        return float(f32(np.sqrt(np.float64(self.variance))))

    def __add__(self, that):
        """`++`, G/MathUtils.scala:9-20."""
        if that.count == 0:
            return self
        if self.count == 0:
            return that
        n = self.count + that.count
        d = f32(self.mean) - f32(that.mean)
        mean = f32(self.mean) + f32(f32(that.count) / f32(n)) * f32(f32(that.mean) - f32(self.mean))
        s = f32(f32(self.s) + f32(that.s)) + f32(f32(f32(f32(d * d) * f32(self.count)) * f32(that.count)) / f32(n))
        return SummaryStats(n, float(mean), float(s))

    @staticmethod
    def of(values):
        """SummaryStats(xs) through the builder, G/MathUtils.scala:43-57 (Welford, fp32)."""
        m, s, n = f32(0), f32(0), 0
        for x in values:
            x = f32(x)
            n += 1
            m0 = m
            m = f32(m0 + f32(f32(x - m0) / f32(n)))
            s = f32(s + f32(f32(x - m0) * f32(x - m)))
        return SummaryStats(n, float(m), float(s))


@dataclass
class Query:
    """Tests.Query(query, results) with results = [(k, maxDistanceSq)], G/Tests.scala:46-47."""
    query: np.ndarray
    results: list


class Tests:
    """Tests(wordVectors, queries), G/Tests.scala:11-41.  `points` are the raw vectors the index was
    built from (DevicePoints, Matrix or a float32 array)."""

    def __init__(self, points, queries):
        self.points = points
        self.queries = list(queries)

    # -- construction -----------------------------------------------------------------------------
    @staticmethod
    def _points(points):
        if isinstance(points, DevicePoints):
            return points
        m = points if isinstance(points, Matrix) else Matrix(np.ascontiguousarray(points, np.float32))
        return m.device()

    @classmethod
    def for_queries(cls, points, queries, ks=DEFAULT_KS):
        """Tests.forQueries / getDistances, G/Tests.scala:89-110: the exact k-th distances."""
        pts = cls._points(points)
        q = np.ascontiguousarray(queries, np.float32)
        kmax = max(ks)
        nn = exact_nearest_neighbours(pts, q, kmax)
        out = []
        for i in range(q.shape[0]):
            n = int(nn.size[i])
            out.append(Query(q[i], [(k, float(nn.values[i, k - 1])) for k in ks if k <= n]))
        return cls(pts, out)

    @classmethod
    def sample(cls, points, rows_host, sample_size=1000, ks=DEFAULT_KS, seed=0):
        """Tests.sample, G/Tests.scala:76-87: queries are database rows drawn with
        java.util.Random(seed).nextInt(size); `rows_host(i)` returns row i as a float32 array."""
        pts = cls._points(points)
        rng = JavaRandom(seed)
        n = pts.rows
        idx = [rng.next_int(n) for _ in range(sample_size)]
        q = np.stack([np.asarray(rows_host(i), np.float32) for i in idx]) if idx else \
            np.zeros((0, pts.cols), np.float32)
        t = cls.for_queries(pts, q, ks)
        t.sampled_rows = idx
        return t

    # -- recallOf, G/Tests.scala:18-41 ----------------------------------------------------------------
    def recall_of(self, index, eps=0.0, to_rows=None):
        """-> {k: SummaryStats} over the queries.  `index.batch_query(k, Q)` must return a TopK; `to_rows`
        maps its keys to rows of `points` (GroupedIndex: `index.original_rows`)."""
        if not self.queries:
            return {}
        if to_rows is None and hasattr(index, "original_rows"):
            to_rows = index.original_rows
        max_k = max((k for qq in self.queries for k, _ in qq.results), default=0)
        if max_k == 0:
            return {}
        Q = np.stack([qq.query for qq in self.queries]).astype(np.float32)
        got = index.batch_query(max_k, Q)
        keys = np.asarray(got.keys)
        rows = np.asarray(to_rows(keys) if to_rows is not None else keys).astype(np.int64)
        rows32 = np.where(keys >= 0, rows, -1).astype(np.int32)
        ex = rerank(self.points, Q, rows32, max_k)             # exact distances, (distance, id) order
        per_k = {}
        for i, qq in enumerate(self.queries):
            n = int(got.size[i])
            # exact distance of each returned key, in the order the index returned them
            order = np.argsort(ex.keys[i, :n], kind="stable")
            at = np.searchsorted(ex.keys[i, :n][order], rows32[i, :n])
            dist = ex.values[i, :n][order][at]
            for k, max_d in qq.results:
                if eps == 0.0:
                    cutoff = f32(max_d)
                else:
                    cutoff = f32(np.float64(np.sqrt(np.float64(f32(max_d))) * np.float64(f32(1.0) + f32(eps))) ** 2)
                tp = int(np.count_nonzero(dist[:k] <= cutoff))
                per_k.setdefault(k, []).append(f32(tp) / f32(k))
        # Monoid.combineAll over per-query single-value stats
        out = {}
        for k, vals in per_k.items():
            acc = SummaryStats()
            for v in vals:
                acc = acc + SummaryStats(1, float(v), 0.0)
            out[k] = acc
        return out
