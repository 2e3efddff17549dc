"""KMeans: host-side mirror of G/KMeans.scala over the C ABI.

Same names and argument meaning as the reference (`KMeans.init`, `KMeans.fromAssignment`,
`KMeans.computeClusters`, `#assign`, `#parAssign`, `#iterate`, `Config`, `ProgressReport`); all
arithmetic runs in the CUDA library.  Argmin ties resolve to the lowest centroid index (the
reference draws `Random(0).nextBoolean()` on exact ties, G/KMeans.scala:47,90).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _native as N
from .vectors import Vectors


@dataclass
class ProgressReport:
    """KMeans.ProgressReport, G/KMeans.scala:119-127 (stepSize as mean / stdDev)."""
    num_iterations: int
    max_iterations: int
    step_mean: float
    step_stddev: float
    converged: bool
    quantizer: int = 0


@dataclass
class Config:
    """KMeans.Config, G/KMeans.scala:129-132."""
    num_clusters: int
    max_iterations: int
    seed: int = 0
    report: Optional[Callable[[ProgressReport], None]] = None
    update_mode: int = N.UPDATE_RUNNING_MEAN


def _progress_cb(report):
    if report is None:
        return N.PROGRESS_FN()

    def cb(_user, r):
        r = r.contents
        report(ProgressReport(r.num_iterations, r.max_iterations, r.step_mean, r.step_stddev,
                              bool(r.converged), r.quantizer))
    return N.PROGRESS_FN(cb)


class KMeans:
    """G/KMeans.scala: `dimension`, `centroids` [K][dimension] float32."""

    def __init__(self, dimension, centroids):
        self.dimension = int(dimension)
        self.centroids = np.ascontiguousarray(centroids, np.float32).reshape(-1, self.dimension)

    @property
    def k(self):
        return self.centroids.shape[0]

    # -- KMeans#assign (G/KMeans.scala:18-22,70-98) / #parAssign (:57-68) ---------------------
    def assign(self, vecs: Vectors, _batch=0):
        if vecs.dimension != self.dimension:
            raise ValueError("dimension mismatch: %d vs %d" % (vecs.dimension, self.dimension))
        out = np.zeros(vecs.size, np.int32)
        N.check(N.lib().gulon_kmeans_assign(vecs.matrix.device().handle, vecs.from_, self.dimension,
                                            self.centroids.ctypes.data, self.k, _batch,
                                            N.TIE_LOWEST, out.ctypes.data))
        return out

    def par_assign(self, vecs: Vectors):
        return self.assign(vecs, _batch=25000)

    # -- KMeans#iterate (G/KMeans.scala:100-106) -------------------------------------------------
    def iterate(self, vecs: Vectors, iters: int, update_mode=N.UPDATE_RUNNING_MEAN):
        cur = self
        for _ in range(iters):
            a = cur.assign(vecs)
            cur = KMeans.from_assignment(self.k, self.dimension, vecs, a, update_mode)
        return cur

    # -- KMeans.init (G/KMeans.scala:188-196) -----------------------------------------------------
    @staticmethod
    def init(k, vecs: Vectors, seed=0, return_rows=False):
        cm = np.zeros((k, vecs.dimension), np.float32)
        rows = np.zeros(k, np.int32)
        N.check(N.lib().gulon_kmeans_init(vecs.matrix.device().handle, vecs.from_, vecs.dimension,
                                          k, seed, cm.ctypes.data, rows.ctypes.data))
        km = KMeans(vecs.dimension, cm)
        return (km, rows) if return_rows else km

    # -- KMeans.fromAssignment (G/KMeans.scala:198-226) -----------------------------------------
    @staticmethod
    def from_assignment(k, dimension, vecs: Vectors, assignments,
                        update_mode=N.UPDATE_RUNNING_MEAN, return_counts=False):
        a = np.ascontiguousarray(assignments, np.int32)
        if a.shape[0] != vecs.size:
            raise ValueError("one assignment per row expected")
        cm = np.zeros((k, dimension), np.float32)
        cnt = np.zeros(k, np.int32)
        N.check(N.lib().gulon_kmeans_update(vecs.matrix.device().handle, vecs.from_, dimension,
                                            a.ctypes.data, k, update_mode, cm.ctypes.data,
                                            cnt.ctypes.data))
        km = KMeans(dimension, cm)
        return (km, cnt) if return_counts else km

    # -- KMeans.computeClusters (G/KMeans.scala:134-157) ----------------------------------------
    @staticmethod
    def compute_clusters(vecs: Vectors, config: Config, comm=None, n_total=0, row_offset=0,
                         return_info=False):
        cm = np.zeros((config.num_clusters, vecs.dimension), np.float32)
        nu, conv = N.i32(0), N.i32(0)
        cb = _progress_cb(config.report)
        N.check(N.lib().gulon_kmeans_train(
            vecs.matrix.device().handle, vecs.from_, vecs.dimension, config.num_clusters,
            config.max_iterations, config.seed, N.TIE_LOWEST, config.update_mode,
            C.byref(comm) if comm is not None else None, n_total, row_offset, cb, None,
            cm.ctypes.data, C.byref(nu), C.byref(conv)))
        km = KMeans(vecs.dimension, cm)
        if return_info:
            return km, dict(updates=nu.value, converged=bool(conv.value))
        return km
