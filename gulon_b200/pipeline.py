"""PQ candidates + exact fp32 re-rank as one pipeline (BASELINE configs[4]).

The reference has the pieces -- PQIndex#batchQuery (G/Index.scala:414-440) for candidates and
Index.exactNearestNeighbours / MathUtils.distanceSq (G/Index.scala:209-229, G/MathUtils.scala:85-95)
for exact distances, combined by its recall harness (G/Tests.scala:24-37) -- but no re-ranked
query; this is the production form of that combination, behind ONE C entry point
(gulon_pq_rerank_query[_dev]).  With more than one rank the code planes AND the raw vectors are
sharded by rows: the global PQ top-R is found as in the sharded query (all-gather + merge), every
rank scores the candidates whose rows it owns, and a second all-gather + merge yields the k best.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .index import PQIndex
from .vectors import DevicePoints


class RerankPipeline:
    """pq: ProductQuantizer; codes: CUDA uint8 [M][stride] planes of this rank's rows; X: CUDA float32
    [n_local][D] raw vectors of the same rows; row_offset: global id of local row 0."""

    def __init__(self, pq, codes, X, row_offset, n_local, world=1, group=None):
        import torch
        self.pq = pq
        self.index = PQIndex.from_device_codes(pq, codes, n_local)
        self.points = DevicePoints.from_torch(X)
        self.row_offset = int(row_offset)
        self.device = X.device
        self.comm = None
        if world > 1:
            from .sharded import TorchComm
            self.comm = TorchComm(group=group, device=self.device)
        self._torch = torch

    def query(self, k, R, queries, normalize=False):
        """queries: CUDA float32 [Q][D] (the same batch on every rank) -> (ids, dists, sizes) CUDA tensors;
        or a host array -> numpy arrays."""
        torch = self._torch
        rc = C.byref(self.comm.struct) if self.comm is not None else None
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            q = queries
            nq = q.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int32, device=q.device)
            ds = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            sz = torch.empty((nq,), dtype=torch.int32, device=q.device)
            ld = q.stride(0) if nq > 1 else max(q.shape[1], 1)
            N.check(N.lib().gulon_pq_rerank_query_dev(
                self.index.handle, self.points.handle, rc, q.data_ptr(), nq, ld, k, R, int(bool(normalize)),
                self.row_offset, ids.data_ptr(), ds.data_ptr(), sz.data_ptr(),
                torch.cuda.current_stream(q.device).cuda_stream))
            return ids, ds, sz
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids, ds, sz = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32), np.empty(nq, np.int32)
        N.check(N.lib().gulon_pq_rerank_query(
            self.index.handle, self.points.handle, rc, q.ctypes.data, nq, q.shape[1], k, R,
            int(bool(normalize)), self.row_offset, ids.ctypes.data, ds.ctypes.data, sz.ctypes.data))
        return ids, ds, sz

    def rerank_only(self, k, R, queries):
        """The exact-distance step alone on a fixed candidate list (for the gather-bandwidth roofline):
        candidates are this rank's rows row_offset + (q * 7919 + i * 104729) mod n_local."""
        torch = self._torch
        nq = queries.shape[0]
        n = self.index.length
        if getattr(self, "_cand", None) is None or self._cand.shape != (nq, R):
            qi = torch.arange(nq, device=self.device, dtype=torch.int64).unsqueeze(1)
            ri = torch.arange(R, device=self.device, dtype=torch.int64).unsqueeze(0)
            self._cand = ((qi * 7919 + ri * 104729) % max(n, 1) + self.row_offset).to(torch.int32).contiguous()
            self._ro = (torch.empty((nq, k), dtype=torch.int32, device=self.device),
                        torch.empty((nq, k), dtype=torch.float32, device=self.device))
        ld = queries.stride(0) if nq > 1 else max(queries.shape[1], 1)
        N.check(N.lib().gulon_rerank_dev(self.points.handle, queries.data_ptr(), nq, ld, self._cand.data_ptr(),
                                         R, k, self.row_offset, self._ro[0].data_ptr(), self._ro[1].data_ptr(),
                                         None, torch.cuda.current_stream(self.device).cuda_stream))
        return self._ro
