"""Index.PQIndex / prepareQuery / exactNearestNeighbours: host-side mirror of G/Index.scala
(:209-229, :352-441) and the result ordering of Result.fromHeap (:83-94) over the C ABI.

Top-k order is (distance ascending, id ascending) -- the stable order T/TopKHeapSpec.scala:16-31
asserts; the reference heap's order inside an equal-distance group is structure-dependent.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .quantizer import EncodedMatrix, ProductQuantizer
from .vectors import DevicePoints, Matrix


@dataclass
class TopK:
    """The drained TopKHeaps of a batch: keys [Q][k] (-1 = empty slot), values [Q][k] squared
    distances ascending (+inf = empty), size [Q]."""
    keys: np.ndarray
    values: np.ndarray
    size: np.ndarray

    def __len__(self):
        return self.keys.shape[0]

    def __getitem__(self, q):
        n = int(self.size[q])
        return self.keys[q, :n], self.values[q, :n]


def prepare_query(product_quantizer: ProductQuantizer, queries):
    """Index.prepareQuery, G/Index.scala:352-383 -> float32 [Q][M][K]."""
    q = np.ascontiguousarray(queries, np.float32)
    if q.ndim != 2 or q.shape[1] != product_quantizer.dimension:
        raise ValueError("expected [Q][%d] queries" % product_quantizer.dimension)
    M = len(product_quantizer.quantizers)
    lut = np.zeros((q.shape[0], M, product_quantizer.num_clusters), np.float32)
    N.check(N.lib().gulon_prepare_query(product_quantizer.handle, q.ctypes.data, q.shape[0],
                                        q.shape[1], lut.ctypes.data))
    return lut


def exact_nearest_neighbours(vectors, queries, k, from_=0, until=None):
    """Index.exactNearestNeighbours, G/Index.scala:209-229, batched over queries."""
    dev = vectors.device() if isinstance(vectors, Matrix) else vectors
    if not isinstance(dev, DevicePoints):
        dev = Matrix(vectors).device()
    n = N.i64(0)
    N.check(N.lib().gulon_points_info(dev.handle, C.byref(n), None, None, None))
    until = n.value if until is None else until
    q = np.ascontiguousarray(queries, np.float32)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    nq = q.shape[0]
    ids = np.full((nq, k), -1, np.int32)
    ds = np.full((nq, k), np.inf, np.float32)
    sz = np.zeros(nq, np.int32)
    N.check(N.lib().gulon_exact_topk(dev.handle, q.ctypes.data, nq, q.shape[1], k, from_, until,
                                     ids.ctypes.data, ds.ctypes.data, sz.ctypes.data))
    return TopK(ids, ds, sz)


class PQIndex:
    """Index.PQIndex(productQuantizer, data), G/Index.scala:385-441: device-resident code planes."""

    def __init__(self, product_quantizer: ProductQuantizer, data, _dev_codes=None, length=None):
        self.product_quantizer = product_quantizer
        h = N.vp()
        if _dev_codes is not None:
            t = _dev_codes  # CUDA uint8 (K <= 256) or uint16 [M][stride] torch tensor, borrowed
            self._keepalive = t
            self.length = int(length)
            fn = N.lib().gulon_index_create16_dev if t.element_size() == 2 else N.lib().gulon_index_create_dev
            if (t.element_size() == 2) != (product_quantizer.num_clusters > 256):
                raise ValueError("16-bit codes are for more than 256 clusters, 8-bit codes for up to 256")
            N.check(fn(product_quantizer.handle, t.data_ptr(), self.length, t.stride(0), C.byref(h)))
        else:
            if not isinstance(data, EncodedMatrix):
                raise ValueError("expected an EncodedMatrix")
            if data.codes.shape[0] != len(product_quantizer.quantizers):
                raise ValueError("one code plane per quantizer expected")
            self._keepalive = None
            self.length = data.length
            wide = product_quantizer.num_clusters > 256
            if (data.codes.dtype == np.uint16) != wide:
                raise ValueError("16-bit codes are for more than 256 clusters, 8-bit codes for up to 256")
            fn = N.lib().gulon_index_create16 if wide else N.lib().gulon_index_create
            N.check(fn(product_quantizer.handle, data.codes.ctypes.data, data.length, data.length,
                       C.byref(h)))
        self.data = data
        self._handle = h

    @classmethod
    def from_device_codes(cls, product_quantizer, codes, length):
        return cls(product_quantizer, None, _dev_codes=codes, length=length)

    @property
    def dimension(self):
        return self.product_quantizer.dimension

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        try:
            if self._handle:
                N.lib().gulon_index_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def decode(self, row):
        return self.product_quantizer.decode(self.data(row))

    def batch_query(self, k, vectors, from_=0, until=None, normalize=False, id_offset=0):
        """PQIndex#batchQuery(k, vectors, from, until), G/Index.scala:414-440; host buffers."""
        q = vectors.data if isinstance(vectors, Matrix) else np.ascontiguousarray(vectors, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dimension:
            raise ValueError("expected [Q][%d] queries" % self.dimension)
        until = self.length if until is None else until
        nq = q.shape[0]
        ids = np.full((nq, k), -1, np.int32)
        ds = np.full((nq, k), np.inf, np.float32)
        sz = np.zeros(nq, np.int32)
        N.check(N.lib().gulon_pq_query(self._handle, q.ctypes.data, nq, q.shape[1], k, from_, until,
                                       int(bool(normalize)), id_offset, ids.ctypes.data,
                                       ds.ctypes.data, sz.ctypes.data))
        return TopK(ids, ds, sz)

    def query(self, k, query, from_=0, until=None, normalize=False):
        """PQIndex#query(k, query, from, until), G/Index.scala:411-412."""
        r = self.batch_query(k, np.asarray(query, np.float32).reshape(1, -1), from_, until, normalize)
        return r[0]

    def batch_query_dev(self, k, q, from_=0, until=None, normalize=False, id_offset=0, out=None,
                        stream=None):
        """Device-resident form: q CUDA float32 [Q][D] torch tensor -> (ids, dists, sizes) tensors."""
        import torch
        until = self.length if until is None else until
        nq = q.shape[0]
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int32, device=q.device),
                   torch.empty((nq, k), dtype=torch.float32, device=q.device),
                   torch.empty((nq,), dtype=torch.int32, device=q.device))
        st = torch.cuda.current_stream(q.device).cuda_stream if stream is None else stream
        ld = q.stride(0) if nq > 1 else max(q.shape[1], 1)
        N.check(N.lib().gulon_pq_query_dev(self._handle, q.data_ptr(), nq, ld, k, from_, until,
                                           int(bool(normalize)), id_offset, out[0].data_ptr(),
                                           out[1].data_ptr(), out[2].data_ptr(), st))
        return out


def debug_tscan(index, q_dev, taus, from_, until):
    """Diagnostics of the tensor scan (gulon_debug_tscan): operands and raw accumulators of one filter
    launch for <= 256 device queries with the given thresholds -> (xb [rows][KP] uint16, qb [256][KP] uint16,
    acc [rows][256] float32)."""
    nq = q_dev.shape[0]
    rows = until - from_
    taus = np.ascontiguousarray(taus, np.float32)
    kp = N.i32(0)
    acc = np.zeros((rows, 256), np.float32)
    # KP is only known after the call: size the operand buffers for the largest contraction (1024)
    xb = np.zeros((rows, 1024), np.uint16)
    qb = np.zeros((256, 1024), np.uint16)
    N.check(N.lib().gulon_debug_tscan(index.handle, q_dev.data_ptr(), nq, q_dev.stride(0), taus.ctypes.data,
                                      from_, until, xb.ctypes.data, qb.ctypes.data, acc.ctypes.data,
                                      C.byref(kp)))
    KP = int(kp.value)
    xb = xb.reshape(-1)[:rows * KP].reshape(rows, KP)
    qb = qb.reshape(-1)[:256 * KP].reshape(256, KP)
    return xb, qb, acc


def rerank(vectors, queries, cand_ids, k):
    """Exact fp32 re-rank of PQ candidates (distanceSq of G/MathUtils.scala:85-95 over the raw
    vectors, as the recall harness G/Tests.scala:24-37 scores returned keys)."""
    dev = vectors.device() if isinstance(vectors, Matrix) else vectors
    q = np.ascontiguousarray(queries, np.float32)
    c = np.ascontiguousarray(cand_ids, np.int32)
    nq, R = c.shape
    ids = np.full((nq, k), -1, np.int32)
    ds = np.full((nq, k), np.inf, np.float32)
    sz = np.zeros(nq, np.int32)
    N.check(N.lib().gulon_rerank(dev.handle, q.ctypes.data, nq, q.shape[1], c.ctypes.data, R, k,
                                 ids.ctypes.data, ds.ctypes.data, sz.ctypes.data))
    return TopK(ids, ds, sz)
