"""ProductQuantizer / Coder / EncodedMatrix: host-side mirror of G/ProductQuantizer.scala,
G/Coder.scala (coder.py; K <= 256, i.e. widths 0/2/4/8, on the device) and G/EncodedMatrix.scala
over the C ABI.

Codes are plane-major: one uint8 array of length N per quantizer (EncodedMatrix.encodings,
G/EncodedMatrix.scala:11-23), held here as one uint8 [M][N] array.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from . import _native as N
from .kmeans import KMeans, ProgressReport as KMeansProgressReport, _progress_cb
from .coder import Coder0, Coder2, Coder4, Coder8, BytePlus, coder, factory_for, max_width
from .vectors import Matrix, subvector_windows


def coder_width(num_clusters):
    """ProductQuantizer.coderFactory, G/ProductQuantizer.scala:11-16: 32 - nlz(K - 1), rounded up to a
    supported coder (G/Coder.scala:35-45).  On the device ids are one byte up to K = 256 (coders 0, 2, 4,
    8: the fast scan kernels) and 16 bits above (the BytePlus coders 10, 12, 16: plain-table scan)."""
    w = max_width(num_clusters)
    f = factory_for(w)
    if f is None:
        raise ValueError("too many clusters: %d" % num_clusters)
    return f.width


class EncodedMatrix:
    """G/EncodedMatrix.scala:11-51: one packed `Code` per quantizer, each holding the ids of all N rows.

    `codes` is the unpacked view the kernels use -- [M][N], uint8 for widths <= 8, uint16 for the BytePlus
    widths -- whatever the coder's packed width; `encodings` / `unwrapped_encodings` are the reference's
    packed planes."""

    def __init__(self, coder, encodings):
        self.coder = coder
        dt = np.uint8 if coder.width <= 8 else np.uint16
        if isinstance(encodings, np.ndarray) and encodings.ndim == 2 and coder.width == 8:
            self.codes = np.ascontiguousarray(encodings, np.uint8)
        elif isinstance(encodings, np.ndarray) and encodings.ndim != 2:
            raise ValueError("expected codes [M][N]")
        else:
            planes = [coder.unpack(coder.wrap_code(e) if e is not None else None) for e in encodings]
            self.codes = (np.stack(planes).astype(dt) if planes else np.zeros((0, coder.length), dt))
        if self.codes.ndim != 2:
            raise ValueError("expected codes [M][N]")

    @classmethod
    def from_planes(cls, coder, planes):
        """uint8 [M][N] ids -> EncodedMatrix with `coder`'s packing."""
        m = cls.__new__(cls)
        m.coder = coder
        m.codes = np.ascontiguousarray(planes, np.uint8 if coder.width <= 8 else np.uint16)
        if m.codes.ndim != 2:
            raise ValueError("expected codes [M][N]")
        return m

    @property
    def encodings(self):
        """Vector[coder.Code]: the packed plane of every quantizer."""
        return [self.coder.build_code(self.codes[m]) for m in range(self.codes.shape[0])]

    @property
    def unwrapped_encodings(self):
        return [self.coder.unwrap_code(c) for c in self.encodings]

    @property
    def length(self):
        return self.codes.shape[1]

    def __call__(self, row):
        """EncodedVector: the M centroid ids of one row."""
        return self.codes[:, row].astype(np.int32)

    def __eq__(self, other):
        return isinstance(other, EncodedMatrix) and self.coder == other.coder and \
            np.array_equal(self.codes, other.codes)

    __hash__ = None


@dataclass
class Quantizer:
    """ProductQuantizer.Quantizer(from, clusters), G/ProductQuantizer.scala:80-86."""
    from_: int
    clusters: KMeans

    @property
    def dimension(self):
        return self.clusters.dimension


@dataclass
class ProgressReport:
    """ProductQuantizer.ProgressReport, G/ProductQuantizer.scala:113-119."""
    kmeans_reports: List[KMeansProgressReport]

    @property
    def completed_iterations(self):
        return sum(r.num_iterations for r in self.kmeans_reports)

    @property
    def total_iterations(self):
        return sum(r.max_iterations for r in self.kmeans_reports)


@dataclass
class Config:
    """ProductQuantizer.Config, G/ProductQuantizer.scala:107-111."""
    num_clusters: int
    num_quantizers: int
    max_iterations: int
    report: Optional[Callable[[ProgressReport], None]] = None
    update_mode: int = N.UPDATE_RUNNING_MEAN


class ProductQuantizer:
    """G/ProductQuantizer.scala.  Owns a device-resident codebook handle (gulon_codebook_t)."""

    def __init__(self, num_clusters, quantizers, _handle=None):
        self.num_clusters = int(num_clusters)
        self.quantizers = list(quantizers)
        coder_width(self.num_clusters)
        self.dimension = sum(q.dimension for q in self.quantizers)
        self._handle = _handle
        if _handle is None:
            M = len(self.quantizers)
            frm, dim, dmax = subvector_windows(self.dimension, M)
            for q, f, d in zip(self.quantizers, frm, dim):
                if q.from_ != f or q.dimension != d:
                    raise ValueError("quantizer windows must follow Vectors.subvectors")
            cb = np.zeros((M, self.num_clusters, dmax), np.float32)
            for m, q in enumerate(self.quantizers):
                if q.clusters.k != self.num_clusters:
                    raise ValueError("every quantizer needs numClusters centroids")
                cb[m, :, :q.dimension] = q.clusters.centroids
            h = N.vp()
            N.check(N.lib().gulon_codebook_create(self.dimension, M, self.num_clusters,
                                                  cb.ctypes.data, C.byref(h)))
            self._handle = h

    @property
    def handle(self):
        return self._handle

    @property
    def coder_factory(self):
        """ProductQuantizer#coderFactory, G/ProductQuantizer.scala:11-16."""
        return factory_for(max_width(self.num_clusters))

    def __del__(self):
        try:
            if self._handle:
                N.lib().gulon_codebook_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def from_codebook(cls, codebook, D):
        """codebook float32 [M][K][dmax] laid out by the split rule."""
        cb = np.ascontiguousarray(codebook, np.float32)
        M, K, _ = cb.shape
        frm, dim, dmax = subvector_windows(D, M)
        if cb.shape[2] != dmax:
            raise ValueError("codebook last dimension must be ceil(D / M) = %d" % dmax)
        qs = [Quantizer(int(f), KMeans(int(d), cb[m, :, :d].copy()))
              for m, (f, d) in enumerate(zip(frm, dim))]
        return cls(K, qs)

    @classmethod
    def _from_handle(cls, h):
        D, M, K, dmax = N.i32(0), N.i32(0), N.i32(0), N.i32(0)
        N.check(N.lib().gulon_codebook_info(h, C.byref(D), C.byref(M), C.byref(K), C.byref(dmax)))
        cb = np.zeros((M.value, K.value, dmax.value), np.float32)
        N.check(N.lib().gulon_codebook_export(h, cb.ctypes.data))
        frm, dim, _ = subvector_windows(D.value, M.value)
        qs = [Quantizer(int(f), KMeans(int(d), cb[m, :, :d].copy()))
              for m, (f, d) in enumerate(zip(frm, dim))]
        return cls(K.value, qs, _handle=h)

    @classmethod
    def train(cls, vectors, config: Config, comm=None, n_total=0, row_offset=0):
        """ProductQuantizer.apply(vectors, config), G/ProductQuantizer.scala:150-153: M independent
        k-means, seed = quantizer index (:139).  `vectors` is a Matrix or DevicePoints."""
        dev = vectors.device() if isinstance(vectors, Matrix) else vectors
        coder_width(config.num_clusters)
        reports = [KMeansProgressReport(0, config.max_iterations, 0.0, 0.0, False, m)
                   for m in range(config.num_quantizers)]
        user_report = config.report

        def on_report(r):
            reports[r.quantizer] = r
            user_report(ProgressReport(list(reports)))

        cb = _progress_cb(on_report if user_report is not None else None)
        h = N.vp()
        N.check(N.lib().gulon_pq_train(dev.handle, config.num_quantizers, config.num_clusters,
                                       config.max_iterations, N.TIE_LOWEST, config.update_mode,
                                       C.byref(comm) if comm is not None else None, n_total,
                                       row_offset, cb, None, C.byref(h)))
        return cls._from_handle(h)

    apply = train

    def codebook(self):
        M = len(self.quantizers)
        _, _, dmax = subvector_windows(self.dimension, M)
        cb = np.zeros((M, self.num_clusters, dmax), np.float32)
        N.check(N.lib().gulon_codebook_export(self._handle, cb.ctypes.data))
        return cb

    # -- encode / decode ------------------------------------------------------------------
    def encode(self, vectors):
        """ProductQuantizer#encode, G/ProductQuantizer.scala:25-35 (host buffers in and out)."""
        x = vectors.data if isinstance(vectors, Matrix) else np.ascontiguousarray(vectors, np.float32)
        if x.ndim != 2 or x.shape[1] != self.dimension:
            raise ValueError("expected [N][%d] vectors" % self.dimension)
        M = len(self.quantizers)
        wide = self.num_clusters > 256
        codes = np.zeros((M, x.shape[0]), np.uint16 if wide else np.uint8)
        fn = N.lib().gulon_pq_encode16 if wide else N.lib().gulon_pq_encode
        N.check(fn(self._handle, x.ctypes.data, x.shape[0], x.shape[1], N.TIE_LOWEST, codes.ctypes.data))
        return EncodedMatrix.from_planes(self.coder_factory(x.shape[0]), codes)

    def encode_dev(self, x, out=None, stream=None):
        """Device-resident encode: x CUDA float32 [N][D] -> CUDA uint8 [M][stride] (torch tensors)."""
        import torch
        n = x.shape[0]
        M = len(self.quantizers)
        stride = (n + 15) // 16 * 16
        wide = self.num_clusters > 256
        if out is None:
            out = torch.zeros((M, max(stride, 16)), dtype=torch.uint16 if wide else torch.uint8,
                              device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        ld = x.stride(0) if n > 1 else max(x.shape[1], 1)
        fn = N.lib().gulon_pq_encode16_dev if wide else N.lib().gulon_pq_encode_dev
        N.check(fn(self._handle, x.data_ptr(), n, ld, N.TIE_LOWEST, out.data_ptr(), out.stride(0), st))
        return out

    def decode(self, encoded):
        """ProductQuantizer#decode(EncodedMatrix), G/ProductQuantizer.scala:58-78; with a single
        EncodedVector (ids [M]) decodes one row (:37-52)."""
        if isinstance(encoded, EncodedMatrix):
            codes = encoded.codes
            one = False
        else:
            codes = np.ascontiguousarray(encoded).reshape(-1, 1)
            one = True
        wide = self.num_clusters > 256
        n = codes.shape[1]
        out = np.zeros((n, self.dimension), np.float32)
        codes = np.ascontiguousarray(codes.astype(np.uint16 if wide else np.uint8))
        fn = N.lib().gulon_pq_decode16 if wide else N.lib().gulon_pq_decode
        N.check(fn(self._handle, codes.ctypes.data, n, n, out.ctypes.data, self.dimension))
        return out[0] if one else Matrix(out)
