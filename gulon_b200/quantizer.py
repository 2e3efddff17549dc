"""ProductQuantizer / Coder / EncodedMatrix: host-side mirror of G/ProductQuantizer.scala,
G/Coder.scala (Coder8 only: 256 centroids => width 8) and G/EncodedMatrix.scala over the C ABI.

Codes are plane-major: one uint8 array of length N per quantizer (EncodedMatrix.encodings,
G/EncodedMatrix.scala:11-23), held here as one uint8 [M][N] array.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from . import _native as N
from .kmeans import KMeans, ProgressReport as KMeansProgressReport, _progress_cb
from .vectors import Matrix, subvector_windows


def coder_width(num_clusters):
    """ProductQuantizer.coderFactory, G/ProductQuantizer.scala:11-16: 32 - nlz(K - 1)."""
    w = max(int(num_clusters) - 1, 0).bit_length()
    if w > 8:
        # G/Coder.scala:35-45 also has 10/12/16-bit coders; this library ships Coder8 only.
        raise ValueError("too many clusters: %d (this build supports Coder8: K <= 256)" % num_clusters)
    return w


class Coder8:
    """G/Coder.scala:129-140: one byte per index."""
    width = 8

    def __init__(self, length):
        self.length = int(length)

    def build_code(self, indices):
        a = np.asarray(indices)
        if a.shape[0] != self.length:
            raise IndexError("expected %d indices" % self.length)
        return a.astype(np.uint8)

    @staticmethod
    def get_index(code, i):
        return int(code[i]) & 0xFF


class EncodedMatrix:
    """G/EncodedMatrix.scala:11-23."""

    def __init__(self, coder, encodings):
        self.coder = coder
        self.codes = np.ascontiguousarray(encodings, np.uint8)
        if self.codes.ndim != 2:
            raise ValueError("expected codes [M][N]")

    @property
    def encodings(self):
        return [self.codes[m] for m in range(self.codes.shape[0])]

    @property
    def length(self):
        return self.codes.shape[1]

    def __call__(self, row):
        """EncodedVector: the M centroid ids of one row."""
        return self.codes[:, row].astype(np.int32)


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
        codes = np.zeros((M, x.shape[0]), np.uint8)
        N.check(N.lib().gulon_pq_encode(self._handle, x.ctypes.data, x.shape[0], x.shape[1],
                                        N.TIE_LOWEST, codes.ctypes.data))
        return EncodedMatrix(Coder8(x.shape[0]), codes)

    def encode_dev(self, x, out=None, stream=None):
        """Device-resident encode: x CUDA float32 [N][D] -> CUDA uint8 [M][stride] (torch tensors)."""
        import torch
        n = x.shape[0]
        M = len(self.quantizers)
        stride = (n + 15) // 16 * 16
        if out is None:
            out = torch.zeros((M, max(stride, 16)), dtype=torch.uint8, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        ld = x.stride(0) if n > 1 else max(x.shape[1], 1)
        N.check(N.lib().gulon_pq_encode_dev(self._handle, x.data_ptr(), n, ld, N.TIE_LOWEST,
                                            out.data_ptr(), out.stride(0), st))
        return out

    def decode(self, encoded):
        """ProductQuantizer#decode(EncodedMatrix), G/ProductQuantizer.scala:58-78; with a single
        EncodedVector (ids [M]) decodes one row (:37-52)."""
        if isinstance(encoded, EncodedMatrix):
            codes = encoded.codes
            one = False
        else:
            codes = np.ascontiguousarray(encoded).astype(np.uint8).reshape(-1, 1)
            one = True
        n = codes.shape[1]
        out = np.zeros((n, self.dimension), np.float32)
        codes = np.ascontiguousarray(codes)
        N.check(N.lib().gulon_pq_decode(self._handle, codes.ctypes.data, n, n, out.ctypes.data,
                                        self.dimension))
        return out[0] if one else Matrix(out)
