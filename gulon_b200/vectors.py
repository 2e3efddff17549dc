"""Matrix / Vectors: the boundary types of the hot path.

Reference: G/Matrix.scala:3 (`Matrix(rows, cols, data: Array[Array[Float]])`, jagged) and
G/Vectors.scala:3,84-104 (`Vectors(matrix, from, until)` column window + the `subvectors` split
rule).  Here a Matrix is one flat row-major float32 array (the C ABI's layout); it lazily owns a
device-resident copy (`gulon_points_t`) so repeated calls do not re-upload.
"""
import ctypes as C

import numpy as np

from . import _native as N


def subvector_windows(D, M):
    """Vectors.subvectors split rule (G/Vectors.scala:84-104) -> (from[M], dim[M], dmax)."""
    frm = np.zeros(M, np.int32)
    dim = np.zeros(M, np.int32)
    dmax = N.check(N.lib().gulon_subvectors(int(D), int(M), frm.ctypes.data, dim.ctypes.data))
    return frm, dim, int(dmax)


class DevicePoints:
    """Owner of a gulon_points_t handle."""

    def __init__(self, handle, keepalive=None):
        self.handle = handle
        self._keepalive = keepalive

    @classmethod
    def from_host(cls, data):
        data = np.ascontiguousarray(data, np.float32)
        if data.ndim != 2:
            raise ValueError("expected a 2-d matrix")
        h = N.vp()
        N.check(N.lib().gulon_points_create(data.ctypes.data, data.shape[0], max(data.shape[1], 0),
                                            max(data.shape[1], 1), C.byref(h)))
        return cls(h)

    @classmethod
    def from_torch(cls, t):
        """Borrow a CUDA float32 tensor [N][D] (row stride = t.stride(0))."""
        if not t.is_cuda or t.dim() != 2 or t.dtype.is_floating_point is False or t.element_size() != 4:
            raise ValueError("expected a 2-d float32 CUDA tensor")
        if t.shape[1] > 1 and t.stride(1) != 1:
            raise ValueError("rows must be contiguous")
        h = N.vp()
        ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
        N.check(N.lib().gulon_points_wrap_dev(t.data_ptr(), t.shape[0], t.shape[1], ld, C.byref(h)))
        return cls(h, keepalive=t)

    def normalize(self):
        N.check(N.lib().gulon_points_normalize(self.handle))

    def info(self):
        n, d, ld = C.c_int64(0), C.c_int32(0), C.c_int64(0)
        N.check(N.lib().gulon_points_info(self.handle, C.byref(n), C.byref(d), C.byref(ld), None))
        return int(n.value), int(d.value), int(ld.value)

    @property
    def rows(self):
        return self.info()[0]

    @property
    def cols(self):
        return self.info()[1]

    def __del__(self):
        try:
            if self.handle:
                N.lib().gulon_points_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Matrix:
    """G/Matrix.scala:3."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(data, np.float32)
        if self.data.ndim != 2:
            raise ValueError("expected a 2-d matrix")
        self._dev = None

    @property
    def rows(self):
        return self.data.shape[0]

    @property
    def cols(self):
        return self.data.shape[1]

    def device(self):
        if self._dev is None:
            self._dev = DevicePoints.from_host(self.data)
        return self._dev


class Vectors:
    """G/Vectors.scala:3: the column window [from, until) of a matrix (no copy)."""

    def __init__(self, matrix, from_=0, until=None):
        if not isinstance(matrix, Matrix):
            matrix = Matrix(matrix)
        self.matrix = matrix
        self.from_ = int(from_)
        self.until = matrix.cols if until is None else int(until)
        if not (0 <= self.from_ <= self.until <= matrix.cols):
            raise ValueError("invalid window [%d, %d) of %d columns" % (self.from_, self.until, matrix.cols))

    @property
    def dimension(self):
        return self.until - self.from_

    @property
    def size(self):
        return self.matrix.rows

    def to_array(self):
        return self.matrix.data[:, self.from_:self.until]

    @staticmethod
    def subvectors(matrix, n):
        """G/Vectors.scala:84-104."""
        if not isinstance(matrix, Matrix):
            matrix = Matrix(matrix)
        frm, dim, _ = subvector_windows(matrix.cols, n)
        return [Vectors(matrix, int(f), int(f + d)) for f, d in zip(frm, dim)]


def normalize(x):
    """MathUtils.normalize per row (G/MathUtils.scala:100-120), computed on the device."""
    x = np.ascontiguousarray(x, np.float32)
    one = x.ndim == 1
    x2 = x.reshape(1, -1) if one else x
    out = np.empty_like(x2)
    N.check(N.lib().gulon_normalize(x2.ctypes.data, x2.shape[0], x2.shape[1], x2.shape[1],
                                    out.ctypes.data, x2.shape[1]))
    return out[0] if one else out
