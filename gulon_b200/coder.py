"""Coder: how the per-quantizer centroid ids of N rows are packed into one byte array
(G/Coder.scala:12-168).  A `Code` here is a numpy uint8 array (None for Coder0); bit layouts are the
reference's, so packed planes can be exchanged with it byte for byte (index.proto `encodings`).

Supported widths follow Coder.factoryFor (G/Coder.scala:35-45): 0, 2, 4, 8 and the BytePlus widths
10, 12, 16 (one most-significant byte per id followed by the packed low bits).  The scan kernels
work on one byte per id (K <= 256), whatever the packed width: `unpack` / `pack` convert at the
boundary; ids wider than 8 bits exist in the packed form only.
"""
import numpy as np

SUPPORTED_WIDTHS = [2, 4, 8, 10, 12, 16]           # Coder.supportedWidths, G/Coder.scala:28-29


class Coder0:
    """G/Coder.scala:62-73: a single centroid, nothing stored."""
    width = 0

    def __init__(self, length):
        self.length = int(length)

    def build_code(self, indices):
        return None

    def get_index(self, code, i):
        if i < 0 or i >= self.length:
            raise IndexError(str(i))
        return 0

    def wrap_code(self, encoded):
        return None

    def unwrap_code(self, code):
        return np.zeros(0, np.uint8)

    def unpack(self, code):
        return np.zeros(self.length, np.int32)

    def __eq__(self, other):
        return type(other) is type(self) and other.length == self.length

    def __hash__(self):
        return hash((type(self).__name__, self.length))


class _BytePacked:
    """BytePackedCoder(width), G/Coder.scala:81-97: 8 / width ids per byte, low bits first."""
    width = 8

    def __init__(self, length):
        self.length = int(length)
        self.codes_per_byte = 8 // self.width
        self.bytes_per_code = (self.length + self.codes_per_byte - 1) // self.codes_per_byte

    # buildCodeWithOffset / getIndexWithOffset, vectorised over the whole array
    def _pack_into(self, code, indices, offset):
        ids = np.asarray(indices).astype(np.int64) & ((1 << self.width) - 1)
        cpb = self.codes_per_byte
        pad = (-len(ids)) % cpb
        if pad:
            ids = np.concatenate((ids, np.zeros(pad, np.int64)))
        lanes = ids.reshape(-1, cpb) << (np.arange(cpb, dtype=np.int64) * self.width)
        packed = np.bitwise_or.reduce(lanes, axis=1).astype(np.uint8)
        code[offset:offset + len(packed)] |= packed

    def _unpack_from(self, raw, offset, n):
        cpb = self.codes_per_byte
        nb = (n + cpb - 1) // cpb
        b = np.asarray(raw[offset:offset + nb], np.uint8).astype(np.int32)
        lanes = (b[:, None] >> (np.arange(cpb, dtype=np.int32) * self.width)) & ((1 << self.width) - 1)
        return lanes.reshape(-1)[:n]

    def build_code(self, indices):
        code = np.zeros(self.bytes_per_code, np.uint8)
        self._pack_into(code, indices, 0)
        return code

    def get_index(self, code, i):
        return self._get_with_offset(code, 0, i)

    def _get_with_offset(self, raw, offset, i):
        cpb = self.codes_per_byte
        if cpb == 1:
            return int(raw[offset + i]) & 0xFF
        sh = {4: 2, 2: 1}[cpb]
        return (int(raw[offset + (i >> sh)]) >> ((i & (cpb - 1)) * self.width)) & ((1 << self.width) - 1)

    def wrap_code(self, encoded):
        return np.frombuffer(bytes(encoded), np.uint8).copy() if not isinstance(encoded, np.ndarray) \
            else np.ascontiguousarray(encoded, np.uint8)

    def unwrap_code(self, code):
        return np.ascontiguousarray(code, np.uint8)

    def unpack(self, code):
        """all N ids of a packed plane (the inverse of build_code)."""
        return self._unpack_from(np.asarray(code, np.uint8), 0, self.length)

    def __eq__(self, other):
        return type(other) is type(self) and other.length == self.length

    def __hash__(self):
        return hash((type(self).__name__, self.length))


class Coder2(_BytePacked):
    """G/Coder.scala:99-112."""
    width = 2


class Coder4(_BytePacked):
    """G/Coder.scala:114-127."""
    width = 4


class Coder8(_BytePacked):
    """G/Coder.scala:129-140: one byte per index (`code[i] = (byte) idx`)."""
    width = 8

    def build_code(self, indices):
        a = np.asarray(indices)
        if a.ndim != 1 or a.shape[0] != self.length:
            raise IndexError("expected %d indices" % self.length)
        return a.astype(np.int64).astype(np.uint8)            # toByte: the low 8 bits

    def get_index(self, code, i):
        return int(code[i]) & 0xFF

    def unpack(self, code):
        return np.asarray(code, np.uint8).astype(np.int32)[:self.length]


class BytePlus:
    """BytePlus(lsb), G/Coder.scala:142-168: `length` most-significant bytes, then the low
    `lsb.width` bits packed by `lsb`."""

    def __init__(self, lsb):
        self.lsb = lsb
        self.length = lsb.length
        self.width = lsb.width + 8

    def build_code(self, indices):
        a = np.asarray(indices).astype(np.int64)
        if a.ndim != 1 or a.shape[0] != self.length:
            raise ValueError("indices.length != %d" % self.length)
        code = np.zeros(self.length + self.lsb.bytes_per_code, np.uint8)
        code[:self.length] = ((a & 0xFFFFFFFF) >> self.lsb.width).astype(np.uint8)   # >>> then toByte
        self.lsb._pack_into(code, a, self.length)
        return code

    def get_index(self, code, i):
        b1 = (int(code[i]) & 0xFF) << self.lsb.width
        b0 = self.lsb._get_with_offset(code, self.length, i) & 0xFF
        return b1 | b0

    def wrap_code(self, encoded):
        return np.frombuffer(bytes(encoded), np.uint8).copy() if not isinstance(encoded, np.ndarray) \
            else np.ascontiguousarray(encoded, np.uint8)

    def unwrap_code(self, code):
        return np.ascontiguousarray(code, np.uint8)

    def unpack(self, code):
        c = np.asarray(code, np.uint8)
        hi = c[:self.length].astype(np.int32) << self.lsb.width
        return hi | self.lsb._unpack_from(c, self.length, self.length)

    def __eq__(self, other):
        return isinstance(other, BytePlus) and other.lsb == self.lsb

    def __hash__(self):
        return hash(("BytePlus", self.lsb))


class Factory:
    """Coder.Factory(width, make), G/Coder.scala:30-33."""

    def __init__(self, width, make):
        self.width = width
        self.make = make

    @property
    def k(self):
        return 1 << self.width

    def __call__(self, length):
        return self.make(length)


def factory_for(width):
    """Coder.factoryFor, G/Coder.scala:35-45: the narrowest supported coder holding `width` bits."""
    if width < 0:
        return None
    if width == 0:
        return Factory(0, Coder0)
    if width <= 2:
        return Factory(2, Coder2)
    if width <= 4:
        return Factory(4, Coder4)
    if width <= 8:
        return Factory(8, Coder8)
    if width <= 10:
        return Factory(10, lambda n: BytePlus(Coder2(n)))
    if width <= 12:
        return Factory(12, lambda n: BytePlus(Coder4(n)))
    if width <= 16:
        return Factory(16, lambda n: BytePlus(Coder8(n)))
    return None


def coder(width, length):
    """Coder.apply(width, length), G/Coder.scala:54-58."""
    f = factory_for(width)
    if f is None:
        raise ValueError("unsupported width: %d" % width)
    return f(length)


def max_width(num_clusters):
    """32 - numberOfLeadingZeros(numClusters - 1), G/ProductQuantizer.scala:12 (Int arithmetic:
    numClusters <= 0 gives 32)."""
    v = (int(num_clusters) - 1) & 0xFFFFFFFF
    return v.bit_length()
