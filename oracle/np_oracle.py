"""Independent numpy-float32 restatement of the reference's hot path, for TINY cases.

TEST INFRASTRUCTURE ONLY (same rule as gulon_oracle.c).  PARITY UNPINNED.
Written separately from the C oracle, straight from the Scala sources, so that the two
restatements cross-check each other bit for bit (tests/test_oracle_cross.py).
G/ = core/src/main/scala/net/tixxit/gulon/.
"""
import math

import numpy as np

f32 = np.float32


class JavaRandom:
    """java.util.Random, from the JDK javadoc."""

    def __init__(self, seed):
        self.s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def _next(self, bits):
        self.s = (self.s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.s >> (48 - bits)
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v >= (1 << 31) else v

    def next_int(self, bound=None):
        if bound is None:
            return self._next(32)
        r = self._next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            t = (u - r + m) & 0xFFFFFFFF
            if t < (1 << 31):
                return r
            u = self._next(31)

    def next_boolean(self):
        return self._next(1) != 0

    def next_float(self):
        return f32(self._next(24)) / f32(1 << 24)


def subvectors(D, M):
    """G/Vectors.scala:84-104."""
    ideal = (D + M - 1) // M
    short = ideal * M - D
    full = M - short
    out = []
    for i in range(M):
        if i < full:
            out.append((i * ideal, i * ideal + ideal))
        else:
            fr = full * ideal + (i - full) * (ideal - 1)
            out.append((fr, fr + ideal - 1))
    return out


def offsets(Cm):
    """G/KMeans.scala:170-186."""
    off = np.zeros(len(Cm), f32)
    for i, c in enumerate(Cm):
        s = f32(0)
        for x in c:
            s = f32(s + f32(x * x))
        off[i] = s
    return off


def assign(X, frm, until, Cm, start=0, end=None, literal=True, prev=None):
    """G/KMeans.scala:24-55 / :70-98 with one Random(0) over [start, end)."""
    X = np.asarray(X, f32)
    Cm = np.asarray(Cm, f32)
    end = len(X) if end is None else end
    off = offsets(Cm)
    out = np.zeros(len(X), np.int32) if prev is None else np.array(prev, np.int32)
    rng = JavaRandom(0)
    fmax = np.finfo(f32).max
    for i in range(start, end):
        row = X[i]
        mn = fmax
        for k in range(len(Cm)):
            c = Cm[k]
            d = f32(0)
            for j in range(len(c)):
                d = f32(d + f32(row[j + frm] * c[j]))
            d = f32(off[k] - f32(f32(2) * d))
            if d < mn or (d == mn and (rng.next_boolean() if literal else False)):
                out[i] = k
                mn = d
    return out


def par_assign(X, frm, until, Cm, literal=True, batch=25000):
    """G/KMeans.scala:57-68."""
    out = np.zeros(len(X), np.int32)
    for s in range(0, len(X), batch):
        out = assign(X, frm, until, Cm, s, min(len(X), s + batch), literal, out)
    return out


def from_assignment(X, frm, until, a, K):
    """G/KMeans.scala:198-226."""
    X = np.asarray(X, f32)
    dim = until - frm
    Cm = np.zeros((K, dim), f32)
    cnt = np.zeros(K, np.int32)
    for i in range(len(X)):
        k = a[i]
        n = cnt[k] + 1
        for j in range(dim):
            p = Cm[k, j]
            Cm[k, j] = f32(p + f32(f32(X[i, j + frm] - p) / f32(n)))
        cnt[k] = n
    return Cm


def kmeans_init(X, frm, until, K, seed=0):
    """G/KMeans.scala:188-196."""
    rng = JavaRandom(seed)
    return np.array([np.asarray(X, f32)[rng.next_int(len(X)), frm:until] for _ in range(K)], f32)


def compute_clusters(X, frm, until, K, max_iter, seed=0, literal=True):
    """G/KMeans.scala:134-157."""
    cur = kmeans_init(X, frm, until, K, seed)
    prev_a = par_assign(X, frm, until, cur, literal)
    i, updates, conv = 0, 0, False
    while i <= max_iter:
        cur = from_assignment(X, frm, until, prev_a, K)
        nxt = par_assign(X, frm, until, cur, literal)
        conv = bool(np.array_equal(prev_a, nxt))
        updates += 1
        prev_a = nxt
        i = max_iter + 1 if conv else i + 1
    return cur, updates, conv, prev_a


def pq_encode(X, codebooks, literal=True):
    """G/ProductQuantizer.scala:25-35 + the coder's getIndex(buildCode(.)) round trip (ids unchanged for
    ids < numClusters); codebooks = list of (from, centroids).  One byte per id up to 256 centroids."""
    planes = []
    for frm, Cm in codebooks:
        a = assign(X, frm, frm + Cm.shape[1], Cm, literal=literal)
        planes.append(a.astype(np.uint8 if len(Cm) <= 256 else np.uint16))
    return np.stack(planes)


def prepare_query(queries, codebooks):
    """G/Index.scala:352-383."""
    queries = np.asarray(queries, f32)
    K = len(codebooks[0][1])
    lut = np.zeros((len(queries), len(codebooks), K), f32)
    for j, (frm, Cm) in enumerate(codebooks):
        for i, c in enumerate(Cm):
            for q in range(len(queries)):
                s = f32(0)
                for k in range(len(c)):
                    d = f32(queries[q, k + frm] - c[k])
                    s = f32(s + f32(d * d))
                lut[q, j, i] = s
    return lut


class TopKHeap:
    """G/TopKHeap.scala, literal."""

    def __init__(self, k):
        self.keys = [0] * k
        self.values = [f32(0)] * k
        self.size = 0

    def _swap(self, i, j):
        self.keys[i], self.keys[j] = self.keys[j], self.keys[i]
        self.values[i], self.values[j] = self.values[j], self.values[i]

    def _up(self, i):
        if i > 0:
            p = (i - 1) // 2
            if self.values[i] > self.values[p]:
                self._swap(i, p)
                self._up(p)

    def _down(self, i):
        top, lc, rc = i, 2 * i + 1, 2 * i + 2
        if lc < self.size and self.values[top] < self.values[lc]:
            top = lc
        if rc < self.size and self.values[top] < self.values[rc]:
            top = rc
        if top != i:
            self._swap(i, top)
            self._down(top)

    def delete(self):
        if self.size <= 0:
            raise RuntimeError("heap is empty")
        self.size -= 1
        removed = self.keys[0]
        self.keys[0] = self.keys[self.size]
        self.values[0] = self.values[self.size]
        self._down(0)
        return removed

    def update(self, k, v):
        if self.size == len(self.keys) and self.size > 0 and self.values[0] > v:
            self.delete()
        if self.size < len(self.keys):
            self.keys[self.size] = k
            self.values[self.size] = v
            self._up(self.size)
            self.size += 1

    def merge(self, that):
        for i in range(that.size):
            self.update(that.keys[i], that.values[i])

    def drain(self):
        """Result.fromHeap, G/Index.scala:83-94."""
        n = self.size
        ids, ds = [0] * n, [f32(0)] * n
        for i in range(n - 1, -1, -1):
            ids[i], ds[i] = self.keys[0], self.values[0]
            self.delete()
        return np.array(ids, np.int32), np.array(ds, f32)


def batch_query(lut, codes, k, frm=0, until=None, literal=True):
    """G/Index.scala:393-440."""
    lut = np.asarray(lut, f32)
    until = codes.shape[1] if until is None else until
    out = []
    for q in range(len(lut)):
        heap = TopKHeap(k)
        pairs = []
        i = frm
        while i < until:
            bs = min(4096, until - i)
            ds = np.zeros(bs, f32)
            for j in range(codes.shape[0]):
                ds = (ds + lut[q, j][codes[j, i:i + bs]]).astype(f32)
            for r in range(bs):
                if literal:
                    heap.update(i + r, ds[r])
                else:
                    pairs.append((ds[r], i + r))
            i += bs
        if literal:
            out.append(heap.drain())
        else:
            pairs.sort()
            pairs = pairs[:k]
            out.append((np.array([p[1] for p in pairs], np.int32),
                        np.array([p[0] for p in pairs], f32)))
    return out


def distance_sq(x, y):
    """G/MathUtils.scala:85-95."""
    s = f32(0)
    for i in range(len(x)):
        dx = f32(f32(y[i]) - f32(x[i]))
        s = f32(s + f32(dx * dx))
    return s


def exact_nn(X, query, k, frm=0, until=None, literal=True):
    """G/Index.scala:209-229."""
    until = len(X) if until is None else until
    if literal:
        heap = TopKHeap(k)
        for i in range(frm, until):
            heap.update(i, distance_sq(X[i], query))
        return heap.drain()
    pairs = sorted((distance_sq(X[i], query), i) for i in range(frm, until))[:k]
    return np.array([p[1] for p in pairs], np.int32), np.array([p[0] for p in pairs], f32)


def normalize(x):
    """G/MathUtils.scala:100-120."""
    x = np.asarray(x, f32)
    s = f32(0)
    for v in x:
        s = f32(s + f32(v * v))
    d = f32(math.sqrt(float(s)))
    return np.array([f32(v / d) for v in x], f32)


# ---- Coder (G/Coder.scala), literal per-index loops ------------------------------------------------
def _to_byte(v):
    return v & 0xFF


def coder_bytes_per_code(width, length):
    """BytePackedCoder.bytesPerCode, G/Coder.scala:82-83."""
    cpb = 8 // width
    return (length + cpb - 1) // cpb


def _packed_build(width, code, indices, offset):
    """Coder2/4/8#buildCodeWithOffset, G/Coder.scala:100-108,115-123,130-136."""
    for i, v in enumerate(indices):
        if width == 2:
            j = offset + (i >> 2)
            code[j] = _to_byte(code[j] | ((v & 0x3) << ((i & 0x3) * 2)))
        elif width == 4:
            j = offset + (i >> 1)
            code[j] = _to_byte(code[j] | ((v & 0xF) << ((i & 0x1) * 4)))
        else:
            code[offset + i] = _to_byte(v)


def _packed_get(width, code, offset, i):
    """Coder2/4/8#getIndexWithOffset, G/Coder.scala:110-111,125-126,138-139."""
    if width == 2:
        return (code[offset + (i >> 2)] >> ((i & 0x3) * 2)) & 0x3
    if width == 4:
        return (code[offset + (i >> 1)] >> ((i & 0x1) * 4)) & 0xF
    return code[offset + i] & 0xFF


def coder_supported_width(width):
    """Coder.factoryFor, G/Coder.scala:35-45 -> the width actually used, or None."""
    if width < 0 or width > 16:
        return None
    for w in (0, 2, 4, 8, 10, 12, 16):
        if width <= w:
            return w
    return None


def coder_build(width, length, indices):
    """Coder(width, length).buildCode(indices) -> list of byte values (None for width 0)."""
    w = coder_supported_width(width)
    if w is None:
        raise ValueError("unsupported width: %d" % width)
    if w == 0:
        return None
    if w <= 8:
        code = [0] * coder_bytes_per_code(w, length)
        _packed_build(w, code, indices, 0)
        return code
    lw = w - 8                                                 # BytePlus, G/Coder.scala:147-161
    if len(indices) != length:
        raise ValueError("indices.length != %d" % length)
    code = [0] * (length + coder_bytes_per_code(lw, length))
    for i in range(length):
        code[i] = _to_byte((indices[i] & 0xFFFFFFFF) >> lw)
    _packed_build(lw, code, indices, length)
    return code


def coder_get_index(width, length, code, i):
    """Coder(width, length).getIndex(code, i), G/Coder.scala:68-72,93-94,163-167."""
    w = coder_supported_width(width)
    if w == 0:
        if i < 0 or i >= length:
            raise IndexError(str(i))
        return 0
    if w <= 8:
        return _packed_get(w, code, 0, i)
    lw = w - 8
    return ((code[i] & 0xFF) << lw) | (_packed_get(lw, code, length, i) & 0xFF)
