"""ctypes wrapper over oracle/libgulon_oracle.so (the C restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of gulon_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs; never by gulon_b200/.
PARITY UNPINNED (no JVM in this image, no golden vectors in the reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgulon_oracle.so")

TIE_LITERAL, TIE_LOWEST = 0, 1
TOPK_LITERAL, TOPK_CANONICAL = 0, 1


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("gulon_oracle.c", "go_codes.inc", "synth.c",
                                             "../gulon_b200/csrc/synth_spec.h")]
    def stale():
        return not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs)

    if force or stale():
        import fcntl
        with open(_SO + ".lock", "w") as lock:          # several processes may find it stale at once
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.go_jr_next_float.restype = C.c_float
        _lib.go_distance_sq.restype = C.c_float
        _lib.go_objective.restype = C.c_double
        _lib.go_heap_new.restype = C.c_void_p
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


i64 = C.c_int64
i32 = C.c_int32


def num_threads():
    """OpenMP's default team size -- obeys OMP_NUM_THREADS (torchrun exports 1); prefer host_cores()."""
    return int(lib().go_num_threads())


def host_cores():
    """Host cores this process may run on; every oracle call takes an explicit `nthreads`."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class JRandom:
    """java.util.Random."""

    def __init__(self, seed):
        self._s = C.c_uint64(0)
        lib().go_jr_init(C.byref(self._s), i64(seed))

    def next_int(self, bound=None):
        if bound is None:
            return int(i32(lib().go_jr_next_int(C.byref(self._s))).value)
        return int(lib().go_jr_next_int_bound(C.byref(self._s), i32(bound)))

    def next_boolean(self):
        return bool(lib().go_jr_next_boolean(C.byref(self._s)))

    def next_float(self):
        return float(lib().go_jr_next_float(C.byref(self._s)))


def subvectors(D, M):
    frm = np.zeros(M, np.int32)
    dim = np.zeros(M, np.int32)
    dmax = lib().go_subvectors(int(D), int(M), _p(frm), _p(dim))
    return frm, dim, int(dmax)


def offsets(Cm):
    Cm = _f32(Cm)
    K, dim = Cm.shape
    off = np.zeros(K, np.float32)
    lib().go_offsets(_p(Cm), K, dim, i64(dim), _p(off))
    return off


def assign(X, frm, dim, Cm, batch=0, tie_mode=TIE_LITERAL, nthreads=1, prev=None, stats=None):
    """KMeans.assign (batch=0, whole array) / parAssign (batch=25000)."""
    X = _f32(X)
    Cm = _f32(Cm)
    N, ld = X.shape
    K = Cm.shape[0]
    out = np.zeros(N, np.int32) if prev is None else np.ascontiguousarray(prev, np.int32).copy()
    st = np.zeros(2, np.int64)
    lib().go_assign(_p(X), i64(N), i64(ld), int(frm), int(dim), _p(Cm), i64(Cm.shape[1]), K,
                    i64(batch), int(tie_mode), int(nthreads), _p(out), _p(st))
    if stats is not None:
        stats += st
    return out


def from_assignment(X, frm, dim, a, K, return_counts=False):
    X = _f32(X)
    N, ld = X.shape
    a = np.ascontiguousarray(a, np.int32)
    Cm = np.zeros((K, dim), np.float32)
    cnt = np.zeros(K, np.int32)
    lib().go_from_assignment(_p(X), i64(N), i64(ld), int(frm), int(dim), _p(a), K, _p(Cm), i64(dim),
                             _p(cnt))
    return (Cm, cnt) if return_counts else Cm


def kmeans_init(X, frm, dim, K, seed=0):
    X = _f32(X)
    N, ld = X.shape
    Cm = np.zeros((K, dim), np.float32)
    rows = np.zeros(K, np.int32)
    lib().go_kmeans_init(_p(X), i64(N), i64(ld), int(frm), int(dim), K, i32(seed), _p(Cm), i64(dim),
                         _p(rows))
    return Cm, rows


def compute_clusters(X, frm, dim, K, max_iter, seed=0, tie_mode=TIE_LITERAL, nthreads=1):
    X = _f32(X)
    N, ld = X.shape
    Cm = np.zeros((K, dim), np.float32)
    nu = i32(0)
    conv = i32(0)
    rep = np.zeros((max_iter + 2, 4), np.float32)
    fa = np.zeros(N, np.int32)
    st = np.zeros(2, np.int64)
    lib().go_compute_clusters(_p(X), i64(N), i64(ld), int(frm), int(dim), K, int(max_iter),
                              i32(seed), int(tie_mode), int(nthreads), _p(Cm), i64(dim),
                              C.byref(nu), C.byref(conv), _p(rep), _p(fa), _p(st))
    return dict(centroids=Cm, updates=nu.value, converged=bool(conv.value),
                report=rep[:nu.value].copy(), assignments=fa, tie_stats=st)


def pq_train(X, M, K, max_iter, tie_mode=TIE_LITERAL, nthreads=1):
    X = _f32(X)
    N, ld = X.shape
    _, _, dmax = subvectors(ld, M)
    cb = np.zeros((M, K, dmax), np.float32)
    nu = np.zeros(M, np.int32)
    conv = np.zeros(M, np.int32)
    lib().go_pq_train(_p(X), i64(N), i64(ld), int(ld), int(M), int(K), int(max_iter), int(tie_mode),
                      int(nthreads), _p(cb), _p(nu), _p(conv))
    return cb, nu, conv


def pq_encode(X, cb, tie_mode=TIE_LITERAL, nthreads=1, stats=None):
    X = _f32(X)
    cb = _f32(cb)
    N, D = X.shape
    M, K, _ = cb.shape
    # one byte per centroid id up to K = 256 (Coder8 and narrower), 16 bits above (BytePlus coders)
    codes = np.zeros((M, N), np.uint8 if K <= 256 else np.uint16)
    st = np.zeros(2, np.int64)
    fn = lib().go_pq_encode if K <= 256 else lib().go_pq_encode16
    fn(_p(X), i64(N), i64(D), int(D), M, K, _p(cb), int(tie_mode), int(nthreads), _p(codes), _p(st))
    if stats is not None:
        stats += st
    return codes


def pq_decode(codes, cb, D):
    cb = _f32(cb)
    K = cb.shape[1]
    codes = np.ascontiguousarray(codes, np.uint8 if K <= 256 else np.uint16)
    M, N = codes.shape
    out = np.zeros((N, D), np.float32)
    fn = lib().go_pq_decode if K <= 256 else lib().go_pq_decode16
    fn(_p(codes), i64(N), i64(N), int(D), M, K, _p(cb), _p(out), i64(D))
    return out


def prepare_query(queries, cb):
    queries = _f32(queries)
    cb = _f32(cb)
    Q, D = queries.shape
    M, K, _ = cb.shape
    lut = np.zeros((Q, M, K), np.float32)
    lib().go_prepare_query(_p(queries), i64(Q), i64(D), int(D), M, K, _p(cb), _p(lut))
    return lut


def batch_query(lut, codes, k, frm=0, until=None, topk_mode=TOPK_CANONICAL, nthreads=1):
    lut = _f32(lut)
    Q, M, K = lut.shape
    codes = np.ascontiguousarray(codes, np.uint8 if K <= 256 else np.uint16)
    N = codes.shape[1]
    until = N if until is None else until
    ids = np.full((Q, max(k, 1)), -1, np.int32)
    ds = np.full((Q, max(k, 1)), np.inf, np.float32)
    sz = np.zeros(Q, np.int32)
    fn = lib().go_batch_query if K <= 256 else lib().go_batch_query16
    fn(_p(lut), i64(Q), M, K, _p(codes), i64(N), i64(frm), i64(until), int(k),
       int(topk_mode), int(nthreads), _p(ids), _p(ds), _p(sz))
    return ids[:, :k], ds[:, :k], sz


def pq_query(queries, cb, codes, k, frm=0, until=None, topk_mode=TOPK_CANONICAL, nthreads=1):
    return batch_query(prepare_query(queries, cb), codes, k, frm, until, topk_mode, nthreads)


def exact_nn(X, queries, k, frm=0, until=None, topk_mode=TOPK_CANONICAL, nthreads=1):
    X = _f32(X)
    queries = _f32(queries)
    N, D = X.shape
    Q = queries.shape[0]
    until = N if until is None else until
    ids = np.full((Q, max(k, 1)), -1, np.int32)
    ds = np.full((Q, max(k, 1)), np.inf, np.float32)
    sz = np.zeros(Q, np.int32)
    lib().go_exact_nn(_p(X), i64(D), int(D), i64(frm), i64(until), _p(queries), i64(Q), i64(D),
                      int(k), int(topk_mode), int(nthreads), _p(ids), _p(ds), _p(sz))
    return ids[:, :k], ds[:, :k], sz


def normalize(X):
    X = _f32(X)
    N, D = X.shape
    out = np.zeros_like(X)
    lib().go_normalize(_p(X), i64(N), i64(D), int(D), _p(out), i64(D))
    return out


def distance_sq(x, y):
    x = _f32(x)
    y = _f32(y)
    return float(lib().go_distance_sq(_p(x), _p(y), int(x.shape[0])))


def objective(X, frm, dim, Cm, a):
    X = _f32(X)
    Cm = _f32(Cm)
    a = np.ascontiguousarray(a, np.int32)
    return float(lib().go_objective(_p(X), i64(X.shape[0]), i64(X.shape[1]), int(frm), int(dim),
                                    _p(Cm), i64(Cm.shape[1]), _p(a)))


def grouped_query(query, centroids, offsets, strategy, limit, cb, codes, k,
                  topk_mode=TOPK_CANONICAL):
    query = _f32(query)
    centroids = _f32(centroids)
    offsets = np.ascontiguousarray(offsets, np.int32)
    cb = _f32(cb)
    codes = np.ascontiguousarray(codes, np.uint8)
    P, D = centroids.shape
    M, K, _ = cb.shape
    N = codes.shape[1]
    ids = np.full(max(k, 1), -1, np.int32)
    ds = np.full(max(k, 1), np.inf, np.float32)
    sz = i32(0)
    probed = np.zeros(P, np.int32)
    npb = i32(0)
    lib().go_grouped_query(_p(query), int(D), _p(centroids), int(P), _p(offsets), int(strategy),
                           int(limit), _p(cb), M, K, _p(codes), i64(N), i64(N), int(k),
                           int(topk_mode), _p(ids), _p(ds), C.byref(sz), _p(probed), C.byref(npb))
    return ids[:sz.value], ds[:sz.value], probed[:npb.value]


class Heap:
    """Literal TopKHeap (G/TopKHeap.scala)."""

    def __init__(self, k):
        self.k = k
        self._h = C.c_void_p(lib().go_heap_new(int(k)))

    def update(self, key, v):
        lib().go_heap_update(self._h, i32(key), C.c_float(v))

    def merge(self, other):
        lib().go_heap_merge(self._h, other._h)

    @property
    def size(self):
        return int(lib().go_heap_size(self._h))

    def raw(self):
        n = self.size
        keys = np.zeros(max(n, 1), np.int32)
        vals = np.zeros(max(n, 1), np.float32)
        lib().go_heap_raw(self._h, _p(keys), _p(vals))
        return keys[:n], vals[:n]

    def delete(self):
        r = i32(0)
        if lib().go_heap_delete(self._h, C.byref(r)) != 0:
            raise RuntimeError("heap is empty")
        return r.value

    def drain(self):
        n = self.size
        ids = np.zeros(max(n, 1), np.int32)
        ds = np.zeros(max(n, 1), np.float32)
        lib().go_heap_drain(self._h, _p(ids), _p(ds))
        return ids[:n], ds[:n]

    def __del__(self):
        try:
            lib().go_heap_free(self._h)
        except Exception:
            pass


# ---- WordVectors.Grouped / Index.GroupedIndex (host restatement over the C primitives) ----------
def grouped_build(X, assignments, coarse_centroids, keys=None):
    """WordVectors#grouped, G/WordVectors.scala:24-58, literally: two stable sorts (by word, then by
    assignment) and the run detection seeded with assignments(0) (:39) -- not with the first sorted
    row's assignment -- which yields an empty leading group unless row 0 lies in the lowest cluster.
    Returns (order, centroids [P][D], offsets [P-1])."""
    X = _f32(X)
    a = np.asarray(assignments)
    n = X.shape[0]
    idx = list(range(n))
    if keys is not None:
        idx.sort(key=lambda j: keys[j])
    idx.sort(key=lambda j: a[j])
    cents, offsets = [], []
    if n > 0:
        prev = a[0]
        cents.append(coarse_centroids[prev])
        for i, j in enumerate(idx):
            if prev != a[j]:
                offsets.append(i)
                prev = a[j]
                cents.append(coarse_centroids[prev])
    return (np.asarray(idx, np.int64), np.asarray(cents, np.float32).reshape(len(cents), X.shape[1]),
            np.asarray(offsets, np.int32))


def grouped_residuals(X, order, cents, offsets):
    """Grouped#residuals, G/WordVectors.scala:118-138."""
    G = _f32(X)[order]
    out = np.empty_like(G)
    k, nxt = -1, 0
    for i in range(G.shape[0]):
        while i >= nxt:
            k += 1
            nxt = offsets[k] if k < len(offsets) else G.shape[0]
        out[i] = G[i] - cents[k]
    return out


def grouped_query(queries, cents, offsets, n, cb, codes, k, strategy, normalized=False):
    """GroupedIndex#query, G/Index.scala:266-299 with the canonical (distance, id) heap order.
    strategy = ("groups", m) | ("vectors", n)."""
    Q = _f32(queries)
    if normalized:
        Q = normalize(Q)
    P = cents.shape[0]

    def bounds(i):
        return (0 if i == 0 else int(offsets[i - 1])), (n if i == len(offsets) else int(offsets[i]))

    ids = np.full((Q.shape[0], k), -1, np.int32)
    ds = np.full((Q.shape[0], k), np.inf, np.float32)
    sz = np.zeros(Q.shape[0], np.int32)
    for qi in range(Q.shape[0]):
        q = Q[qi:qi + 1]
        if strategy[0] == "groups":
            o, _, s = exact_nn(cents, q, min(strategy[1], P))
            nn = o[0, :s[0]]
        else:
            o, _, s = exact_nn(cents, q, P)
            order = o[0, :s[0]]
            i = cnt = 0
            while i < len(order) and cnt < strategy[1]:
                f, u = bounds(order[i])
                cnt += u - f
                i += 1
            nn = order[:i]
        cand = []
        for c in nn:
            f, u = bounds(int(c))
            if u <= f:
                continue
            r = q - cents[c]                                  # MathUtils.subtract, fp32
            ci, cd, cs = pq_query(r, cb, codes, k, f, u)
            cand += [(float(cd[0, j]), int(ci[0, j])) for j in range(cs[0])]
        cand.sort()
        cand = cand[:k]
        sz[qi] = len(cand)
        for j, (d, i) in enumerate(cand):
            ids[qi, j] = i
            ds[qi, j] = np.float32(d)
    return ids, ds, sz


# ---- CPU twin of the synthetic-data generator (oracle/synth.c; spec: gulon_b200/csrc/synth_spec.h) ----
class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("D", i32), ("centres", i32), ("latent", i32), ("nonneg", i32),
                ("noise", C.c_float), ("eps", C.c_float), ("span", C.c_float),
                ("inv_sqrt_latent", C.c_float)]


class SynthMixture:
    """Same data as gulon_b200.synth.Mixture, on host cores (identical bits)."""

    def __init__(self, D, centres=4096, noise=0.5, seed=20261018, nonneg=False, span=1.0, latent=32,
                 eps=0.05):
        L = latent if (latent and centres > 0) else 0
        inv = np.float32(1.0 / np.sqrt(np.float64(L))) if L else np.float32(0.0)
        self.p = SynthParams(int(seed), int(D), int(centres), int(L), int(bool(nonneg)), float(noise),
                             float(eps), float(span), float(inv))
        self.D = D
        W = L if L > 0 else D
        self.c = np.zeros((max(centres, 1), W), np.float32)
        self.P = np.zeros((max(L, 1), D), np.float32)
        lib().go_synth_tables(C.byref(self.p), _p(self.c), _p(self.P))

    def rows(self, lo, hi, stream_seed=0, out=None, nthreads=0):
        n = hi - lo
        if out is None:
            out = np.empty((n, self.D), np.float32)
        lib().go_synth_rows(C.byref(self.p), i64(stream_seed), i64(lo), i64(n), _p(self.c), _p(self.P),
                            _p(out), i64(out.strides[0] // 4 if n > 0 else self.D), i32(nthreads))
        return out
