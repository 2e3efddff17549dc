"""The reference's on-disk index format (core/src/main/protobuf/index.proto, proto2), read and written
without generated code: Index.read / toProtobuf / fromProtobuf (G/Index.scala:147-207),
PQIndex / ProductQuantizer / EncodedMatrix mappings (G/ProductQuantizer.scala:88-106,
G/EncodedMatrix.scala:38-51).

An index file written by the reference CLI loads straight into device-resident code planes
(`repeated bytes encodings` ARE the plane-major uint8 planes the scan kernels read), and an index built
here can be written for the reference to read.  Every code width of the reference is supported: packed
planes (Coder2 / Coder4 / BytePlus, G/Coder.scala:99-168) are unpacked on load -- to one byte per id up
to 256 clusters, 16 bits above -- and packed again on write.

Two layers: `encode_index` / `decode_index` work on plain dicts of numpy arrays (no GPU needed; the
CPU tests pin them against google.protobuf with the same schema), `to_protobuf` / `from_protobuf`
map dicts to `SortedIndex` / `GroupedIndex` objects.
"""
import struct

import numpy as np

L2, COSINE = 0, 1
LIMIT_GROUPS, LIMIT_VECTORS = 0, 2


# ---- protobuf wire format ------------------------------------------------------------------------
def _varint(n):
    n &= (1 << 64) - 1          # negative int32 are written as 10-byte two's complement varints
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field, wire):
    return _varint((field << 3) | wire)


def _f_varint(field, value):
    return _tag(field, 0) + _varint(int(value))


def _f_bytes(field, payload):
    return _tag(field, 2) + _varint(len(payload)) + bytes(payload)


def _float_vector(values):
    """FloatVector { repeated float values = 1; } -- proto2 repeated scalars are NOT packed."""
    v = np.ascontiguousarray(values, "<f4")
    raw = np.empty((v.shape[0], 5), np.uint8)
    raw[:, 0] = 0x0D            # field 1, wire type 5 (32-bit)
    raw[:, 1:] = v.view(np.uint8).reshape(-1, 4)
    return raw.tobytes()


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.i = 0

    def done(self):
        return self.i >= len(self.b)

    def varint(self):
        shift = v = 0
        while True:
            if self.i >= len(self.b):
                raise ValueError("truncated varint")
            c = self.b[self.i]
            self.i += 1
            v |= (c & 0x7F) << shift
            if not c & 0x80:
                return v
            shift += 7
            if shift > 70:
                raise ValueError("varint too long")

    def field(self):
        t = self.varint()
        return t >> 3, t & 7

    def bytes_(self):
        n = self.varint()
        if self.i + n > len(self.b):
            raise ValueError("truncated length-delimited field")
        out = self.b[self.i:self.i + n]
        self.i += n
        return out

    def fixed32(self):
        out = self.b[self.i:self.i + 4]
        self.i += 4
        return out

    def skip(self, wire):
        if wire == 0:
            self.varint()
        elif wire == 1:
            self.i += 8
        elif wire == 2:
            self.bytes_()
        elif wire == 5:
            self.i += 4
        else:
            raise ValueError("unsupported wire type %d" % wire)


def _int32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _parse_float_vector(buf):
    # fast path: the unpacked layout scalapb writes (tag 0x0D + 4 bytes, repeated)
    n = len(buf)
    if n % 5 == 0:
        a = np.frombuffer(buf, np.uint8).reshape(-1, 5)
        if n == 0 or np.all(a[:, 0] == 0x0D):
            return np.ascontiguousarray(a[:, 1:]).view("<f4").reshape(-1).astype(np.float32)
    r = _Reader(buf)
    out = []
    while not r.done():
        f, w = r.field()
        if f == 1 and w == 5:
            out.append(np.frombuffer(r.fixed32(), "<f4"))
        elif f == 1 and w == 2:                                  # packed
            out.append(np.frombuffer(r.bytes_(), "<f4"))
        else:
            r.skip(w)
    return np.concatenate(out).astype(np.float32) if out else np.zeros(0, np.float32)


# ---- messages <-> dicts ----------------------------------------------------------------------------
def _enc_product_quantizer(pq):
    out = _f_varint(1, pq["num_clusters"])
    for q in pq["quantizers"]:
        body = _f_varint(1, q["start_index"]) + _f_varint(2, q["dimension"])
        body += b"".join(_f_bytes(3, _float_vector(c)) for c in q["centroids"])
        out += _f_bytes(2, body)
    return out


def _dec_product_quantizer(buf):
    r = _Reader(buf)
    pq = {"num_clusters": None, "quantizers": []}
    while not r.done():
        f, w = r.field()
        if f == 1 and w == 0:
            pq["num_clusters"] = _int32(r.varint())
        elif f == 2 and w == 2:
            rq = _Reader(r.bytes_())
            q = {"start_index": None, "dimension": None, "centroids": []}
            while not rq.done():
                f2, w2 = rq.field()
                if f2 == 1 and w2 == 0:
                    q["start_index"] = _int32(rq.varint())
                elif f2 == 2 and w2 == 0:
                    q["dimension"] = _int32(rq.varint())
                elif f2 == 3 and w2 == 2:
                    q["centroids"].append(_parse_float_vector(rq.bytes_()))
                else:
                    rq.skip(w2)
            if q["start_index"] is None or q["dimension"] is None:
                raise ValueError("Quantizer: missing required field")
            q["centroids"] = (np.stack(q["centroids"]) if q["centroids"]
                              else np.zeros((0, q["dimension"]), np.float32))
            pq["quantizers"].append(q)
        else:
            r.skip(w)
    if pq["num_clusters"] is None:
        raise ValueError("ProductQuantizer: missing num_clusters")
    return pq


def _enc_encoded_matrix(m):
    out = _f_varint(1, m["code_width"]) + _f_varint(2, m["length"])
    for plane in m["encodings"]:
        out += _f_bytes(3, np.ascontiguousarray(plane, np.uint8).tobytes())
    return out


def _dec_encoded_matrix(buf):
    r = _Reader(buf)
    m = {"code_width": None, "length": None, "encodings": []}
    while not r.done():
        f, w = r.field()
        if f == 1 and w == 0:
            m["code_width"] = _int32(r.varint())
        elif f == 2 and w == 0:
            m["length"] = _int32(r.varint())
        elif f == 3 and w == 2:
            m["encodings"].append(np.frombuffer(r.bytes_(), np.uint8).copy())
        else:
            r.skip(w)
    if m["code_width"] is None or m["length"] is None:
        raise ValueError("EncodedMatrix: missing required field")
    return m


def _enc_pq_index(v):
    return (_f_bytes(1, _enc_product_quantizer(v["product_quantizer"])) +
            _f_bytes(2, _enc_encoded_matrix(v["data"])))


def _dec_pq_index(buf):
    r = _Reader(buf)
    v = {}
    while not r.done():
        f, w = r.field()
        if f == 1 and w == 2:
            v["product_quantizer"] = _dec_product_quantizer(r.bytes_())
        elif f == 2 and w == 2:
            v["data"] = _dec_encoded_matrix(r.bytes_())
        else:
            r.skip(w)
    if "product_quantizer" not in v or "data" not in v:
        raise ValueError("PQIndex: missing required field")
    return v


def encode_index(ix):
    """dict -> bytes of message Index.  ix = {"kind": "sorted", "words": [...], "vector_index": {...},
    "metric": L2|COSINE} or {"kind": "grouped", ..., "centroids": [P][D], "offsets": [P-1],
    "strategy": LIMIT_GROUPS|LIMIT_VECTORS, "limit": n}."""
    body = b"".join(_f_bytes(1, w.encode("utf-8")) for w in ix["words"])
    body += _f_bytes(2, _enc_pq_index(ix["vector_index"])) + _f_varint(3, ix["metric"])
    if ix["kind"] == "sorted":
        return _f_bytes(1, body)
    if ix["kind"] != "grouped":
        raise ValueError("missing index implementation")
    body += b"".join(_f_bytes(4, _float_vector(c)) for c in ix["centroids"])
    body += b"".join(_f_varint(5, int(o)) for o in ix["offsets"])
    body += _f_varint(6, ix["strategy"]) + _f_varint(7, ix["limit"])
    return _f_bytes(2, body)


def decode_index(buf):
    """bytes of message Index -> dict (see encode_index); the last implementation field wins (oneof)."""
    r = _Reader(buf)
    out = None
    while not r.done():
        f, w = r.field()
        if f in (1, 2) and w == 2:
            rb = _Reader(r.bytes_())
            ix = {"kind": "sorted" if f == 1 else "grouped", "words": [], "metric": None}
            cents, offs = [], []
            while not rb.done():
                f2, w2 = rb.field()
                if f2 == 1 and w2 == 2:
                    ix["words"].append(bytes(rb.bytes_()).decode("utf-8"))
                elif f2 == 2 and w2 == 2:
                    ix["vector_index"] = _dec_pq_index(rb.bytes_())
                elif f2 == 3 and w2 == 0:
                    ix["metric"] = _int32(rb.varint())
                elif f == 2 and f2 == 4 and w2 == 2:
                    cents.append(_parse_float_vector(rb.bytes_()))
                elif f == 2 and f2 == 5 and w2 == 0:
                    offs.append(_int32(rb.varint()))
                elif f == 2 and f2 == 5 and w2 == 2:                # packed
                    rp = _Reader(rb.bytes_())
                    while not rp.done():
                        offs.append(_int32(rp.varint()))
                elif f == 2 and f2 == 6 and w2 == 0:
                    ix["strategy"] = _int32(rb.varint())
                elif f == 2 and f2 == 7 and w2 == 0:
                    ix["limit"] = _int32(rb.varint())
                else:
                    rb.skip(w2)
            if "vector_index" not in ix or ix["metric"] is None:
                raise ValueError("index: missing required field")
            if ix["metric"] not in (L2, COSINE):
                raise ValueError("unrecognized metric: %d" % ix["metric"])
            if f == 2:
                if "strategy" not in ix or "limit" not in ix:
                    raise ValueError("GroupedIndex: missing required field")
                if ix["strategy"] not in (LIMIT_GROUPS, LIMIT_VECTORS):
                    raise ValueError("strategy must be one of LIMIT_GROUPS or LIMIT_VECTORS")
                D = sum(q["dimension"] for q in ix["vector_index"]["product_quantizer"]["quantizers"])
                ix["centroids"] = np.stack(cents) if cents else np.zeros((0, D), np.float32)
                ix["offsets"] = np.asarray(offs, np.int32)
            out = ix
        else:
            r.skip(w)
    if out is None:
        raise ValueError("missing index implementation")
    return out


# ---- dicts <-> index objects (device side) -----------------------------------------------------------
class SortedIndex:
    """Index.SortedIndex(keyIndex, vectorIndex, metric), G/Index.scala:308-336: a full scan per query."""

    def __init__(self, words, vector_index, normalized=False):
        self.words = list(words)
        self.vector_index = vector_index
        self.normalized = bool(normalized)
        self._keys = None      # UTF-16 sort keys, built on the first lookup

    @property
    def size(self):
        return self.vector_index.length

    @property
    def dimension(self):
        return self.vector_index.dimension

    def batch_query(self, k, vectors):
        return self.vector_index.batch_query(k, vectors, normalize=self.normalized)

    def query(self, k, vector):
        return self.batch_query(k, np.asarray(vector, np.float32).reshape(1, -1))[0]

    def position(self, word):
        """KeyIndex.Sorted#lookup: binary search in java.lang.String#compareTo order (UTF-16 code units; it
        differs from Python's code-point order for supplementary-plane characters)."""
        import bisect
        if self._keys is None:
            self._keys = [w.encode("utf-16-be", "surrogatepass") for w in self.words]
        key = word.encode("utf-16-be", "surrogatepass")
        i = bisect.bisect_left(self._keys, key)
        return i if i < len(self._keys) and self._keys[i] == key else None

    def lookup(self, word):
        i = self.position(word)
        return None if i is None else self.vector_index.decode(i)

    # -- Index.sorted, G/Index.scala:107-113 ---------------------------------------------------------
    @staticmethod
    def build(words, matrix, quantizer, normalized=False):
        """Encode the rows of `matrix` (one per word, words ascending as in WordVectors.Sorted) with
        `quantizer` and wrap them as a full-scan index."""
        from .index import PQIndex
        from .vectors import Matrix
        m = matrix if isinstance(matrix, Matrix) else Matrix(matrix)
        words = list(words)
        if len(words) != m.rows:
            raise ValueError("one word per row expected")
        keys = [w.encode("utf-16-be", "surrogatepass") for w in words]
        if any(a > b for a, b in zip(keys[:-1], keys[1:])):
            raise ValueError("words must be sorted (KeyIndex.Sorted looks them up by binary search, in "
                             "java.lang.String#compareTo order)")
        return SortedIndex(words, PQIndex(quantizer, quantizer.encode(m)), normalized)

    def query_by_word(self, k, word):
        """Index#queryByWord, G/Index.scala:44-45: None for an unknown word."""
        v = self.lookup(word)
        return None if v is None else self.query(k, v)

    def results(self, k, vectors):
        """Index.Result per query, G/Index.scala:62-94: [(word, squared distance), ...] ascending."""
        r = self.batch_query(k, vectors)
        return [[(self.words[int(i)], float(d)) for i, d in zip(r.keys[q, :r.size[q]], r.values[q, :r.size[q]])]
                for q in range(len(r))]


def _pq_index_dict(vector_index):
    pq = vector_index.product_quantizer
    cb = pq.codebook()
    quantizers = [{"start_index": int(q.from_), "dimension": int(q.dimension),
                   "centroids": cb[m, :, :q.dimension]} for m, q in enumerate(pq.quantizers)]
    if vector_index.data is not None:
        planes = vector_index.data.codes
    else:
        planes = vector_index._keepalive[:, :vector_index.length].cpu().numpy()
    # EncodedMatrix.toProtobuf, G/EncodedMatrix.scala:38-42: the coder's width and its packed planes
    cd = pq.coder_factory(int(vector_index.length))
    packed = [cd.unwrap_code(cd.build_code(p)) for p in planes]
    return {"product_quantizer": {"num_clusters": int(pq.num_clusters), "quantizers": quantizers},
            "data": {"code_width": int(cd.width), "length": int(vector_index.length), "encodings": packed}}


def to_protobuf(index):
    """Index.toProtobuf, G/Index.scala:151-173 -> bytes."""
    from .grouped import GroupedIndex, LimitGroups
    if isinstance(index, GroupedIndex):
        gv = index.grouped
        words = gv.keys if gv.keys is not None else ["%d" % int(r) for r in gv.order]
        return encode_index({
            "kind": "grouped", "words": words, "vector_index": _pq_index_dict(index.vector_index),
            "metric": COSINE if index.normalized else L2, "centroids": gv.centroids, "offsets": gv.offsets,
            "strategy": LIMIT_GROUPS if isinstance(index.strategy, LimitGroups) else LIMIT_VECTORS,
            "limit": int(index.strategy.count)})
    if isinstance(index, SortedIndex):
        return encode_index({"kind": "sorted", "words": index.words,
                             "vector_index": _pq_index_dict(index.vector_index),
                             "metric": COSINE if index.normalized else L2})
    raise ValueError("expected a SortedIndex or a GroupedIndex")


def _pq_index_from_dict(v):
    from .index import PQIndex
    from .coder import coder as make_coder
    from .quantizer import EncodedMatrix, ProductQuantizer
    pqd, data = v["product_quantizer"], v["data"]
    cd = make_coder(data["code_width"], data["length"])          # "unsupported width" as the reference
    M, K = len(pqd["quantizers"]), pqd["num_clusters"]
    D = sum(q["dimension"] for q in pqd["quantizers"])
    dmax = max([q["dimension"] for q in pqd["quantizers"]] + [1])
    cb = np.zeros((M, K, dmax), np.float32)
    at = 0
    for m, q in enumerate(pqd["quantizers"]):
        if q["start_index"] != at:
            raise ValueError("quantizer windows must tile the dimensions in order")
        if q["centroids"].shape != (K, q["dimension"]):
            raise ValueError("quantizer %d: expected %d x %d centroids" % (m, K, q["dimension"]))
        cb[m, :, :q["dimension"]] = q["centroids"]
        at += q["dimension"]
    pq = ProductQuantizer.from_codebook(cb, D)
    want = getattr(cd, "bytes_per_code", 0) if cd.width <= 8 else cd.length + cd.lsb.bytes_per_code
    if len(data["encodings"]) != M or any(len(p) != want for p in data["encodings"]):
        raise ValueError("one code plane of %d bytes per quantizer expected" % want)
    dt = np.uint8 if K <= 256 else np.uint16
    if (cd.width > 8) != (K > 256):
        raise ValueError("code width %d does not match numClusters = %d" % (cd.width, K))
    codes = (np.stack([cd.unpack(p) for p in data["encodings"]]).astype(dt) if M
             else np.zeros((0, data["length"]), dt))
    if codes.size and int(codes.max()) >= max(K, 1):
        raise ValueError("centroid id %d out of range (numClusters = %d)" % (int(codes.max()), K))
    # the planes go straight to HBM (16-byte padded stride); the host copy stays for decode / lookup
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    stride = max(16, (data["length"] + 15) // 16 * 16)
    planes = torch.zeros((M, stride), dtype=torch.uint8 if K <= 256 else torch.uint16, device=dev)
    if data["length"]:
        planes[:, :data["length"]] = torch.from_numpy(codes).to(dev)
    ix = PQIndex.from_device_codes(pq, planes, data["length"])
    ix.data = EncodedMatrix.from_planes(cd, codes)
    return ix


def from_protobuf(buf):
    """Index.fromProtobuf, G/Index.scala:175-207: bytes -> SortedIndex | GroupedIndex."""
    from .grouped import GroupedIndex, GroupedVectors, LimitGroups, LimitVectors
    ix = decode_index(buf)
    vi = _pq_index_from_dict(ix["vector_index"])
    normalized = ix["metric"] == COSINE
    if ix["kind"] == "sorted":
        return SortedIndex(ix["words"], vi, normalized)
    gv = GroupedVectors(np.arange(vi.length, dtype=np.int64), None, ix["centroids"], ix["offsets"],
                        keys=ix["words"])
    strategy = LimitGroups(ix["limit"]) if ix["strategy"] == LIMIT_GROUPS else LimitVectors(ix["limit"])
    return GroupedIndex(gv, vi, normalized, strategy)


def write(index, path):
    with open(path, "wb") as f:
        f.write(to_protobuf(index))


def read(path):
    """Index.read, G/Index.scala:147-149."""
    with open(path, "rb") as f:
        return from_protobuf(f.read())
