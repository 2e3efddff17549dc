"""GPU: more than 256 clusters per quantizer -- the reference's BytePlus coders (widths 10, 12, 16;
G/Coder.scala:142-168, G/ProductQuantizer.scala:11-16).  Ids are 16 bits on the device; training,
encoding, decoding, the ADC tables and the scan are compared bit for bit with the oracle."""
import numpy as np
import pytest

from test_gpu_grouped import clustered

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def codebook_from_rows(rng, X, M, K):
    from gulon_b200 import subvector_windows
    frm, dim, dmax = subvector_windows(X.shape[1], M)
    cb = np.zeros((M, K, dmax), np.float32)
    for m in range(M):
        rows = rng.integers(0, X.shape[0], K)            # with replacement: duplicates exercise the tie rule
        cb[m, :, :dim[m]] = X[rows, frm[m]:frm[m] + dim[m]]
    return cb


@pytest.mark.parametrize("K,width", [(257, 10), (1000, 10), (1025, 12), (5000, 16)])
def test_wide_encode_decode_query(g, oracle, K, width):
    rng = np.random.default_rng(K)
    n, D, M = 6000, 22, 4                                # ragged windows 6 6 5 5
    X = clustered(rng, n, D, centres=30)
    cb = codebook_from_rows(rng, X, M, K)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    assert pq.coder_factory.width == width
    enc = pq.encode(g.Matrix(X))
    want = oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST)
    assert enc.codes.dtype == np.uint16 and np.array_equal(enc.codes, want)
    assert enc.coder.width == width
    # packed planes: N most-significant bytes + the packed low bits (BytePlus)
    packed = enc.unwrapped_encodings
    assert all(len(p) == n + (n * (width - 8) + 7) // 8 for p in packed)
    assert g.EncodedMatrix(enc.coder, packed) == enc
    # decode
    dec = pq.decode(enc)
    assert np.array_equal(dec.data.view(np.uint32), oracle.pq_decode(want, cb, D).view(np.uint32))
    assert np.array_equal(pq.decode(enc(17)).view(np.uint32), dec.data[17].view(np.uint32))
    # ADC tables and the scan
    Q = np.concatenate((X[rng.integers(0, n, 5)] + 0.01, clustered(rng, 6, D, centres=30))).astype(np.float32)
    assert np.array_equal(g.prepare_query(pq, Q).view(np.uint32), oracle.prepare_query(Q, cb).view(np.uint32))
    ix = g.PQIndex(pq, enc)
    for k, frm, until in ((10, 0, n), (1, 7, n - 3), (300, 100, 4100)):
        got = ix.batch_query(k, Q, frm, until)
        wi, wd, ws = oracle.pq_query(Q, cb, want, k, frm, until)
        assert np.array_equal(got.size, ws)
        assert np.array_equal(got.keys, wi) and np.array_equal(got.values.view(np.uint32), wd.view(np.uint32))
    # device-resident encode + adopted 16-bit planes
    import torch
    dcodes = pq.encode_dev(torch.from_numpy(X).cuda())
    torch.cuda.synchronize()
    assert dcodes.dtype == torch.uint16 and np.array_equal(dcodes[:, :n].cpu().numpy(), want)
    ix2 = g.PQIndex.from_device_codes(pq, dcodes, n)
    got2 = ix2.batch_query(10, Q)
    wi, wd, ws = oracle.pq_query(Q, cb, want, 10)
    assert np.array_equal(got2.keys, wi) and np.array_equal(got2.values.view(np.uint32), wd.view(np.uint32))


def test_wide_train_and_file_round_trip(g, oracle, tmp_path):
    from gulon_b200 import storage
    rng = np.random.default_rng(9)
    n, D, M, K = 5000, 12, 3, 300
    X = clustered(rng, n, D, centres=25)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(K, M, 3))
    cb = pq.codebook()
    want_cb, _, _ = oracle.pq_train(X, M, K, 3, tie_mode=oracle.TIE_LOWEST)
    assert np.array_equal(cb.view(np.uint32), want_cb.view(np.uint32))
    enc = pq.encode(g.Matrix(X))
    assert np.array_equal(enc.codes, oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST))
    ix = storage.SortedIndex(["w%04d" % i for i in range(n)], g.PQIndex(pq, enc), False)
    path = tmp_path / "wide.index"
    storage.write(ix, path)
    d = storage.decode_index(path.read_bytes())
    assert d["vector_index"]["data"]["code_width"] == 10
    back = storage.read(path)
    assert back.vector_index.data == enc
    Q = clustered(rng, 7, D, centres=25)
    a, b = ix.batch_query(10, Q), back.batch_query(10, Q)
    assert np.array_equal(a.keys, b.keys) and np.array_equal(a.values.view(np.uint32), b.values.view(np.uint32))
    assert storage.to_protobuf(back) == path.read_bytes()


@pytest.mark.parametrize("K,M,n,k", [(1000, 4, 300_000, 10), (4096, 3, 280_000, 100), (300, 8, 270_000, 1)])
def test_wide_lower_bound_scan_matches_oracle(g, oracle, K, M, n, k):
    """Ranges of >= 262 144 rows on a wide index take the lower-bound scan (round 2): the bound pass runs
    over 8-bit GROUP ids (the K centroids of a quantizer clustered into 256 groups, a group's table entry
    = the minimum over its members), survivors are evaluated with the real 16-bit ids.  Same answer as the
    plain-table scan and as the oracle, ids and distance bits."""
    rng = np.random.default_rng(K + M)
    D = 6 * M - 1                                          # ragged windows
    X = clustered(rng, n, D, centres=60)
    cb = codebook_from_rows(rng, X, M, K)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(g.Matrix(X))
    ix = g.PQIndex(pq, enc)
    Q = np.concatenate((X[rng.integers(0, n, 9)] + 0.01, clustered(rng, 12, D, centres=60))).astype(np.float32)
    launches0 = g.kernel_launches()
    got = ix.batch_query(k, Q, 5, n - 11)                  # automatic: the pruned path
    assert g._native.counter("pscan_lb_quantizers") >= 1
    wi, wd, ws = oracle.pq_query(Q, cb, enc.codes, k, 5, n - 11)
    assert np.array_equal(got.size, ws)
    assert np.array_equal(got.keys, wi) and np.array_equal(got.values.view(np.uint32), wd.view(np.uint32))
    g.set_option("scan_impl", g.SCAN_SIMPLE)               # the plain-table scan, same index
    try:
        plain = ix.batch_query(k, Q, 5, n - 11)
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
    assert np.array_equal(plain.keys, got.keys) and np.array_equal(plain.values.view(np.uint32), got.values.view(np.uint32))
    # 16-bit fields and a pinned subset of the quantizers
    for bits, lb in ((16, 0), (8, max(1, M // 2))):
        g.set_option("pruned_bits", bits)
        g.set_option("pruned_lb_quantizers", lb)
        try:
            again = ix.batch_query(k, Q[:5], 5, n - 11)
        finally:
            g.set_option("pruned_bits", 0)
            g.set_option("pruned_lb_quantizers", 0)
        assert np.array_equal(again.keys, wi[:5]) and np.array_equal(again.values.view(np.uint32), wd[:5].view(np.uint32))
