"""The C oracle and the independent numpy restatement must agree bit for bit (tiny cases)."""
import numpy as np
import pytest

from oracle import np_oracle as npo


def _cb_list(cb, frm, dim):
    return [(int(f), cb[m, :, :d].copy()) for m, (f, d) in enumerate(zip(frm, dim))]


def _data(rng, N, D, scale=1.0):
    return (rng.standard_normal((N, D)) * scale).astype(np.float32)


@pytest.mark.parametrize("seed", range(4))
def test_assign_matches_numpy_including_ties(oracle, seed):
    rng = np.random.default_rng(seed)
    N, D, K = 300, 7, 9
    X = _data(rng, N, D)
    Cm = _data(rng, K, 3)
    Cm[4] = Cm[1]          # duplicate centroids -> exact ties -> Random(0) draws
    Cm[7] = Cm[1]
    X[10:40] = X[5]        # duplicate rows too
    frm = 2
    for literal in (True, False):
        tie = oracle.TIE_LITERAL if literal else oracle.TIE_LOWEST
        st = np.zeros(2, np.int64)
        a_c = oracle.assign(X, frm, 3, Cm, batch=0, tie_mode=tie, stats=st)
        a_n = npo.assign(X, frm, frm + 3, Cm, literal=literal)
        assert (a_c == a_n).all()
        assert st[0] > 0          # ties did occur
        a_c = oracle.assign(X, frm, 3, Cm, batch=100, tie_mode=tie, nthreads=3)
        a_n = npo.par_assign(X, frm, frm + 3, Cm, literal=literal, batch=100)
        assert (a_c == a_n).all()


def test_from_assignment_and_init_match_numpy(oracle):
    rng = np.random.default_rng(1)
    X = _data(rng, 500, 6, 3.0)
    a = rng.integers(0, 5, 500).astype(np.int32)
    a[a == 3] = 2                                  # cluster 3 empty -> all-zero centroid
    Cc = oracle.from_assignment(X, 1, 4, a, 5)
    Cn = npo.from_assignment(X, 1, 5, a, 5)
    assert Cc.tobytes() == Cn.tobytes()
    assert (Cc[3] == 0).all()
    for seed in (0, 1, 7):
        Ci, rows = oracle.kmeans_init(X, 1, 4, 6, seed)
        assert Ci.tobytes() == npo.kmeans_init(X, 1, 5, 6, seed).tobytes()
        r = npo.JavaRandom(seed)
        assert rows.tolist() == [r.next_int(500) for _ in range(6)]


@pytest.mark.parametrize("literal", [True, False])
def test_compute_clusters_matches_numpy(oracle, literal):
    rng = np.random.default_rng(2)
    cents = _data(rng, 4, 5, 4.0)
    X = (cents[rng.integers(0, 4, 400)] + _data(rng, 400, 5, 0.3)).astype(np.float32)
    tie = oracle.TIE_LITERAL if literal else oracle.TIE_LOWEST
    rc = oracle.compute_clusters(X, 1, 3, 6, 8, seed=3, tie_mode=tie, nthreads=2)
    Cn, upd, conv, fa = npo.compute_clusters(X, 1, 4, 6, 8, seed=3, literal=literal)
    assert rc["centroids"].tobytes() == Cn.tobytes()
    assert rc["updates"] == upd and rc["converged"] == conv
    assert (rc["assignments"] == fa).all()


def test_pq_encode_lut_query_match_numpy(oracle):
    rng = np.random.default_rng(3)
    N, D, M, K, Q, k = 700, 11, 4, 16, 5, 7
    X = _data(rng, N, D)
    frm, dim, dmax = oracle.subvectors(D, M)
    cb = np.zeros((M, K, dmax), np.float32)
    for m in range(M):
        cb[m, :, :dim[m]] = _data(rng, K, dim[m])
    cb[1, 5] = cb[1, 2]
    cbl = _cb_list(cb, frm, dim)
    for literal in (True, False):
        tie = oracle.TIE_LITERAL if literal else oracle.TIE_LOWEST
        codes_c = oracle.pq_encode(X, cb, tie_mode=tie, nthreads=2)
        codes_n = npo.pq_encode(X, cbl, literal=literal)
        assert (codes_c == codes_n).all()
    codes = oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST)
    qs = _data(rng, Q, D)
    lut_c = oracle.prepare_query(qs, cb)
    lut_n = npo.prepare_query(qs, cbl)
    assert lut_c.tobytes() == lut_n.tobytes()
    for literal in (True, False):
        mode = oracle.TOPK_LITERAL if literal else oracle.TOPK_CANONICAL
        ids, ds, sz = oracle.batch_query(lut_c, codes, k, topk_mode=mode, nthreads=2)
        ref = npo.batch_query(lut_n, codes, k, literal=literal)
        for q in range(Q):
            assert sz[q] == k
            assert ids[q].tolist() == ref[q][0].tolist()
            assert ds[q].tobytes() == ref[q][1].tobytes()
        ids, ds, sz = oracle.batch_query(lut_c, codes, k, 100, 130, topk_mode=mode)
        ref = npo.batch_query(lut_n, codes, k, 100, 130, literal=literal)
        for q in range(Q):
            assert ids[q, :sz[q]].tolist() == ref[q][0].tolist()


def test_scan_crosses_4096_block_boundary(oracle):
    rng = np.random.default_rng(4)
    N, M, K = 4096 * 2 + 17, 3, 8
    codes = rng.integers(0, K, (M, N)).astype(np.uint8)     # few codes -> many equal distances
    lut = rng.random((2, M, K)).astype(np.float32)
    for literal in (True, False):
        mode = oracle.TOPK_LITERAL if literal else oracle.TOPK_CANONICAL
        ids, ds, sz = oracle.batch_query(lut, codes, 20, topk_mode=mode)
        ref = npo.batch_query(lut, codes, 20, literal=literal)
        for q in range(2):
            assert ids[q].tolist() == ref[q][0].tolist()
            assert ds[q].tobytes() == ref[q][1].tobytes()


def test_exact_nn_and_normalize_match_numpy(oracle):
    rng = np.random.default_rng(5)
    X = _data(rng, 200, 9)
    X[50] = X[20]
    qs = _data(rng, 3, 9)
    for literal in (True, False):
        mode = oracle.TOPK_LITERAL if literal else oracle.TOPK_CANONICAL
        ids, ds, sz = oracle.exact_nn(X, qs, 6, topk_mode=mode)
        for q in range(3):
            ri, rd = npo.exact_nn(X, qs[q], 6, literal=literal)
            assert ids[q].tolist() == ri.tolist() and ds[q].tobytes() == rd.tobytes()
    nc = oracle.normalize(X)
    for i in (0, 17, 199):
        assert nc[i].tobytes() == npo.normalize(X[i]).tobytes()


def test_wide_ids_match_numpy(oracle):
    """More than 256 centroids: the C oracle's 16-bit twins (oracle/go_codes.inc compiled for uint16_t)
    against the numpy restatement, encode / decode / tables / scan, literal and canonical rules."""
    rng = np.random.default_rng(5)
    N, D, M, K, Q, k = 260, 7, 3, 300, 3, 6
    X = _data(rng, N, D)
    frm, dim, dmax = oracle.subvectors(D, M)
    cb = np.zeros((M, K, dmax), np.float32)
    for m in range(M):
        cb[m, :, :dim[m]] = _data(rng, K, dim[m])
    cb[2, 299] = cb[2, 7]                                 # a duplicate beyond id 255: tie rule on wide ids
    cbl = _cb_list(cb, frm, dim)
    for literal in (True, False):
        tie = oracle.TIE_LITERAL if literal else oracle.TIE_LOWEST
        codes_c = oracle.pq_encode(X, cb, tie_mode=tie, nthreads=2)
        codes_n = npo.pq_encode(X, cbl, literal=literal)
        assert codes_c.dtype == np.uint16 and (codes_c == codes_n).all()
    codes = oracle.pq_encode(X, cb, tie_mode=oracle.TIE_LOWEST)
    assert codes.max() > 255
    dec = oracle.pq_decode(codes, cb, D)
    for m in range(M):
        assert np.array_equal(dec[:, frm[m]:frm[m] + dim[m]], cb[m, codes[m], :dim[m]])
    qs = _data(rng, Q, D)
    lut_c = oracle.prepare_query(qs, cb)
    assert lut_c.tobytes() == npo.prepare_query(qs, cbl).tobytes()
    for literal in (True, False):
        mode = oracle.TOPK_LITERAL if literal else oracle.TOPK_CANONICAL
        ids, ds, sz = oracle.batch_query(lut_c, codes, k, topk_mode=mode, nthreads=2)
        ref = npo.batch_query(lut_c, codes, k, literal=literal)
        for q in range(Q):
            assert ids[q].tolist() == ref[q][0].tolist() and ds[q].tobytes() == ref[q][1].tobytes()
