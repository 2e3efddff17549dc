"""GPU: the tensor scan (GULON_SCAN_TENSOR, csrc/tscan.cuh) -- a tcgen05 lower bound over the decoded rows
followed by the exact re-evaluation of its survivors -- against the oracle and the exact kernels.

Three layers:
  * the filter's arithmetic: raw accumulators of one launch (gulon_debug_tscan) against a float64
    contraction of the same bf16 operands (the accumulation-error allowance eps_acc(KP) >= 2^-12 of
    tscan.cuh must hold with a wide margin), and the BOUND itself: accumulator <= d* - tau(1 + 2^-11)
    for every (row, query), d* the float64 squared distance to the reconstruction, so that every pair
    the reference could rank at or below tau has an accumulator <= 0;
  * the whole path against oracle.pq_query (ids and distance bits), on the shapes the other scan tests
    use, with ragged ranges, several stage ratios, k = 1 .. 128 and query counts that are not multiples
    of a block;
  * the hand-backs: non-finite queries, adversarial ties that overflow the survivor lists.
"""
import numpy as np
import pytest

from test_gpu_parity import build_index, check_query, clustered, random_codebook

pytestmark = pytest.mark.gpu

def eps_acc(KP):
    """tscan::eps_acc: the allowance for the tensor core's accumulation error per unit of sum |products|."""
    return 2.0 ** -12 if KP <= 320 else KP / 2.0 ** 20


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def bf16_to_f64(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def encoded_index(g, rng, n, D, M, centres=40, scale=3.0):
    X = clustered(rng, n, D, centres=centres, scale=scale)
    cb = random_codebook(rng, X, M, 256)
    pq = g.ProductQuantizer.from_codebook(cb, D)
    enc = pq.encode(X)
    return X, cb, pq, enc, g.PQIndex(pq, enc)


@pytest.fixture(params=[1, 0], ids=["pair", "single"])
def pair(request, g):
    """Both forms of the filter kernel: CTA pairs (tcgen05 cta_group::2, the default) and single CTAs."""
    g.set_option("tensor_pair", request.param)
    yield request.param
    g.set_option("tensor_pair", 1)


@pytest.mark.parametrize("n,D,M,frm,until", [
    (3000, 300, 30, 0, 3000),        # c2 shape: KP = 320, five chunks
    (2500, 128, 16, 37, 2401),       # c4 shape: KP = 144 (last chunk partly outside the operand), ragged range
    (1800, 60, 6, 0, 1800),          # KP = 64: exactly one chunk
    (2000, 100, 10, 1, 1999),        # c1 shape: KP = 112
    (1500, 37, 5, 0, 1500),          # ragged windows: KP = 48
    (1100, 316, 31, 0, 1100),        # the widest index whose query block stays in shared memory
    (900, 317, 31, 0, 900),          # one column more: the streamed-B form of the kernel (KP = 336)
    (1300, 1000, 100, 5, 1290),      # c5 shape: KP = 1008, 16 chunks, streamed B
])
def test_filter_accumulators_and_bound(g, pair, n, D, M, frm, until):
    import torch
    from gulon_b200.index import debug_tscan
    rng = np.random.default_rng(D * 7 + n)
    X, cb, pq, enc, ix = encoded_index(g, rng, n, D, M)
    nq = 200
    Q = clustered(rng, nq, D, centres=40)
    Q[0] = X[frm + 3]
    Q[1] = 0.0
    # thresholds: around each query's 20th smallest squared distance to the reconstructions
    dec = pq.decode(enc).data.astype(np.float64)
    dstar = ((Q.astype(np.float64)[:, None, :] - dec[None, frm:until, :]) ** 2).sum(axis=2)   # [nq][rows]
    taus = np.sort(dstar, axis=1)[:, 20].astype(np.float32)
    taus[2] = 0.0
    xb, qb, acc = debug_tscan(ix, torch.from_numpy(Q).cuda(), taus, frm, until)
    rows = until - frm
    KP = xb.shape[1]
    assert KP == (D + 4 + 15) // 16 * 16 and acc.shape == (rows, 256)
    A, B = bf16_to_f64(xb), bf16_to_f64(qb)
    EPS_ACC = eps_acc(KP)
    # operand rows: coordinates rounded to nearest, the norm terms on the safe side
    x32 = pq.decode(enc).data[frm:until]
    assert np.array_equal(A[:, :D], bf16_to_f64((x32.view(np.uint32) + 0x7FFF + ((x32.view(np.uint32) >> 16) & 1) >> 16).astype(np.uint16)))
    n2 = (dec[frm:until] ** 2).sum(axis=1)
    assert np.all(A[:, D:D + 2].sum(axis=1) <= n2 * (1 - EPS_ACC)) and np.all(A[:, D:D + 2].sum(axis=1) >= n2 * (1 - 2 * EPS_ACC))
    assert np.all(A[:, D + 2] >= np.sqrt(n2)) and np.all(A[:, D + 3] == 1.0) and np.all(A[:, D + 4:] == 0.0)
    assert np.all(B[nq:, :D] == 0) and np.all(B[nq:, D + 3] > 1e37)     # padding slots: nothing survives
    # 1. the tensor core's accumulation: far inside the allowance
    ref = A @ B.T
    mag = np.abs(A) @ np.abs(B).T
    err = np.abs(acc.astype(np.float64) - ref)[:, :nq]
    assert np.all(err <= (EPS_ACC / 16) * mag[:, :nq] + 1e-30), float((err / (mag[:, :nq] + 1e-300)).max())
    # 2. the bound: D <= d* - tau (1 + 2^-11) for every pair
    taup = taus.astype(np.float64) * (1 + 2.0 ** -11)
    slack = (dstar.T - taup[None, :]) - acc[:, :nq].astype(np.float64)
    assert np.all(slack >= 0), float(slack.min())
    # ... and it is tight: the gap stays within ~1 % of |q||x^| + the norms
    qn = np.sqrt((Q.astype(np.float64) ** 2).sum(axis=1))
    scale = np.sqrt(n2)[:, None] * qn[None, :] + n2[:, None] + (qn ** 2)[None, :] + taup[None, :]
    assert np.all(slack <= 0.02 * scale + 1e-30)
    # 3. every pair the reference could rank at or below tau is a survivor
    assert np.all(acc[:, :nq][dstar.T <= taup[None, :]] <= 0)
    assert np.all(acc[:, nq:] > 0)


SHAPES = [
    # n, D, M, nq, k, frm, until, ratio, boot
    (300000, 100, 10, 300, 10, 0, None, 0, 8192),       # c1 shape, two query blocks (one partial)
    (200000, 300, 30, 70, 10, 17, 199001, 4, 8192),     # c2 shape, ragged range
    (150000, 128, 16, 257, 100, 5, 149990, 0, 4096),    # c4 shape, large k (ratio 2), one query past a block
    (120000, 37, 5, 33, 1, 0, None, 16, 16),            # ragged windows, k = 1, boot shorter than a tile
    (90000, 24, 3, 5, 128, 100, 89000, 2, 4096),        # k at the limit
    (50000, 316, 31, 9, 10, 0, None, 3, 8192),          # widest resident-B index
    (120000, 1000, 100, 300, 100, 0, None, 0, 4096),    # c5 shape (streamed B), k = 100
    (60000, 400, 40, 40, 10, 7, 59000, 0, 8192),        # streamed B, KP = 416
    (20000, 64, 8, 12, 10, 0, 8192, 0, 0),              # range == boot: no stage at all
    (20000, 64, 8, 12, 10, 3, 8400, 0, 8192),           # one short stage (< 2 tiles) after the exact kernel's boot rows
    (20000, 64, 8, 12, 10, 100, 300, 0, 0),             # range shorter than the 256 boot rows: no stage at all
    (5000, 16, 2, 7, 128, 10, 160, 0, 0),               # ... with k close to the range
    (9000, 16, 2, 7, 128, 10, 400, 0, 0),               # one tiny stage after the boot rows
]


@pytest.mark.parametrize("n,D,M,nq,k,frm,until,ratio,boot", SHAPES)
def test_tensor_scan_matches_oracle(g, oracle, pair, n, D, M, nq, k, frm, until, ratio, boot):
    rng = np.random.default_rng(n + M + k)
    # (wide rows: keep the norms comparable to the neighbour distances, as centred embeddings have them --
    # the bound's slack is ~1 % of |q||x^|, and 40 far-apart clusters of 1000-d points would let a whole
    # cluster through it: that case is the hand-back test's)
    scale = 3.0 if D < 400 else 0.5
    X, cb, pq, enc, ix = encoded_index(g, rng, n, D, M, scale=scale)
    Q = clustered(rng, nq, D, centres=40, scale=scale)
    Q[0] = X[n // 2]
    until = n if until is None else until
    g.set_option("tensor_stage_ratio", ratio)
    g.set_option("tensor_boot_rows", boot)
    g.set_option("profile", 1)
    try:
        check_query(g, oracle, ix, cb, enc.codes, Q, k, frm, until, (g.SCAN_TENSOR, 0, 0))
        from gulon_b200 import _native as N
        assert N.counter("tscan_fallbacks") == 0
        if until - frm > max(boot or 8192, k):
            assert N.counter("tscan_stages") >= 1 and N.counter("tscan_tiles") > 0
            # the filter lets through little more than what the lists need
            assert N.counter("tscan_survivors") <= 64 * max(k, 8) * nq * N.counter("tscan_stages")
    finally:
        g.set_option("profile", 0)
        g.set_option("tensor_stage_ratio", 0)
        g.set_option("tensor_boot_rows", 0)


def test_tensor_scan_uniform_random_codes(g, oracle):
    # no structure at all: every row is about as far as every other, the bound prunes the least
    rng = np.random.default_rng(5)
    pq, cb, codes, ix = build_index(g, rng, 150000, 100, 10)
    Q = clustered(rng, 40, 100)
    check_query(g, oracle, ix, cb, codes, Q, 10, 0, 150000, (g.SCAN_TENSOR, 0, 0))


def test_tensor_scan_auto_selection_and_cosine(g, oracle):
    rng = np.random.default_rng(11)
    X, cb, pq, enc, ix = encoded_index(g, rng, 140000, 64, 8)
    Q = clustered(rng, 600, 64, centres=40)
    from gulon_b200 import _native as N
    g.set_option("tensor_min_rows", 100000)
    g.set_option("tensor_min_queries", 512)
    g.set_option("tensor_min_pairs", 0)
    g.set_option("profile", 1)
    try:
        got = ix.batch_query(10, Q, 0, 140000, normalize=True)
        assert N.counter("tscan_batches") == 1 and N.counter("tscan_fallbacks") == 0
        small = ix.batch_query(10, Q[:100], 0, 140000, normalize=True)      # below the query threshold: pruned scan
        assert N.counter("tscan_batches") == 1
    finally:
        g.set_option("profile", 0)
        g.set_option("tensor_min_rows", 1 << 16)
        g.set_option("tensor_min_queries", 256)
        g.set_option("tensor_min_pairs", 1 << 27)
    assert np.array_equal(small.keys, got.keys[:100]) and np.array_equal(small.values.view(np.uint32), got.values[:100].view(np.uint32))
    g.set_option("scan_impl", g.SCAN_FUSED)
    try:
        want = ix.batch_query(10, Q, 0, 140000, normalize=True)
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
    assert np.array_equal(want.keys, got.keys) and np.array_equal(want.values.view(np.uint32), got.values.view(np.uint32))


def test_tensor_scan_hands_back_what_it_cannot_bound(g, oracle):
    from gulon_b200 import _native as N
    rng = np.random.default_rng(3)
    n, D, M = 60000, 32, 4
    X, cb, pq, enc, ix = encoded_index(g, rng, n, D, M)
    Q = clustered(rng, 20, D, centres=40)
    Q[3, 5] = np.nan
    Q[7, 0] = np.inf
    Q[11, 2] = 1e38
    g.set_option("profile", 1)
    try:
        g.set_option("scan_impl", g.SCAN_TENSOR)
        got = ix.batch_query(10, Q, 0, n)
        # only the three queries go to the pruned scan, the batch itself stays
        assert N.counter("tscan_fallbacks") == 0 and N.counter("tscan_handed_back_queries") == 3
        g.set_option("scan_impl", g.SCAN_PRUNED)
        want = ix.batch_query(10, Q, 0, n)
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
        g.set_option("profile", 0)
    assert np.array_equal(want.keys, got.keys) and np.array_equal(want.values.view(np.uint32), got.values.view(np.uint32))
    ok = [q for q in range(20) if q not in (3, 7, 11)]
    check = oracle.pq_query(Q[ok], cb, enc.codes, 10, 0, n, topk_mode=oracle.TOPK_CANONICAL)
    assert np.array_equal(check[0], got.keys[ok]) and np.array_equal(check[1].view(np.uint32), got.values[ok].view(np.uint32))


def test_tensor_scan_overflowing_ties_fall_back(g, oracle):
    # three distinct codes: hundreds of thousands of rows tie with the k-th best, every one of them survives the bound
    from gulon_b200 import _native as N
    rng = np.random.default_rng(2)
    n, D, M = 400000, 16, 2
    codes = rng.integers(0, 2, (M, n)).astype(np.uint8)
    pq, cb, codes, ix = build_index(g, rng, n, D, M, codes=codes)
    Q = clustered(rng, 40, D)
    g.set_option("profile", 1)
    try:
        check_query(g, oracle, ix, cb, codes, Q, 25, 3, n - 1, (g.SCAN_TENSOR, 0, 0))
        assert N.counter("tscan_fallbacks") == 1
    finally:
        g.set_option("profile", 0)


def test_tensor_scan_refuses_what_it_cannot_serve(g):
    rng = np.random.default_rng(1)
    pq, cb, codes, ix = build_index(g, rng, 20000, 16, 2)
    Q = clustered(rng, 3, 16)
    g.set_option("scan_impl", g.SCAN_TENSOR)
    try:
        with pytest.raises(Exception):
            ix.batch_query(129, Q, 0, 20000)          # k beyond the in-kernel lists
        pq2, cb2, codes2, ix2 = build_index(g, rng, 3000, 1024, 64)
        with pytest.raises(Exception):
            ix2.batch_query(10, clustered(rng, 3, 1024), 0, 3000)   # D + 4 > 1024
    finally:
        g.set_option("scan_impl", g.SCAN_AUTO)
    r = ix.batch_query(129, Q, 0, 20000)              # automatic: served by the k-chunked scans
    assert r.keys.shape == (3, 129)
