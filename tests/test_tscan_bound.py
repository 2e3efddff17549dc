"""CPU: the algebra of the tensor scan's lower bound (csrc/tscan.cuh), restated in numpy.

The filter keeps a (row, query) pair when the contraction of two bf16 operand rows is <= 0 and the header
of tscan.cuh claims  acc <= d* - tau (1 + 2^-11)  for EVERY pair, whatever the tensor core's accumulation
order does within its allowance EPS_ACC.  This test rebuilds the operand rows exactly as decode_rows_kernel /
qprep_kernel do (round-to-nearest coordinates, norm pieces rounded to the safe side), evaluates the
contraction in float64, ADDS the whole allowance (EPS_ACC x the sum of the absolute products: the worst
case the derivation admits) and asserts the inequality pair by pair -- on centred, uncentred, tiny, huge,
zero and duplicated vectors.  The GPU test (tests/test_gpu_tscan.py) checks the same inequality on the real
accumulators and that the measured accumulation error stays 16x below the allowance.
"""
import numpy as np
import pytest

NEXTRA = 4
TAU_SLACK = 2.0 ** -11


def eps_acc(KP):
    return 2.0 ** -12 if KP <= 320 else KP / 2.0 ** 20


def bound_coef(KP):
    return 2.0 ** -7 + 2.0 ** -17 + 2.03 * eps_acc(KP)


def padded_k(D):
    return (D + NEXTRA + 15) // 16 * 16


def f32_bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def bf_rn(x):
    """float32 -> bf16 bits, round to nearest even (finite inputs)."""
    u = f32_bits(x).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def bf_rz(x):
    return (f32_bits(x) >> 16).astype(np.uint16)


def bf_rd(x):
    u = f32_bits(x)
    b = (u >> 16).astype(np.uint32)
    b = b + (((u & 0xFFFF) != 0) & ((u & 0x80000000) != 0))
    return b.astype(np.uint16)


def bf_ru(x):
    u = f32_bits(x)
    b = (u >> 16).astype(np.uint32)
    b = b + (((u & 0xFFFF) != 0) & ((u & 0x80000000) == 0))
    return b.astype(np.uint16)


def bf_val(b):
    return (np.asarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def f32_rz(v):
    """float64 -> float32 towards zero."""
    f = np.asarray(v, np.float64).astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(v)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def f32_rd(v):
    f = np.asarray(v, np.float64).astype(np.float32)
    return np.where(f.astype(np.float64) > v, np.nextafter(f, np.float32(-np.inf)), f).astype(np.float32)


def f32_ru(v):
    f = np.asarray(v, np.float64).astype(np.float32)
    return np.where(f.astype(np.float64) < v, np.nextafter(f, np.float32(np.inf)), f).astype(np.float32)


def row_operand(xhat, KP):
    """decode_rows_kernel: [bf16(x^) | a1 a2 | nb | 1 | 0..]."""
    n, D = xhat.shape
    A = np.zeros((n, KP), np.float64)
    A[:, :D] = bf_val(bf_rn(xhat))
    nrm = (xhat.astype(np.float64) ** 2).sum(axis=1)
    v = nrm * (1.0 - eps_acc(KP) - 1e-9)
    p1 = bf_val(bf_rz(f32_rz(v)))
    p2 = bf_val(bf_rd(f32_rd(v - p1)))
    A[:, D], A[:, D + 1] = p1, p2
    A[:, D + 2] = bf_val(bf_ru(f32_ru(np.sqrt(nrm) * (1.0 + 1e-9))))
    A[:, D + 3] = 1.0
    return A


def query_operand(q, tau, KP):
    """qprep_kernel: [-2 bf16(q) | 1 1 | -e | c | 0..]."""
    n, D = q.shape
    B = np.zeros((n, KP), np.float64)
    B[:, :D] = bf_val(bf_rn((-2.0 * q).astype(np.float32)))
    nrm = (q.astype(np.float64) ** 2).sum(axis=1)
    taup = tau.astype(np.float64) * (1.0 + TAU_SLACK)
    c = nrm * (1.0 - 1e-9) - taup
    B[:, D], B[:, D + 1] = 1.0, 1.0
    B[:, D + 2] = -bf_val(bf_ru(f32_ru(bound_coef(KP) * np.sqrt(nrm) * (1.0 + 1e-9))))
    B[:, D + 3] = bf_val(bf_rd(f32_rd(c - 2.0 * eps_acc(KP) * np.abs(c) - 1e-300)))
    return B


CASES = [
    # D, rows, queries, row scale, query scale, mean shift
    (300, 400, 64, 1.0, 1.0, 0.0),      # c2-like, centred
    (128, 400, 64, 40.0, 40.0, 20.0),   # c4-like: non-negative coordinates up to ~100
    (100, 300, 48, 1.0, 1.0, 0.0),
    (37, 300, 48, 3.0, 0.01, 0.0),      # queries much smaller than the rows
    (16, 300, 48, 1e-6, 1e-6, 0.0),     # tiny everything
    (64, 300, 48, 1e6, 1e6, 5e5),       # large, uncentred
    (1000, 120, 24, 1.0, 1.0, 0.3),     # c5-like: the deep contraction (KP = 1008)
    (316, 150, 24, 1.0, 2.0, 0.0),
]


@pytest.mark.parametrize("D,n,nq,rs,qs,shift", CASES)
def test_bound_holds_with_the_whole_accumulation_allowance(D, n, nq, rs, qs, shift):
    rng = np.random.default_rng(D * 31 + n)
    KP = padded_k(D)
    X = (rng.normal(size=(n, D)) * rs + shift).astype(np.float32)
    Q = (rng.normal(size=(nq, D)) * qs + shift).astype(np.float32)
    X[0] = 0.0
    Q[0] = 0.0
    Q[1] = X[5]                      # a query that coincides with a row
    X[7] = X[5]                      # a duplicated row
    dstar = ((Q.astype(np.float64)[:, None, :] - X.astype(np.float64)[None, :, :]) ** 2).sum(axis=2)   # [nq][n]
    # thresholds: each query's 10th smallest distance (as fp32), one at 0, one far above everything
    tau = np.sort(dstar, axis=1)[:, 9].astype(np.float32)
    tau[2] = 0.0
    tau[3] = np.float32(dstar.max() * 4 + 1)
    A, B = row_operand(X, KP), query_operand(Q, tau, KP)
    acc = A @ B.T                                             # [n][nq], exact
    worst = acc + eps_acc(KP) * (np.abs(A) @ np.abs(B).T)     # every rounding of the accumulation against us
    taup = tau.astype(np.float64) * (1.0 + TAU_SLACK)
    assert np.all(worst <= dstar.T - taup[None, :] + 1e-300)
    # hence: every pair the reference could rank at or below tau survives (acc <= 0), even in the worst case
    assert np.all(worst[dstar.T <= taup[None, :]] <= 0)
    # and the bound is not vacuous: far rows are rejected (exact accumulation), the slack stays ~1 % of the norms
    qn, xn = np.sqrt((Q.astype(np.float64) ** 2).sum(1)), np.sqrt((X.astype(np.float64) ** 2).sum(1))
    slack = (dstar.T - taup[None, :]) - acc
    scale = xn[:, None] * qn[None, :] + (xn ** 2)[:, None] + (qn ** 2)[None, :] + taup[None, :]
    assert np.all(slack <= 0.02 * scale + 1e-300)
    assert np.all(acc[:, 3] <= 0)                             # the huge threshold keeps every row


def test_blocking_row_lets_nothing_through():
    """qprep's row for a query it cannot serve (and for padding slots): c = 2^126 against the rows' 1.0."""
    D, KP = 100, padded_k(100)
    rng = np.random.default_rng(1)
    A = row_operand((rng.normal(size=(50, D)) * 1e3).astype(np.float32), KP)
    b = np.zeros(KP)
    b[D + 3] = bf_val(np.uint16(0x7E80))
    assert b[D + 3] == 2.0 ** 126
    assert np.all(A @ b > 0)
