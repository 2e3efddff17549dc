"""Known-answer tests that pin the oracle's integer plumbing.

The reference's own tests hold no golden vectors for this path, and the JVM is
absent, so the only external anchors are the JDK-documented java.util.Random outputs and
hand-derived TopKHeap tie traces (G/TopKHeap.scala).  PARITY UNPINNED beyond these.
"""
import numpy as np

from oracle import np_oracle as npo


def test_java_random_next_int_seed0(oracle):
    # the well-known JDK values for new Random(0).nextInt()
    for R in (oracle.JRandom, npo.JavaRandom):
        r = R(0)
        assert [r.next_int() for _ in range(3)] == [-1155484576, -723955400, 1033096058]


def test_java_random_next_boolean_seed0(oracle):
    want = [int(c) for c in "1101101011000111"]
    for R in (oracle.JRandom, npo.JavaRandom):
        r = R(0)
        assert [int(r.next_boolean()) for _ in range(16)] == want


def test_java_random_bounded(oracle):
    kat = {
        (0, 1000000): [741360, 505948, 548029, 116447, 843515],
        (1, 1000000): [548985, 764588, 641847, 970313, 64254],
        (2, 1000000): [126108, 21372, 844040, 925067, 918389],
        (0, 10000): [1360, 5948, 8029, 6447, 3515],
    }
    for (seed, bound), want in kat.items():
        for R in (oracle.JRandom, npo.JavaRandom):
            r = R(seed)
            assert [r.next_int(bound) for _ in range(5)] == want


def test_java_random_power_of_two_and_agreement(oracle):
    a, b = oracle.JRandom(12345), npo.JavaRandom(12345)
    for bound in (1, 2, 256, 65536, 3, 7, 1000, 2**31 - 1, 2**30 + 1):
        for _ in range(50):
            assert a.next_int(bound) == b.next_int(bound)
    a, b = oracle.JRandom(-7), npo.JavaRandom(-7)
    assert [a.next_float() for _ in range(20)] == [float(b.next_float()) for _ in range(20)]


def test_subvectors_split_rule(oracle):
    # G/Vectors.scala:84-104; T/VectorsSpec.scala:42-65: n parts, sizes differ by <= 1, contiguous
    for D in range(1, 40):
        for M in range(1, D + 1):
            frm, dim, dmax = oracle.subvectors(D, M)
            assert dmax == -(-D // M)
            assert frm[0] == 0 and (frm[1:] == (frm + dim)[:-1]).all() and frm[-1] + dim[-1] == D
            assert dim.max() - dim.min() <= 1 and (np.diff(dim) <= 0).all()
            assert [(f, f + d) for f, d in zip(frm, dim)] == npo.subvectors(D, M)
    frm, dim, _ = oracle.subvectors(300, 30)
    assert (dim == 10).all()
    frm, dim, _ = oracle.subvectors(10, 4)   # ideal 3, shortfall 2 -> 3,3,2,2
    assert dim.tolist() == [3, 3, 2, 2] and frm.tolist() == [0, 3, 6, 8]


def test_heap_tie_traces(oracle):
    # hand-traced on G/TopKHeap.scala: k=2, (0,5),(1,5) drains to ids [1,0]
    for H in (oracle.Heap, npo.TopKHeap):
        h = H(2)
        h.update(0, 5.0)
        h.update(1, 5.0)
        ids, ds = h.drain()
        assert list(ids) == [1, 0] and list(ds) == [5.0, 5.0]
        h = H(2)
        h.update(0, 5.0)
        h.update(1, 5.0)
        h.update(2, 3.0)       # evicts the root (id 0), keeps id 1
        ids, ds = h.drain()
        assert list(ids) == [2, 1] and list(ds) == [3.0, 5.0]
        h = H(2)
        h.update(0, 5.0)
        h.update(1, 5.0)
        h.update(2, 5.0)       # equal to the root: strict '>' rejects
        ids, _ = h.drain()
        assert sorted(ids) == [0, 1]


def test_heap_empty_delete(oracle):
    import pytest
    with pytest.raises(RuntimeError):
        oracle.Heap(3).delete()
    with pytest.raises(RuntimeError):
        npo.TopKHeap(3).delete()
