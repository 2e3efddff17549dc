"""Coder (G/Coder.scala): the vectorised packers of gulon_b200/coder.py against the literal per-index
restatement in oracle/np_oracle.py, and the reference's own CoderSpec properties
(T/CoderSpec.scala:17-39).  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gulon_b200 import coder as C
from gulon_b200.quantizer import EncodedMatrix, coder_width
from oracle import np_oracle as O


@st.composite
def coder_with_values(draw):
    # T/CoderSpec.scala:10-14 (note Scala precedence: `1 << width - 1` is 1 << (width - 1))
    width = draw(st.integers(1, 16))
    max_value = 1 << (width - 1)
    values = draw(st.lists(st.integers(0, max_value), min_size=1, max_size=70))
    return width, values


@settings(max_examples=300, deadline=None, derandomize=True)
@given(coder_with_values())
def test_coder_round_trips_indices(cv):
    width, values = cv
    cd = C.coder(width, len(values))
    code = cd.build_code(np.array(values))
    assert [cd.get_index(code, i) for i in range(len(values))] == values
    assert list(cd.unpack(code)) == values
    # byte for byte what the reference's loops write
    assert list(code) == O.coder_build(width, len(values), values)
    assert [O.coder_get_index(width, len(values), [int(b) for b in code], i) for i in range(len(values))] == values


@settings(max_examples=100, deadline=None, derandomize=True)
@given(st.integers(1, 16), st.lists(st.integers(-2 ** 31, 2 ** 31 - 1), min_size=1, max_size=40))
def test_out_of_range_indices_are_masked_like_the_reference(width, values):
    # buildCode masks (`& 0x3`, `& 0xF`, `.toByte`, `>>> width`): no range check in the reference either
    cd = C.coder(width, len(values))
    assert list(cd.build_code(np.array(values, np.int64))) == O.coder_build(width, len(values), values)


def test_factory_for_returns_factories_for_width_1_to_16():
    for w in range(1, 17):
        f = C.factory_for(w)
        assert f is not None and f.width == O.coder_supported_width(w) and f.k == 1 << f.width
    assert C.factory_for(17) is None and C.factory_for(-1) is None
    assert C.factory_for(0).width == 0
    with pytest.raises(ValueError, match="unsupported width: 17"):
        C.coder(17, 3)


def test_produces_small_code():
    indices = np.array([1, 1, 1, 1, 1])
    for w in C.SUPPORTED_WIDTHS:
        cd = C.coder(w, len(indices))
        assert len(cd.unwrap_code(cd.build_code(indices))) == (len(indices) * w + 7) // 8


def test_coder0():
    cd = C.coder(0, 4)
    assert cd.build_code([0, 0, 0, 0]) is None and cd.get_index(None, 3) == 0
    assert len(cd.unwrap_code(None)) == 0 and list(cd.unpack(None)) == [0, 0, 0, 0]
    with pytest.raises(IndexError):
        cd.get_index(None, 4)


def test_max_width_and_cluster_limits():
    # G/ProductQuantizer.scala:11-16
    assert [C.max_width(k) for k in (1, 2, 3, 4, 5, 16, 17, 256, 257, 65536, 65537)] == \
        [0, 1, 2, 2, 3, 4, 5, 8, 9, 16, 17]
    assert [coder_width(k) for k in (1, 2, 4, 5, 16, 17, 256, 257, 1024, 1025, 4096, 4097, 65536)] == \
        [0, 2, 2, 4, 4, 8, 8, 10, 10, 12, 12, 16, 16]
    with pytest.raises(ValueError, match="too many clusters: 65537"):
        coder_width(65537)          # G/ProductQuantizer.scala:13-15


@pytest.mark.parametrize("K", [1, 3, 4, 9, 16, 200])
def test_encoded_matrix_packs_by_cluster_count(K):
    rng = np.random.default_rng(K)
    n, M = 23, 3
    planes = rng.integers(0, K, (M, n)).astype(np.uint8)
    cd = C.factory_for(C.max_width(K))(n)
    em = EncodedMatrix.from_planes(cd, planes)
    assert em.length == n and list(em(5)) == list(planes[:, 5])
    packed = em.unwrapped_encodings
    assert all(len(p) == (n * cd.width + 7) // 8 for p in packed)
    back = EncodedMatrix(cd, packed)
    assert back == em and np.array_equal(back.codes, planes)
    for m in range(M):
        want = O.coder_build(cd.width, n, [int(v) for v in planes[m]])
        assert list(packed[m]) == (want or [])
