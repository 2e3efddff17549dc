"""WordVectors.readWord2Vec / sorted (G/WordVectors.scala:60-68,143-268): header detection, field
splitting, empty lines, progress reports, error cases, Java string order."""
import io

import numpy as np
import pytest

from gulon_b200.wordvectors import ProgressReport, Sorted, read_word2vec, utf16_key, write_word2vec


def test_header_and_rows():
    r = read_word2vec(io.StringIO("3 2\na 1 2\nb -0.5 1e-3\nc 3.25 4\n"))
    assert r.keys == ["a", "b", "c"] and r.dimension == 2 and r.size == 3
    assert np.array_equal(r.matrix.data, np.array([[1, 2], [-0.5, 1e-3], [3.25, 4]], np.float32))


def test_no_header_first_line_is_data_and_last_line_without_newline():
    r = read_word2vec(io.StringIO("12 7 8 9\nx 1 2 3"))
    assert r.keys == ["12", "x"] and r.dimension == 3          # "12 7 8 9" is not a header: four fields
    assert np.array_equal(r.matrix.data, np.array([[7, 8, 9], [1, 2, 3]], np.float32))
    two = read_word2vec(io.StringIO("5 6\n"))                   # exactly two integers: a header, no rows
    assert two.size == 0 and two.dimension == 6


def test_empty_lines_are_skipped_and_words_may_hold_odd_characters():
    r = read_word2vec(io.StringIO('2 1\n"quoted,word" 1\n\n\tnan 2\n'))
    assert r.keys == ['"quoted,word"', "\tnan"]
    assert np.array_equal(r.matrix.data[:, 0], np.array([1, 2], np.float32))


def test_malformed_lines_raise():
    with pytest.raises(ValueError):
        read_word2vec(io.StringIO("2 2\na 1\n"))                # too few fields
    with pytest.raises(ValueError):
        read_word2vec(io.StringIO("2 2\na 1 2 3\n"))            # too many
    with pytest.raises(ValueError):
        read_word2vec(io.StringIO("2 2\na 1 x\n"))              # NumberFormatException


def test_progress_reports_every_chunk():
    n = 25000
    text = "%d 1\n" % n + "".join("w%d %d\n" % (i, i) for i in range(n))
    seen = []
    r = read_word2vec(io.StringIO(text), report=seen.append)
    assert r.size == n
    assert [p.lines_read for p in seen] == [10000, 20000, 25000, 25000]
    assert seen[0].lines_total == n and seen[-1].lines_total == n and seen[0].dimension == 1
    assert seen[0].percentage_read == pytest.approx(0.4)
    assert isinstance(seen[-1], ProgressReport) and seen[-1].size_estimate > 0


def test_float_fields_round_like_float_parsefloat_on_short_decimals():
    rng = np.random.default_rng(0)
    x = (rng.normal(size=(50, 7)) * 10.0 ** rng.integers(-20, 20, (50, 7))).astype(np.float32)
    buf = io.StringIO()
    write_word2vec(buf, ["w%d" % i for i in range(50)], x)
    r = read_word2vec(io.StringIO(buf.getvalue()))
    assert np.array_equal(r.matrix.data.view(np.uint32), x.view(np.uint32))      # shortest round-trip decimals


def test_sorted_uses_java_string_order():
    # U+1F600 (surrogates D83D DE00) sorts BEFORE U+FF21 in UTF-16 code units, after it by code point
    words = ["Ａ", "\U0001F600", "b", "a"]
    u = read_word2vec(io.StringIO("".join("%s %d\n" % (w, i) for i, w in enumerate(words))))
    s = u.sorted()
    assert isinstance(s, Sorted)
    assert s.keys == ["a", "b", "\U0001F600", "Ａ"]
    assert sorted(words) == ["a", "b", "Ａ", "\U0001F600"]                  # Python's order differs
    assert np.array_equal(s.matrix.data[:, 0], np.array([3, 2, 1, 0], np.float32))
    assert utf16_key("\U0001F600") < utf16_key("Ａ")
    assert s.sorted() is s


def test_sorted_index_lookup_uses_java_string_order():
    """KeyIndex.Sorted is searched with String.compareTo; a Python-ordered bisect would miss
    words behind a supplementary-plane character."""
    from gulon_b200.storage import SortedIndex
    words = ["a", "b", "\U0001F600", "Ａ", "Ｂ"]           # ascending in UTF-16 code units
    ix = SortedIndex(words, None)
    assert [ix.position(w) for w in words] == [0, 1, 2, 3, 4]
    assert ix.position("zzz") is None and ix.position("") is None


@pytest.mark.gpu
def test_read_word2vec_normalizes_rows_like_mathutils(oracle):
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(1)
    x = rng.normal(size=(40, 9)).astype(np.float32)
    buf = io.StringIO()
    write_word2vec(buf, ["w%d" % i for i in range(40)], x, header=False)
    r = read_word2vec(io.StringIO(buf.getvalue()), normalize_rows=True)
    assert np.array_equal(r.matrix.data.view(np.uint32), oracle.normalize(x).view(np.uint32))
