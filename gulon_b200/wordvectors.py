"""WordVectors.readWord2Vec and the Unindexed / Sorted containers -- host-side mirror of
G/WordVectors.scala:60-97,143-268 (word2vec TEXT format: an optional "<rows> <dim>" header line, then
"word f f f ..." lines).  Host I/O: outside the GPU hot path, here so that an index can be built from
the same files `gulon build-index` reads.

Semantics kept from the reference
  * header detection: a first line of exactly two integers is the header (`Word2VecHeader`, :143);
    anything else is already a data line and gives the dimension (:150-156);
  * a line is split at its spaces: the text before the first space is the word, every following field
    is a float (`readFast`, :161-196); empty lines are skipped; a line with a different number of
    fields is an error (the reference throws ArrayIndexOutOfBounds / NumberFormatException);
  * `normalize = true` applies MathUtils.normalize to every row (:226) -- on the GPU here
    (gulon_normalize, bit-exact with G/MathUtils.scala:100-120);
  * progress is reported every 10 000 lines as (dimension, linesRead, linesTotal, charsPerWord) (:198-212).
Strings sort like Java's String.compareTo: by UTF-16 code unit (`utf16_key`), which differs from Python's
code-point order for supplementary-plane characters.
Decimal fields are parsed to double and rounded to float32; Java's Float.parseFloat rounds the decimal
once.  The two agree unless a field has more than 17 significant digits and sits within 2^-53 of a
float32 midpoint (word2vec writers print 6-9 digits).
"""
import io
import re
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from .vectors import Matrix, normalize

_HEADER = re.compile(r"^(\d+) (\d+)$")
CHUNK = 10000


def utf16_key(word):
    """Sort key equal to java.lang.String#compareTo (UTF-16 code units)."""
    return word.encode("utf-16-be", "surrogatepass")


@dataclass
class ProgressReport:
    """WordVectors.ProgressReport, G/WordVectors.scala:198-206."""
    dimension: int
    lines_read: int
    lines_total: Optional[int]
    chars_per_word: float

    @property
    def percentage_read(self):
        return None if self.lines_total is None else self.lines_read / self.lines_total

    @property
    def size_estimate(self):
        return int(2 * self.chars_per_word * self.lines_read) + int(4 * self.dimension * self.lines_read)


@dataclass
class Unindexed:
    """WordVectors.Unindexed(keys, toMatrix), G/WordVectors.scala:72-80."""
    keys: List[str]
    matrix: Matrix

    @property
    def size(self):
        return self.matrix.rows

    @property
    def dimension(self):
        return self.matrix.cols

    def word(self, i):
        return self.keys[i]

    def __getitem__(self, i):
        return self.matrix.data[i]

    def sorted(self):
        """WordVectors#sorted, G/WordVectors.scala:60-68: rows ordered by word (stable)."""
        order = sorted(range(self.size), key=lambda i: utf16_key(self.keys[i]))
        return Sorted([self.keys[i] for i in order], Matrix(np.ascontiguousarray(self.matrix.data[order])))

    def grouped(self, clustering, device=None):
        """WordVectors#grouped(clustering), G/WordVectors.scala:24-58 (device side: grouped.py)."""
        from .grouped import GroupedVectors
        return GroupedVectors.group(self.matrix, clustering, keys=self.keys, device=device)


@dataclass
class Sorted(Unindexed):
    """WordVectors.Sorted, G/WordVectors.scala:86-96: keys ascending."""

    def sorted(self):
        return self


def _parse_lines(lines, dimension, first_no):
    words, rows = [], np.empty((len(lines), dimension), np.float32)
    for i, line in enumerate(lines):
        parts = line.split(" ")
        if len(parts) != dimension + 1:
            raise ValueError("line %d: expected a word and %d numbers, found %d fields"
                             % (first_no + i, dimension, len(parts) - 1))
        words.append(parts[0])
        try:
            rows[i] = np.array(parts[1:], dtype=np.float64)      # correctly rounded doubles -> float32
        except ValueError as e:
            raise ValueError("line %d: %s" % (first_no + i, e)) from None
    return words, rows


def read_word2vec(source, normalize_rows=False, report: Optional[Callable[[ProgressReport], None]] = None):
    """WordVectors.readWord2Vec(reader, normalize, report), G/WordVectors.scala:214-256.
    `source`: a path, or a text file object."""
    if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__"):
        with open(source, "r", encoding="utf-8", newline="\n") as f:
            return read_word2vec(f, normalize_rows, report)
    reader = source
    first = reader.readline()
    total = None
    pending = []
    if first.endswith("\n"):
        first = first[:-1]
    m = _HEADER.match(first)
    if m:
        total, dimension = int(m.group(1)), int(m.group(2))
    else:
        dimension = len(first.split(" ")) - 1
        pending = [first]                  # already a data line (the reference pushes it back)
    if dimension < 0:
        dimension = 0
    words: List[str] = []
    blocks = []
    chars = 0
    n = 0
    eof = False
    line_no = 2 if m else 1
    while not eof:
        lines = pending
        pending = []
        consumed = len(lines)              # readFast calls that returned true, empty lines included
        while consumed < CHUNK:
            line = reader.readline()
            if line == "":
                eof = True
                break
            consumed += 1
            if line.endswith("\n"):
                line = line[:-1]
            lines.append(line)
        data = [ln for ln in lines if ln != ""]     # `if (i > 0)`: empty lines add nothing
        w, rows = _parse_lines(data, dimension, line_no)
        line_no += len(lines)
        words.extend(w)
        blocks.append(rows)
        chars += sum(len(x) for x in w)
        # the reference counts lines by the number of readFast calls that saw input, empty lines included
        n += consumed
        if report is not None:
            report(ProgressReport(dimension, n, total, chars / n if n else float("nan")))
    if report is not None:
        report(ProgressReport(dimension, n, n, chars / n if n else float("nan")))
    data = np.concatenate(blocks) if blocks else np.zeros((0, dimension), np.float32)
    if normalize_rows and len(data):
        data = normalize(data)             # MathUtils.normalize per row, on the device
    return Unindexed(words, Matrix(np.ascontiguousarray(data, np.float32)))


def write_word2vec(path_or_file, words, matrix, header=True):
    """The inverse (test fixtures, round trips): repr-exact float32 decimals."""
    data = matrix.data if isinstance(matrix, Matrix) else np.asarray(matrix, np.float32)

    def dump(f):
        if header:
            f.write("%d %d\n" % data.shape)
        for w, row in zip(words, data):
            f.write(w + " " + " ".join(np.format_float_scientific(np.float32(x), unique=True, trim="-")
                                       for x in row) + "\n")

    if hasattr(path_or_file, "write"):
        dump(path_or_file)
    else:
        with open(path_or_file, "w", encoding="utf-8", newline="\n") as f:
            dump(f)
