"""The oracle against its own committed snapshot (tests/golden/oracle_snapshot.json, made by
scripts/make_golden.py).  These are NOT reference outputs -- the Scala reference cannot run here -- they
pin the oracle against itself so that a change to oracle/ that alters any bit of its answers is noticed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_its_snapshot(oracle):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import make_golden
    want = json.load(open(make_golden.OUT))
    got = make_golden.cases()
    assert sorted(got) == sorted(want)
    for name in want:
        for field in want[name]:
            assert got[name][field] == want[name][field], (name, field)
