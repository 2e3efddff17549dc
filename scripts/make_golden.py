"""Snapshot of the CPU oracle on seeded inputs -> tests/golden/oracle_snapshot.json.

NOT reference outputs: the reference (Scala) cannot run in this image, so these digests only pin the
oracle against ITSELF over time (a refactor of oracle/ that changes any bit of its answers fails
tests/test_oracle_golden.py).  Inputs are regenerated from the seeds below; digests are sha256 of the raw
little-endian arrays.  usage: python scripts/make_golden.py [--check]"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "oracle_snapshot.json")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cases():
    from oracle import oracle as o
    out = {}
    for name, (n, D, M, K, nq, k, seed) in {
            "c1_small": (4000, 100, 10, 256, 7, 10, 1), "c2_small": (3000, 300, 30, 256, 5, 10, 2),
            "c4_small": (5000, 128, 16, 256, 6, 100, 3), "ragged": (2500, 37, 5, 64, 4, 3, 4),
            "wide": (2000, 12, 3, 300, 5, 10, 5)}.items():
        rng = np.random.default_rng(seed)
        c = rng.normal(size=(20, D)).astype(np.float32) * 2
        X = (c[rng.integers(0, 20, n)] + 0.4 * rng.normal(size=(n, D))).astype(np.float32)
        Q = (c[rng.integers(0, 20, nq)] + 0.4 * rng.normal(size=(nq, D))).astype(np.float32)
        cb, nu, conv = o.pq_train(X, M, K, 3, tie_mode=o.TIE_LOWEST)
        cb_lit, _, _ = o.pq_train(X[:600], M, min(K, 64), 2, tie_mode=o.TIE_LITERAL)
        codes = o.pq_encode(X, cb, tie_mode=o.TIE_LOWEST)
        lut = o.prepare_query(Q, cb)
        ids, ds, sz = o.pq_query(Q, cb, codes, k)
        lids, lds, lsz = o.pq_query(Q, cb, codes, k, topk_mode=o.TOPK_LITERAL)
        ei, ed, es = o.exact_nn(X, Q, k)
        out[name] = {"shape": [n, D, M, K, nq, k, seed], "codebook": digest(cb), "updates": nu.tolist(),
                     "codebook_literal_ties": digest(cb_lit), "codes": digest(codes), "lut": digest(lut),
                     "ids": digest(ids), "dists": digest(ds), "sizes": sz.tolist(),
                     "ids_literal_heap": digest(lids), "dists_literal_heap": digest(lds),
                     "exact_ids": digest(ei), "exact_dists": digest(ed),
                     "decode": digest(o.pq_decode(codes, cb, D)), "normalize": digest(o.normalize(X[:50]))}
    return out


if __name__ == "__main__":
    got = cases()
    if "--check" in sys.argv:
        want = json.load(open(OUT))
        bad = [(k, f) for k in want for f in want[k] if got[k][f] != want[k][f]]
        print("mismatches:", bad)
        sys.exit(1 if bad else 0)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(got, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)
