"""GPU: the recall harness (G/Tests.scala:11-121) against a host restatement built on the oracle's
exact kNN and ADC query: same sampled rows, same k-th true distances, same per-k recall statistics."""
import numpy as np
import pytest

from test_gpu_grouped import build, clustered

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gulon_b200 as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: gulon_b200 has no CPU fallback")
    return g


def host_recall(oracle, X, Q, results, ids, eps):
    """Tests#recallOf, G/Tests.scala:18-41, on the host: exact distance of each returned id in returned
    order (MathUtils.distanceSq, sequential fp32), count within the cut-off."""
    per_k = {}
    for i in range(Q.shape[0]):
        row_ids = [int(v) for v in ids[i] if v >= 0]
        dist = np.array([oracle.distance_sq(Q[i], X[r]) for r in row_ids], np.float32)
        for k, max_d in results[i]:
            cutoff = np.float32(max_d) if eps == 0 else \
                np.float32(np.float64(np.sqrt(np.float64(np.float32(max_d))) * np.float64(np.float32(1) + np.float32(eps))) ** 2)
            per_k.setdefault(k, []).append(np.float32(np.count_nonzero(dist[:k] <= cutoff)) / np.float32(k))
    return per_k


@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_recall_of_pq_index_matches_host(g, oracle, eps):
    rng = np.random.default_rng(21)
    n, D, M = 4000, 24, 4
    X = clustered(rng, n, D)
    pq = g.ProductQuantizer.train(g.Matrix(X), g.ProductQuantizerConfig(256, M, 3))
    ix = g.PQIndex(pq, pq.encode(g.Matrix(X)))
    ks = (1, 2, 3, 5, 10, 25)
    t = g.Tests.sample(g.Matrix(X), lambda i: X[i], sample_size=40, ks=ks, seed=0)
    # Tests.sample: rows drawn by java.util.Random(0).nextInt(n)
    jr = oracle.JRandom(0)
    assert t.sampled_rows == [jr.next_int(n) for _ in range(40)]
    Q = X[t.sampled_rows]
    gi, gd, gs = oracle.exact_nn(X, Q, max(ks))
    for i, qq in enumerate(t.queries):
        assert [k for k, _ in qq.results] == list(ks)
        assert np.array_equal(np.array([d for _, d in qq.results], np.float32).view(np.uint32),
                              gd[i, [k - 1 for k in ks]].view(np.uint32))
    got = t.recall_of(ix, eps=eps)
    wi, wd, ws = oracle.pq_query(Q, pq.codebook(), ix.data.codes, max(ks))
    want = host_recall(oracle, X, Q, [qq.results for qq in t.queries], wi, eps)
    assert sorted(got) == sorted(want)
    for k in ks:
        w = g.SummaryStats()
        for v in want[k]:
            w = w + g.SummaryStats(1, float(v), 0.0)
        assert got[k] == w and 0.0 <= got[k].mean <= 1.0
    assert got[1].mean > 0.2        # a database row finds itself or a code twin


def test_recall_of_grouped_index_and_ks_beyond_the_data(g, oracle):
    rng = np.random.default_rng(22)
    X, coarse, ks_, gv, res, pq, ix = build(g, oracle, rng, n=3000, D=16, P=6, M=4, K=64)
    ix.strategy = g.LimitGroups(6)                  # probe everything: recall is the PQ recall
    Q = clustered(rng, 15, 16, centres=6)
    t = g.Tests.for_queries(g.Matrix(X), Q, ks=(1, 10, 5000))
    assert all([k for k, _ in qq.results] == [1, 10] for qq in t.queries)   # ks.filter(_ <= result.length)
    got = t.recall_of(ix)
    keys = ix.batch_query(10, Q).keys
    want = host_recall(oracle, X, Q, [qq.results for qq in t.queries], ix.original_rows(keys), 0.0)
    for k in (1, 10):
        assert abs(got[k].mean - float(np.mean(np.asarray(want[k], np.float64)))) < 1e-6
        assert got[k].count == 15
