"""The controller that sizes the pruned scan's lower bound (gulon_b200/csrc/mlctl.h, host logic of
scan_batch) on simulated cost curves: it finds the minimum of a convex curve from a cold start within a
few launches, parks there with exponentially rarer probes, stays at M on curves with a cliff, ignores
stale and foreign-shape measurements, and follows a workload change.  CPU only (g++)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("mlctl") / "mlctl_shim.so")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-shared", "-fPIC", "-I",
                           os.path.join(ROOT, "gulon_b200", "csrc"), "-o", so,
                           os.path.join(ROOT, "tests", "mlctl_shim.cpp")])
    L = C.CDLL(so)
    L.mlc_new.restype = C.c_void_p
    L.mlc_new.argtypes = [C.c_int]
    L.mlc_free.argtypes = [C.c_void_p]
    L.mlc_measure.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_longlong]
    for f in (L.mlc_next, L.mlc_searching, L.mlc_best, L.mlc_hold):
        f.argtypes = [C.c_void_p]
        f.restype = C.c_int
    return L


def run(lib, M, cost, launches, lag=0, shape=lambda i: 7, noise=None):
    """Simulate `launches` launches; a measurement reaches the controller `lag` launches late (callers
    enqueue ahead of the GPU).  -> list of sizes used."""
    c = lib.mlc_new(M)
    used, inflight = [], []
    try:
        for i in range(launches):
            while inflight and inflight[0][0] <= i - lag:
                _, ml, sh = inflight.pop(0)
                v = cost(ml) * (1.0 + (noise(i) if noise else 0.0))
                lib.mlc_measure(c, ml, v, sh)
            ml = lib.mlc_next(c)
            assert min(M, 4) <= ml <= M
            used.append(ml)
            inflight.append((i + 1, ml, shape(i)))
        return used, lib.mlc_best(c), lib.mlc_hold(c)
    finally:
        lib.mlc_free(c)


def convex(opt, M):
    # bound pass ~ ml, survivor evaluation doubling for every 20 % fewer quantizers: the measured shape
    return lambda ml: ml / M + 0.25 * 2.0 ** ((opt - ml) / (0.2 * M))


@pytest.mark.parametrize("M,opt", [(30, 15), (30, 24), (100, 60), (16, 9), (10, 10), (4, 4), (2, 2), (1, 1)])
def test_finds_the_minimum_and_parks(lib, M, opt):
    f = convex(opt, M)
    true_opt = min(range(min(M, 4), M + 1), key=f)
    used, best, hold = run(lib, M, f, 400)
    assert used[0] == M                                   # a cold index starts with the full bound
    assert f(best) <= f(true_opt) * 1.03                  # within 3 % of the best achievable cost
    settle = next(i for i in range(len(used)) if all(u == used[i] for u in used[i:i + 20]))
    assert settle <= 16                                   # searching ends within a dozen or so launches
    near = next(i for i, u in enumerate(used) if f(u) <= f(true_opt) * 1.10)
    assert near <= 6                                      # ... and is near the optimum after a handful
    tail = used[200:]
    assert sum(1 for u in tail if u != best) <= 6         # parked: exponentially rarer probes
    assert all(abs(u - best) <= 1 for u in tail)          # ... of a single quantizer


def test_cliff_keeps_the_full_bound(lib):
    # unclustered data: any quantizer left out multiplies the survivors
    f = lambda ml: 1.0 if ml == 10 else 5.0 * (11 - ml)
    used, best, _ = run(lib, 10, f, 300)
    assert best == 10
    assert sum(1 for u in used if u != 10) <= 8           # 24, 48, 96, ... launches between probes
    assert all(u >= 9 for u in used)


def test_stale_measurements_and_lag(lib):
    # measurements arrive 5 launches late: decisions are made only on launches that used the current size
    f = convex(15, 30)
    used, best, _ = run(lib, 30, f, 600, lag=5)
    assert f(best) <= min(f(m) for m in range(4, 31)) * 1.03
    assert used[-1] == best


def test_foreign_shapes_are_ignored_and_a_new_workload_restarts(lib):
    f = convex(15, 30)
    g = lambda ml: 10.0 * ml                              # a short last batch: costs on another scale
    shape = lambda i: 9 if i % 43 == 42 else 7            # one odd-shaped launch per step
    cost = lambda ml: f(ml)
    c_used, best, _ = run(lib, 30, cost, 400, shape=shape)
    assert f(best) <= min(f(m) for m in range(4, 31)) * 1.03
    # the workload changes for good after 200 launches: the search restarts from the current size
    f2 = convex(26, 30)
    phase = {"i": 0}

    def cost2(ml):
        return f(ml) if phase["i"] < 200 else f2(ml)

    c = lib.mlc_new(30)
    used = []
    for i in range(500):
        phase["i"] = i
        ml = lib.mlc_next(c)
        used.append(ml)
        lib.mlc_measure(c, ml, cost2(ml), 7 if i < 200 else 11)
    assert f2(lib.mlc_best(c)) <= min(f2(m) for m in range(4, 31)) * 1.03
    lib.mlc_free(c)


def test_noise_does_not_run_away(lib):
    import random
    rnd = random.Random(3)
    f = convex(15, 30)
    used, best, _ = run(lib, 30, f, 800, noise=lambda i: rnd.uniform(-0.01, 0.01))
    assert f(best) <= min(f(m) for m in range(4, 31)) * 1.06
    assert all(f(u) <= 2.5 * f(15) for u in used[60:])
