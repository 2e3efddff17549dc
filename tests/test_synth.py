"""The synthetic data sets (csrc/synth_spec.h): the CPU twin against a numpy restatement of the spec
(CPU), and the CUDA generator against the CPU twin (GPU) -- identical bits, any row range."""
import numpy as np
import pytest

M64 = (1 << 64) - 1


def mix(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def key(seed, table, a, b):
    return mix((mix((mix(seed ^ ((table << 48) & M64)) + a) & M64) + b) & M64)


def gauss(h):
    s = (h & 0xffff) + ((h >> 16) & 0xffff) + ((h >> 32) & 0xffff) + ((h >> 48) & 0xffff) - 131070
    return np.float32(s) * np.float32(2.6429e-5)


def np_rows(p, stream, lo, n):
    """Literal per-element restatement of synth_spec.h in numpy float32 scalars."""
    f = np.float32
    D, L = p["D"], p["latent"]
    scale = [f(f((key(p["seed"], 1, 0, d) >> 40)) * f(1.0 / 16777216.0)) + f(0.5) for d in range(D)]
    inv = f(1.0 / np.sqrt(np.float64(L))) if L else f(0)
    out = np.zeros((n, D), np.float32)
    for i in range(n):
        r = lo + i
        for d in range(D):
            nz = gauss(key(p["seed"], 16 + 4 * stream + 2, r, d))
            if p["centres"] <= 0:
                x = nz
            else:
                idx = key(p["seed"], 16 + 4 * stream, r, 0) % p["centres"]
                if L:
                    x = f(0)
                    for l in range(L):
                        z = gauss(key(p["seed"], 2, idx, l)) + f(p["noise"]) * gauss(key(p["seed"], 16 + 4 * stream + 1, r, l))
                        P = (gauss(key(p["seed"], 3, l, d)) * scale[d]) * inv
                        x = x + z * P
                    x = x + f(p["eps"]) * nz
                else:
                    x = gauss(key(p["seed"], 2, idx, d)) * scale[d] + f(p["noise"]) * nz
            if p["nonneg"]:
                x = abs(x) * f(p["span"])
            out[i, d] = x
    return out


CASES = [
    dict(D=7, centres=16, latent=5, noise=0.5, eps=0.05, nonneg=False, span=1.0, seed=20261018),
    dict(D=6, centres=8, latent=0, noise=0.25, eps=0.0, nonneg=True, span=40.0, seed=5),
    dict(D=9, centres=0, latent=0, noise=0.5, eps=0.05, nonneg=False, span=1.0, seed=77),
]


@pytest.mark.parametrize("p", CASES)
def test_cpu_twin_follows_the_spec(oracle, p):
    kw = {k: v for k, v in p.items() if k != "D"}
    mixr = oracle.SynthMixture(p["D"], **kw)
    for stream, lo, n in ((0, 0, 5), (1, 123456789, 4)):
        got = mixr.rows(lo, lo + n, stream_seed=stream)
        want = np_rows(p, stream, lo, n)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_cpu_twin_is_range_and_thread_independent(oracle):
    m = oracle.SynthMixture(20, centres=64, latent=8)
    whole = m.rows(1000, 1400, nthreads=1)
    assert np.array_equal(whole[100:250], m.rows(1100, 1250, nthreads=3))
    x = m.rows(0, 20000)
    assert abs(float(x.mean())) < 0.1 and 0.3 < float(x.std()) < 3.0


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(D=300), dict(D=128, centres=16384, nonneg=True, span=40.0),
                                dict(D=100, centres=0), dict(D=33, centres=50, latent=0)])
def test_device_generator_equals_cpu_twin(oracle, kw):
    import torch
    import gulon_b200 as g
    from gulon_b200.synth import Mixture
    if g.device_count() < 1:
        pytest.skip("no CUDA device")
    D = kw.pop("D")
    md = Mixture(D, **kw)
    mc = oracle.SynthMixture(D, **kw)
    for stream, lo, hi in ((0, 0, 1000), (1, 9_999_000, 10_000_037), (0, 65530, 65545)):
        a = md.rows(lo, hi, stream_seed=stream).cpu().numpy()
        b = mc.rows(lo, hi, stream_seed=stream)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # a strided destination (rows written into a wider matrix)
    wide = torch.zeros((50, D + 4), dtype=torch.float32, device="cuda")
    md.rows(10, 60, out=wide[:, :D])
    assert np.array_equal(wide[:, :D].cpu().numpy(), mc.rows(10, 60)) and float(wide[:, D:].abs().sum()) == 0.0
