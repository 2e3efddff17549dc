"""Synthetic inputs of the shapes BASELINE.json names, generated on the device.

Every value is a pure function of (seed, stream, row, column) -- csrc/synth_spec.h: an integer hash,
Irwin-Hall "Gaussians" and individually rounded fp32 operations -- so any rank can regenerate any row
range, the data do not depend on the GPU count, and the CPU twin (oracle/synth.c, used by bench.py's
reference arm and by the parity tests) reproduces the same bits without a device.
"""
import ctypes as C

import numpy as np

from . import _native as N


def params(D, centres=4096, noise=0.5, seed=20261018, nonneg=False, span=1.0, latent=32, eps=0.05):
    """gulon_synth_params_t for a data set (shared by the device generator and the CPU twin)."""
    L = latent if (latent and centres > 0) else 0
    inv = np.float32(1.0 / np.sqrt(np.float64(L))) if L else np.float32(0.0)
    return N.SynthParams(int(seed), int(D), int(centres), int(L), int(bool(nonneg)), float(noise),
                         float(eps), float(span), float(inv))


class Mixture:
    """Word-embedding-shaped / clustered data with a low intrinsic dimension, as real embeddings
    have: a mixture of `centres` centres in a `latent`-dimensional space (point latent = centre +
    noise * g), mapped to D dimensions by a fixed random linear map with a per-dimension scale in
    [0.5, 1.5), plus a little isotropic noise (`eps`).  latent = 0 keeps the mixture directly in D
    dimensions; centres = 0 gives plain iid unit-variance noise (config 1).  `nonneg` / `span`: |x| * span
    (SIFT-shaped, config 4)."""

    def __init__(self, D, centres=4096, noise=0.5, seed=20261018, device="cuda", nonneg=False,
                 span=1.0, latent=32, eps=0.05):
        import torch
        self.D = D
        self.device = torch.device(device)
        self.kw = dict(centres=centres, noise=noise, seed=seed, nonneg=nonneg, span=span, latent=latent,
                       eps=eps)
        self.p = params(D, **self.kw)
        W = self.p.latent if self.p.latent > 0 else D
        self.c = torch.empty((max(self.p.centres, 1), W), dtype=torch.float32, device=self.device)
        self.P = torch.empty((max(self.p.latent, 1), D), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().gulon_synth_tables_dev(C.byref(self.p), self.c.data_ptr(), self.P.data_ptr(),
                                                   torch.cuda.current_stream(self.device).cuda_stream))

    def rows(self, lo, hi, stream_seed=0, out=None):
        """float32 [hi - lo][D] CUDA tensor holding global rows [lo, hi) of stream `stream_seed`."""
        import torch
        n = hi - lo
        if out is None:
            out = torch.empty((n, self.D), dtype=torch.float32, device=self.device)
        ld = out.stride(0) if n > 1 else self.D
        with torch.cuda.device(self.device):
            N.check(N.lib().gulon_synth_rows_dev(C.byref(self.p), int(stream_seed), int(lo), int(n),
                                                 self.c.data_ptr(), self.P.data_ptr(), out.data_ptr(), ld,
                                                 torch.cuda.current_stream(self.device).cuda_stream))
        return out


def numpy_clustered(rng, n, d, centres=64, noise=0.5):
    """Small clustered test matrices (numpy RNG; unrelated to the seeded data sets above)."""
    c = rng.normal(size=(centres, d)).astype(np.float32)
    x = c[rng.integers(0, centres, n)] + noise * rng.normal(size=(n, d)).astype(np.float32)
    return np.ascontiguousarray(x, np.float32)
