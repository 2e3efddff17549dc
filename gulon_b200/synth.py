"""Synthetic inputs of the shapes BASELINE.json names, generated on the device.

Rows are produced in blocks of BLOCK rows; block b of a data set is a pure function of
(seed, b), so any rank can regenerate any row range and the data do not depend on the GPU count.
torch is used here as plumbing only (random numbers and device memory), never on the hot path.
"""
import numpy as np

BLOCK = 65536


def _gen(device, seed):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


class Mixture:
    """Word-embedding-shaped / clustered data with a low intrinsic dimension, as real embeddings
    have: a Gaussian mixture of `centres` centres in a `latent`-dimensional space (point latent =
    centre + noise * N(0, I)), mapped to D dimensions by a fixed random linear map with a
    per-dimension scale in [0.5, 1.5), plus a little isotropic noise (`eps`).  latent = 0 keeps the
    mixture directly in D dimensions; centres = 0 gives plain iid N(0, 1) (config 1)."""

    def __init__(self, D, centres=4096, noise=0.5, seed=20261018, device="cuda", nonneg=False,
                 span=1.0, latent=32, eps=0.05):
        import torch
        self.D, self.centres, self.noise, self.seed = D, centres, noise, seed
        self.device = torch.device(device)
        self.nonneg, self.span, self.eps = nonneg, span, eps
        self.L = latent if (latent and centres > 0) else 0
        if centres > 0:
            g = _gen(self.device, seed)
            scale = torch.rand((1, D), generator=g, device=self.device) + 0.5
            if self.L:
                self.c = torch.randn((centres, self.L), generator=g, device=self.device)
                self.P = torch.randn((self.L, D), generator=g, device=self.device) * scale \
                    / float(self.L) ** 0.5
            else:
                self.c = torch.randn((centres, D), generator=g, device=self.device) * scale
                self.P = None
        else:
            self.c = None
            self.P = None

    def block(self, b, stream_seed=0):
        import torch
        g = _gen(self.device, self.seed * 1000003 + 7919 * stream_seed + b + 1)
        x = torch.randn((BLOCK, self.D), generator=g, device=self.device)
        if self.c is not None:
            idx = torch.randint(0, self.centres, (BLOCK,), generator=g, device=self.device)
            if self.L:
                z = self.c[idx] + self.noise * torch.randn((BLOCK, self.L), generator=g,
                                                           device=self.device)
                # plain fp32 matmul (no TF32): data generation only, not the measured path
                x = torch.matmul(z, self.P) + self.eps * x
            else:
                x = self.c[idx] + self.noise * x
        if self.nonneg:
            x = x.abs() * self.span
        return x

    def rows(self, lo, hi, stream_seed=0, out=None):
        """float32 [hi - lo][D] CUDA tensor holding global rows [lo, hi)."""
        import torch
        n = hi - lo
        if out is None:
            out = torch.empty((n, self.D), dtype=torch.float32, device=self.device)
        b = lo // BLOCK
        at = 0
        while at < n:
            blk = self.block(b, stream_seed)
            s = lo + at - b * BLOCK
            take = min(BLOCK - s, n - at)
            out[at:at + take] = blk[s:s + take]
            at += take
            b += 1
        return out


def numpy_clustered(rng, n, d, centres=64, noise=0.5):
    """CPU twin for the reference arm (no GPU involved)."""
    c = rng.normal(size=(centres, d)).astype(np.float32)
    x = c[rng.integers(0, centres, n)] + noise * rng.normal(size=(n, d)).astype(np.float32)
    return np.ascontiguousarray(x, np.float32)
