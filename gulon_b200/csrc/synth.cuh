// synth.cuh -- the synthetic data sets of SURVEY.md 8d on the device: a pure function of (seed, stream,
// row, column) specified in synth_spec.h, whose CPU twin is oracle/synth.c (identical bits).
// Benchmark / test tooling, not part of the reference path.
#pragma once
#include "common.cuh"
#include "synth_spec.h"

namespace gulon {
namespace synth {

__global__ void tables_kernel(const gs_params p, float *__restrict__ c, float *__restrict__ P) {
  const int W = p.latent > 0 ? p.latent : p.D;
  const i64 nc = (i64)p.centres * W, np = (i64)p.latent * p.D;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < nc + np; t += (i64)gridDim.x * blockDim.x) {
    if (t < nc)
      c[t] = gs_centre(&p, (int)(t / W), (int)(t % W));
    else
      P[t - nc] = gs_map(&p, (int)((t - nc) / p.D), (int)((t - nc) % p.D));
  }
}

constexpr int ROWS = 8;     // rows per block pass
constexpr int LMAX = 64;    // latent dimensions held in shared memory
// grid-stride over groups of ROWS rows; block 256.  Latent coordinates of the group in shared memory,
// then thread = (row, column) walks the L products in order (two roundings per term).
__global__ void __launch_bounds__(256) rows_kernel(const gs_params p, i64 stream, i64 lo, i64 n,
                                                   const float *__restrict__ c,
                                                   const float *__restrict__ P, float *__restrict__ out,
                                                   i64 ld) {
  __shared__ float z[ROWS][LMAX];
  __shared__ int cidx[ROWS];
  const int tid = threadIdx.x, D = p.D, L = p.latent;
  for (i64 g0 = (i64)blockIdx.x * ROWS; g0 < n; g0 += (i64)gridDim.x * ROWS) {
    const int nr = (int)(n - g0 < ROWS ? n - g0 : ROWS);
    if (p.centres > 0) {
      if (tid < nr) cidx[tid] = gs_row_centre(&p, (u64)stream, (u64)(lo + g0 + tid));
      __syncthreads();
      for (int t = tid; t < nr * L; t += 256) {
        const int r = t / L, l = t % L;
        z[r][l] = gs_row_latent(&p, (u64)stream, (u64)(lo + g0 + r), l, c[(i64)cidx[r] * L + l]);
      }
      __syncthreads();
    }
    for (int t = tid; t < nr * D; t += 256) {
      const int r = t / D, d = t % D;
      const u64 row = (u64)(lo + g0 + r);
      float x;
      if (p.centres <= 0) {
        x = gs_row_noise(&p, (u64)stream, row, d);
      } else if (L > 0) {
        x = 0.0f;
        for (int l = 0; l < L; l++) x = __fadd_rn(x, __fmul_rn(z[r][l], P[(i64)l * D + d]));
        x = __fadd_rn(x, __fmul_rn(p.eps, gs_row_noise(&p, (u64)stream, row, d)));
      } else {
        x = __fadd_rn(c[(i64)cidx[r] * D + d], __fmul_rn(p.noise, gs_row_noise(&p, (u64)stream, row, d)));
      }
      out[(g0 + r) * ld + d] = gs_finish(&p, x);
    }
    __syncthreads();
  }
}

}  // namespace synth
}  // namespace gulon
