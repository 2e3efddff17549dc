/*
 * synth_spec.h -- the synthetic data sets of SURVEY.md 8d as a PURE FUNCTION of (seed, stream, row,
 * column), in integer arithmetic plus individually rounded fp32 operations, so that the CUDA
 * generator (synth.cuh) and its CPU twin (oracle/synth.c) produce the same bits: the reference arm
 * of bench.py can rebuild, on host cores, exactly the index the GPU arm measures, and the parity
 * tests can regenerate any row range without copying it from the device.  Plain C99; no FMA
 * contraction may be applied to this file (nvcc: the _rn intrinsics below; gcc: -ffp-contract=off).
 *
 * Not part of the reference: benchmark / test tooling.
 *
 *   hash      splitmix64 finaliser chained over (seed, stream, a, b)
 *   gauss     sum of four 16-bit uniforms, centred and scaled to unit variance (Irwin-Hall, |g| < 3.47)
 *   tables    scale[d] in [0.5, 1.5); centres c[i][l]; map P[l][d] = gauss * scale[d] / sqrt(L)
 *   row r     idx = hash % centres;  z[l] = c[idx][l] + noise * gauss;
 *             x[d] = (((z0 P[0][d]) + z1 P[1][d]) + ...) + eps * gauss      (sequential fp32)
 *             latent = 0: x[d] = c[idx][d] * scale[d] + noise * gauss;  centres = 0: x[d] = gauss
 *             nonneg: x[d] = |x[d]| * span
 */
#ifndef GULON_SYNTH_SPEC_H
#define GULON_SYNTH_SPEC_H
#include <stdint.h>

#ifdef __CUDACC__
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD static inline
#endif

#define GS_T_SCALE 1
#define GS_T_CENTRE 2
#define GS_T_MAP 3
#define GS_T_ROW0 16 /* streams of row data: GS_T_ROW0 + 4 * stream + {0 idx, 1 latent noise, 2 noise} */

GS_HD uint64_t gs_mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
GS_HD uint64_t gs_key(uint64_t seed, uint64_t table, uint64_t a, uint64_t b) {
  return gs_mix(gs_mix(gs_mix(seed ^ (table << 48)) + a) + b);
}
/* exact: the integer is < 2^24, the constant is one fp32 literal, one rounded multiply */
GS_HD float gs_gauss(uint64_t h) {
  const int32_t s = (int32_t)(h & 0xffff) + (int32_t)((h >> 16) & 0xffff) + (int32_t)((h >> 32) & 0xffff) +
                    (int32_t)((h >> 48) & 0xffff) - 131070;
#ifdef __CUDA_ARCH__
  return __fmul_rn((float)s, 2.6429e-5f);
#else
  return (float)s * 2.6429e-5f;
#endif
}
GS_HD float gs_uniform(uint64_t h) { return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f); }

#ifdef __CUDA_ARCH__
#define GS_MUL(a, b) __fmul_rn((a), (b))
#define GS_ADD(a, b) __fadd_rn((a), (b))
#else
#define GS_MUL(a, b) ((a) * (b))
#define GS_ADD(a, b) ((a) + (b))
#endif

typedef struct gs_params {
  uint64_t seed;
  int32_t D, centres, latent, nonneg;
  float noise, eps, span, inv_sqrt_latent;
} gs_params;

GS_HD float gs_scale(const gs_params *p, int d) {
  return GS_ADD(gs_uniform(gs_key(p->seed, GS_T_SCALE, 0, (uint64_t)d)), 0.5f);
}
/* centre table entry: [centres][latent] (latent > 0) or [centres][D] pre-scaled (latent == 0) */
GS_HD float gs_centre(const gs_params *p, int i, int l) {
  const float g = gs_gauss(gs_key(p->seed, GS_T_CENTRE, (uint64_t)i, (uint64_t)l));
  return p->latent > 0 ? g : GS_MUL(g, gs_scale(p, l));
}
GS_HD float gs_map(const gs_params *p, int l, int d) {
  const float g = gs_gauss(gs_key(p->seed, GS_T_MAP, (uint64_t)l, (uint64_t)d));
  return GS_MUL(GS_MUL(g, gs_scale(p, d)), p->inv_sqrt_latent);
}
GS_HD int gs_row_centre(const gs_params *p, uint64_t stream, uint64_t row) {
  return (int)(gs_key(p->seed, GS_T_ROW0 + 4 * stream, row, 0) % (uint64_t)p->centres);
}
GS_HD float gs_row_latent(const gs_params *p, uint64_t stream, uint64_t row, int l, float centre_l) {
  return GS_ADD(centre_l, GS_MUL(p->noise, gs_gauss(gs_key(p->seed, GS_T_ROW0 + 4 * stream + 1, row, (uint64_t)l))));
}
GS_HD float gs_row_noise(const gs_params *p, uint64_t stream, uint64_t row, int d) {
  return gs_gauss(gs_key(p->seed, GS_T_ROW0 + 4 * stream + 2, row, (uint64_t)d));
}
GS_HD float gs_finish(const gs_params *p, float x) {
  if (p->nonneg) {
    x = x < 0.0f ? -x : x;
    x = GS_MUL(x, p->span);
  }
  return x;
}
#endif
