/*
 * synth.c -- CPU twin of the CUDA synthetic-data generator (gulon_b200/csrc/synth.cuh): both follow
 * gulon_b200/csrc/synth_spec.h operation by operation and give identical bits.
 *
 * TEST / BENCHMARK INFRASTRUCTURE ONLY (like the rest of oracle/): it lets `bench.py --impl reference`
 * rebuild on host cores exactly the data set the GPU arm measures, and lets the parity tests
 * regenerate full-size inputs without a device.  Not a restatement of anything in the reference
 * (the reference ships no data generator for this path beyond T/Generators.scala's ScalaCheck gens).
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../gulon_b200/csrc/synth_spec.h"

/* tables: c [centres][latent or D], P [latent][D] (latent > 0) */
void go_synth_tables(const gs_params *p, float *c, float *P) {
  const int W = p->latent > 0 ? p->latent : p->D;
  for (int i = 0; i < p->centres; i++)
    for (int l = 0; l < W; l++) c[(int64_t)i * W + l] = gs_centre(p, i, l);
  for (int l = 0; l < p->latent; l++)
    for (int d = 0; d < p->D; d++) P[(int64_t)l * p->D + d] = gs_map(p, l, d);
}

/* rows [lo, lo + n) of stream `stream` -> out [n][ld] */
void go_synth_rows(const gs_params *p, int64_t stream, int64_t lo, int64_t n, const float *c,
                   const float *P, float *out, int64_t ld, int nthreads) {
  const int D = p->D, L = p->latent;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    float *z = (float *)malloc(sizeof(float) * (size_t)(L > 0 ? L : 1));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      const uint64_t row = (uint64_t)(lo + i);
      float *x = out + i * ld;
      if (p->centres <= 0) {
        for (int d = 0; d < D; d++) x[d] = gs_finish(p, gs_row_noise(p, (uint64_t)stream, row, d));
        continue;
      }
      const int idx = gs_row_centre(p, (uint64_t)stream, row);
      if (L > 0) {
        for (int l = 0; l < L; l++)
          z[l] = gs_row_latent(p, (uint64_t)stream, row, l, c[(int64_t)idx * L + l]);
        for (int d = 0; d < D; d++) x[d] = 0.0f;
        for (int l = 0; l < L; l++) {
          const float zl = z[l];
          const float *Pl = P + (int64_t)l * D;
          for (int d = 0; d < D; d++) x[d] = x[d] + zl * Pl[d]; /* -ffp-contract=off: two roundings */
        }
        for (int d = 0; d < D; d++)
          x[d] = gs_finish(p, x[d] + p->eps * gs_row_noise(p, (uint64_t)stream, row, d));
      } else {
        for (int d = 0; d < D; d++)
          x[d] = gs_finish(p, c[(int64_t)idx * D + d] + p->noise * gs_row_noise(p, (uint64_t)stream, row, d));
      }
    }
    free(z);
  }
}
