// kupdate.cuh -- the shardable centroid update (GULON_UPDATE_SUM) as an exact segmented sum.
//
// Reference: KMeans.fromAssignment, G/KMeans.scala:198-226 computes per-cluster running means in row
// order; that recurrence cannot be sharded.  The sum/count form can: every rank adds up
// its rows per cluster, the sums and counts are all-reduced, every rank divides.  Floating-point
// sums depend on the order of addition, i.e. on the CTA schedule and on the number of GPUs.  This
// kernel therefore accumulates in FIXED POINT:
//
//   v = round(x * 2^S) + 2^28,   2^S chosen per window so that |x| * 2^S < 2^28 for every row
//
// (one pass over the matrix finds the per-window max |x| when training starts).  v is a positive
// 30-bit integer; the per-(cluster, dimension) sums are 64-bit integers kept in shared memory as two
// 32-bit words updated with NATIVE shared-memory atomics (ATOMS.ADD; a carry out of the low word is
// detected from the returned old value and added to the high word -- exact under any interleaving,
// and needed for ~1 value in 16).  Integer addition is associative, so the result is bit-identical
// whatever the order: run to run, CTA schedule to CTA schedule, and for ANY number of GPUs once the
// partial sums are all-reduced as int64.  The centroid is fl32((sum - n 2^28) 2^-S / n) in double.
// Precision: every value is kept to 2^-29 of the window's largest magnitude (fp32 keeps 2^-24 of
// each value's own magnitude); the sum itself is exact.
//
// Data path (persistent CTAs, unit = (group of <= 3 adjacent windows, row range)): one thread issues
// TMA boxes X[256 rows][32 floats] (128-byte swizzle) into a ring of raw tiles, 8 warps (thread =
// row) read their row's windows and the row's assignments and issue the atomics; at the end of the
// unit the shared accumulators are flushed with 64-bit global atomics (RED.ADD.64).
#pragma once
#include "tcassign.cuh"

namespace gulon {
namespace upd {

constexpr int ROWS = 256;                      // rows per raw tile
constexpr int NRAW = 3;
constexpr int RAW_BYTES = ROWS * tca::BOX_COLS * 4;  // 32768
constexpr int NT = 32 * 9;                     // 8 row warps + the TMA warp
constexpr int KMAX = 256;
constexpr int FIX_BITS = 28;
constexpr int GRP_MAX = tca::GRP_MAX;
constexpr int BAR_BYTES = 64;

__host__ __device__ constexpr int acc_words(int dim) { return GRP_MAX * (2 * KMAX * dim + 2 * KMAX); }
__host__ __device__ constexpr int smem_bytes(int dim) { return NRAW * RAW_BYTES + acc_words(dim) * 4 + BAR_BYTES; }

struct Params {
  i64 N;
  int unit_rows;                 // multiple of ROWS
  const int32_t *groups;         // [n_groups][GRP_MAX] window ids (tcassign.cuh)
  const int32_t *from;           // [M]
  int n_groups, K, dmax;
  const int32_t *assign;         // [M][astride]
  i64 astride;
  const float *scale;            // [M] 2^S
  unsigned long long *sums;      // [M][K][dmax] biased fixed-point sums
  int32_t *counts;               // [M][K]
  int32_t *bad;                  // [M][K] != 0: a non-finite coordinate was assigned to the cluster
};

// ---- per-window max |x| over the finite values ---------------------------------------------------
// grid-stride over rows, one warp per row, lanes over columns; col2win[c] = window of column c or -1.
__global__ void absmax_kernel(const float *__restrict__ X, i64 N, int ncols, i64 ld,
                              const int32_t *__restrict__ col2win, int n_windows,
                              unsigned int *__restrict__ absmax_bits) {
  extern __shared__ unsigned int s_max[];  // [ncols]
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) s_max[c] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c0 = 0; c0 < ncols; c0 += 32 * 8) {
    unsigned int m[8];
#pragma unroll
    for (int u = 0; u < 8; u++) m[u] = 0;
    for (i64 r = (i64)blockIdx.x * nwarp + warp; r < N; r += (i64)gridDim.x * nwarp) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int c = c0 + u * 32 + lane;
        if (c < ncols) {
          const unsigned int b = __float_as_uint(X[r * ld + c]) & 0x7fffffffu;
          if (b < 0x7f800000u && b > m[u]) m[u] = b;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int c = c0 + u * 32 + lane;
      if (c < ncols && m[u]) atomicMax(&s_max[c], m[u]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    const int w = col2win[c];
    if (w >= 0 && w < n_windows && s_max[c]) atomicMax(&absmax_bits[w], s_max[c]);
  }
}

// scale[w] = 2^(FIX_BITS - e) with 2^e > absmax >= 2^(e-1)   (exponent clamped to +-100)
__global__ void scale_kernel(const unsigned int *__restrict__ absmax_bits, int n, float *__restrict__ scale) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const float a = __uint_as_float(absmax_bits[w]);
  int e = a > 0.0f ? ilogbf(a) + 1 : 0;
  int s = FIX_BITS - e;
  s = s < -100 ? -100 : (s > 100 ? 100 : s);
  scale[w] = exp2f((float)s);
}

// ---- the accumulation kernel ------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(NT, 1) update_fixed_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using namespace tca;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *raw_s = smem;
  unsigned int *acc = reinterpret_cast<unsigned int *>(smem + NRAW * RAW_BYTES);
  // per window slot w: LO[K][DIM], HI[K][DIM], CNT[K], BAD[K]
  constexpr int WSTRIDE = 2 * KMAX * DIM + 2 * KMAX;
  uint64_t *bars = reinterpret_cast<uint64_t *>(acc + GRP_MAX * WSTRIDE);
  uint64_t *raw_full = bars, *raw_empty = bars + NRAW;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const i64 n_ranges = (p.N + p.unit_rows - 1) / p.unit_rows;
  const i64 n_units = n_ranges * p.n_groups;
  if (tid == 0) {
    for (int i = 0; i < NRAW; i++) {
      mb_init(raw_full + i, 1);
      mb_init(raw_empty + i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = tid; t < GRP_MAX * WSTRIDE; t += NT) acc[t] = 0;
  __syncthreads();

  i64 blk = 0;
  for (i64 unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int g = (int)(unit % p.n_groups);
    const i64 r_begin = (unit / p.n_groups) * p.unit_rows;
    const i64 r_end = r_begin + p.unit_rows < p.N ? r_begin + p.unit_rows : p.N;
    const int32_t *gw = p.groups + g * GRP_MAX;
    const int nw = 1 + (gw[1] >= 0 ? 1 : 0) + (gw[2] >= 0 ? 1 : 0);
    const int n_blocks = (int)((r_end - r_begin + ROWS - 1) / ROWS);
    const int from0 = p.from[gw[0]];
    const int col0 = from0 & ~3, m0 = from0 & 3;

    if (warp == 8) {
      if (lane == 0) {
        for (int b = 0; b < n_blocks; b++) {
          const i64 bi = blk + b;
          const int rb = (int)(bi % NRAW);
          mb_wait(raw_empty + rb, (uint32_t)(((bi / NRAW) & 1) ^ 1));
          mb_expect_tx(raw_full + rb, RAW_BYTES);
          tma_box(raw_s + rb * RAW_BYTES, &tmap, col0, (int)(r_begin + (i64)b * ROWS), raw_full + rb);
        }
      }
      __syncwarp();
    } else {
      const unsigned char *row_off = raw_s + tid * 128;
      const int sw = tid & 7;
      float scl[GRP_MAX];
#pragma unroll
      for (int w = 0; w < GRP_MAX; w++) scl[w] = w < nw ? p.scale[gw[w]] : 0.0f;
      for (int b = 0; b < n_blocks; b++) {
        const i64 bi = blk + b;
        const int rb = (int)(bi % NRAW);
        const i64 row = r_begin + (i64)b * ROWS + tid;
        const bool live = row < r_end;
        int a[GRP_MAX];
#pragma unroll
        for (int w = 0; w < GRP_MAX; w++) a[w] = (live && w < nw) ? p.assign[(i64)gw[w] * p.astride + row] : 0;
        mb_wait(raw_full + rb, (uint32_t)((bi / NRAW) & 1));
        if (live) {
          for (int w = 0; w < nw; w++) {
            float x[DIM];
            load_window<DIM>(row_off + rb * RAW_BYTES, sw, m0 + w * DIM, x);
            unsigned int *lo = acc + w * WSTRIDE + a[w] * DIM;
            unsigned int *hi = lo + KMAX * DIM;
            // all low-word atomics first (back to back, their returns in flight together), carries after
            bool bad = false;
            unsigned int v[DIM], old[DIM];
#pragma unroll
            for (int j = 0; j < DIM; j++) {
              const float sx = x[j] * scl[w];            // exact: a power of two
              const bool ok = fabsf(sx) <= 268435456.0f;  // finite and within 2^28 (always, for finite x)
              bad = bad || !ok;
              v[j] = ok ? (unsigned int)(__float2int_rn(sx) + (1 << FIX_BITS)) : 0u;
            }
#pragma unroll
            for (int j = 0; j < DIM; j++) old[j] = atomicAdd(lo + j, v[j]);
#pragma unroll
            for (int j = 0; j < DIM; j++)
              if (old[j] + v[j] < old[j]) atomicAdd(hi + j, 1u);  // carry out of the low word
            atomicAdd(acc + w * WSTRIDE + 2 * KMAX * DIM + a[w], 1u);
            if (bad) atomicOr(acc + w * WSTRIDE + 2 * KMAX * DIM + KMAX + a[w], 1u);
          }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(raw_empty + rb);
      }
      // flush the unit's sums (and clear them for the next unit)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int w = 0; w < nw; w++) {
        unsigned int *lo = acc + w * WSTRIDE, *hi = lo + KMAX * DIM, *cnt = hi + KMAX * DIM, *bd = cnt + KMAX;
        const int m = gw[w];
        for (int t = tid; t < p.K * DIM; t += 256) {
          const unsigned long long v = ((unsigned long long)hi[t] << 32) | lo[t];
          if (v) {
            atomicAdd(p.sums + ((i64)m * p.K + t / DIM) * p.dmax + t % DIM, v);
            lo[t] = 0;
            hi[t] = 0;
          }
        }
        for (int t = tid; t < p.K; t += 256) {
          if (cnt[t]) {
            atomicAdd(p.counts + (i64)m * p.K + t, (int)cnt[t]);
            cnt[t] = 0;
          }
          if (bd[t]) {
            atomicOr(p.bad + (i64)m * p.K + t, 1);
            bd[t] = 0;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    blk += n_blocks;
  }
}

// centroid = (sum - n 2^28) 2^-S / n in double, rounded once to fp32; empty cluster -> zeros (as the
// reference leaves it); a cluster that received a non-finite coordinate -> NaN.  grid covers M*K*dmax
__global__ void finalize_fixed_kernel(const unsigned long long *__restrict__ sums,
                                      const int32_t *__restrict__ counts, const int32_t *__restrict__ bad,
                                      const float *__restrict__ scale, const int32_t *__restrict__ dims,
                                      const int32_t *__restrict__ active, int M, int K, int dmax,
                                      float *__restrict__ cb) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (i64)M * K * dmax) return;
  const int m = (int)(t / ((i64)K * dmax));
  if (active && !active[m]) return;
  const int j = (int)(t % dmax);
  const int c = counts[t / dmax];
  float out = 0.0f;
  if (c > 0 && j < dims[m]) {
    const long long s = (long long)sums[t] - (long long)c * (1LL << FIX_BITS);
    out = (float)(((double)s / (double)scale[m]) / (double)c);
    if (bad[t / dmax]) out = __int_as_float(0x7fc00000);
  }
  cb[t] = out;
}

}  // namespace upd
}  // namespace gulon
