// kmeans.cuh -- Lloyd k-means pieces and PQ encoding, exact-fp32 CUDA-core kernels.
//
// Reference: KMeans.apply offsets (G/KMeans.scala:170-186), KMeans.assign (G/KMeans.scala:24-55,
// 70-98), KMeans.fromAssignment (G/KMeans.scala:198-226), ProductQuantizer.encode
// (G/ProductQuantizer.scala:25-35) + Coder8 (G/Coder.scala:129-140).
//
// Exactness contract (JVM numeric model: binary32, RN, no FMA, no reassociation):
//   score_k = fl(off_k - 2*dot_k),  dot_k = (((0 + x0*c0) + x1*c1) + ...) sequential;
//   argmin over k ascending with strict '<' starting from Float.MaxValue => lowest index on ties;
//   a row where nothing is accepted (all scores NaN or >= MaxValue) yields 0 (fresh array).
#pragma once
#include <float.h>

#include "common.cuh"
#include "select.cuh"

namespace gulon {

// ---- offsets ---------------------------------------------------------------------------------
__global__ void offsets_kernel(const float *__restrict__ cb, const int32_t *__restrict__ dim,
                               int M, int K, int dmax, float *__restrict__ off) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (i64)M * K) return;
  int m = (int)(t / K);
  const float *c = cb + t * dmax;
  float s = 0.0f;
  for (int j = 0; j < dim[m]; j++) s = __fadd_rn(s, __fmul_rn(c[j], c[j]));
  off[t] = s;
}

// ---- exact assignment, dim <= 16 ---------------------------------------------------------------
// grid (row tiles of 512, number of sub-quantizers in `subs`), block 256, 2 rows per thread.
// Centroids of the CTA's sub-quantizer are staged in shared memory in chunks of KC (padded to a
// multiple of 4 floats so the inner loop reads them as broadcast LDS.128).
constexpr int ASSIGN_NT = 256;
constexpr int ASSIGN_ROWS = 2;
constexpr int ASSIGN_TILE = ASSIGN_NT * ASSIGN_ROWS;

template <int DIM, typename OutT>
__global__ void __launch_bounds__(ASSIGN_NT) assign_exact_kernel(
    const float *__restrict__ X, i64 N, i64 ld, const float *__restrict__ cb,
    const float *__restrict__ off, int K, int dmax, const int32_t *__restrict__ subs,
    const int32_t *__restrict__ from, OutT *__restrict__ out, i64 out_stride, int KC) {
  constexpr int DP = (DIM + 3) & ~3;
  extern __shared__ __align__(16) float sm[];
  float *sc = sm;              // [KC][DP]
  float *so = sm + KC * DP;    // [KC]
  const int m = subs[blockIdx.y];
  const int f = from[m];
  const int tid = threadIdx.x;

  float x[ASSIGN_ROWS][DIM];
  i64 rows[ASSIGN_ROWS];
#pragma unroll
  for (int r = 0; r < ASSIGN_ROWS; r++) {
    rows[r] = (i64)blockIdx.x * ASSIGN_TILE + r * ASSIGN_NT + tid;
    const float *src = X + (rows[r] < N ? rows[r] : 0) * ld + f;
#pragma unroll
    for (int j = 0; j < DIM; j++) x[r][j] = src[j];
  }
  float best[ASSIGN_ROWS];
  int idx[ASSIGN_ROWS];
#pragma unroll
  for (int r = 0; r < ASSIGN_ROWS; r++) {
    best[r] = FLT_MAX;
    idx[r] = 0;
  }
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int kn = K - k0 < KC ? K - k0 : KC;
    __syncthreads();
    for (int t = tid; t < kn * DP; t += ASSIGN_NT) {
      int kk = t / DP, j = t % DP;
      sc[t] = j < DIM ? cb[((i64)m * K + k0 + kk) * dmax + j] : 0.0f;
    }
    for (int t = tid; t < kn; t += ASSIGN_NT) so[t] = off[(i64)m * K + k0 + t];
    __syncthreads();
    for (int kk = 0; kk < kn; kk++) {
      float c[DP];
#pragma unroll
      for (int j4 = 0; j4 < DP / 4; j4++) {
        float4 v = *reinterpret_cast<const float4 *>(sc + kk * DP + 4 * j4);
        c[4 * j4 + 0] = v.x;
        c[4 * j4 + 1] = v.y;
        c[4 * j4 + 2] = v.z;
        c[4 * j4 + 3] = v.w;
      }
      const float o = so[kk];
#pragma unroll
      for (int r = 0; r < ASSIGN_ROWS; r++) {
        float d = 0.0f;
#pragma unroll
        for (int j = 0; j < DIM; j++) d = __fadd_rn(d, __fmul_rn(x[r][j], c[j]));
        const float sdist = __fsub_rn(o, __fmul_rn(2.0f, d));
        if (sdist < best[r]) {
          best[r] = sdist;
          idx[r] = k0 + kk;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ASSIGN_ROWS; r++)
    if (rows[r] < N) out[(i64)m * out_stride + rows[r]] = (OutT)idx[r];
}

// ---- exact assignment, any dim (coarse quantizers; slow generic path) ----------------------------
// One row per thread; the dot products of a chunk of KB centroids are carried in registers while
// the row streams through in slices of 8 floats, so every (row, k) sum is still sequential in j.
constexpr int AG_KB = 16;
template <typename OutT>
__global__ void __launch_bounds__(128) assign_generic_kernel(
    const float *__restrict__ X, i64 N, i64 ld, const float *__restrict__ cb,
    const float *__restrict__ off, int K, int dmax, const int32_t *__restrict__ subs,
    const int32_t *__restrict__ from, const int32_t *__restrict__ dims, OutT *__restrict__ out,
    i64 out_stride) {
  extern __shared__ __align__(16) float sm[];  // [AG_KB][dimP]
  const int m = subs[blockIdx.y];
  const int f = from[m], dm = dims[m];
  const int tid = threadIdx.x;
  const i64 row = (i64)blockIdx.x * 128 + tid;
  const float *xr = X + (row < N ? row : 0) * ld + f;
  float best = FLT_MAX;
  int idx = 0;
  for (int k0 = 0; k0 < K; k0 += AG_KB) {
    const int kn = K - k0 < AG_KB ? K - k0 : AG_KB;
    __syncthreads();
    for (int t = tid; t < kn * dm; t += 128) {
      int kk = t / dm, j = t % dm;
      sm[kk * dm + j] = cb[((i64)m * K + k0 + kk) * dmax + j];
    }
    __syncthreads();
    float d[AG_KB];
#pragma unroll
    for (int kk = 0; kk < AG_KB; kk++) d[kk] = 0.0f;
    for (int j0 = 0; j0 < dm; j0 += 8) {
      float xv[8];
#pragma unroll
      for (int j = 0; j < 8; j++) xv[j] = j0 + j < dm ? xr[j0 + j] : 0.0f;
#pragma unroll
      for (int kk = 0; kk < AG_KB; kk++) {
        if (kk < kn) {
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j0 + j < dm) d[kk] = __fadd_rn(d[kk], __fmul_rn(xv[j], sm[kk * dm + j0 + j]));
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < AG_KB; kk++) {
      if (kk < kn) {
        const float sdist = __fsub_rn(off[(i64)m * K + k0 + kk], __fmul_rn(2.0f, d[kk]));
        if (sdist < best) {
          best = sdist;
          idx = k0 + kk;
        }
      }
    }
  }
  if (row < N) out[(i64)m * out_stride + row] = (OutT)idx;
}

// ---- convergence test: Arrays.equals(prev, next), G/KMeans.scala:149 ----------------------------
// diff[slot] += number of rows whose assignment changed; grid (tiles, nsub)
__global__ void count_diff_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b,
                                  i64 N, i64 stride, const int32_t *__restrict__ subs,
                                  int32_t *__restrict__ diff) {
  const int m = subs[blockIdx.y];
  int local = 0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (i64)gridDim.x * blockDim.x)
    local += a[(i64)m * stride + i] != b[(i64)m * stride + i];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(diff + m, local);
}

// ---- centroid update, sum/count form (GULON_UPDATE_SUM) -----------------------------------------
// Deterministic segmented sum without atomics: every warp owns a contiguous row range and a
// private [K][dim] accumulator in shared memory, lanes = dimensions (lane j is the only thread
// that ever touches column j, so there is no cross-lane hazard), rows visited in order; the
// warps of a CTA are then folded in warp order and the CTAs by update_reduce_kernel in CTA order.
// grid (B, nsub), block 32*W; dynamic smem W*K*(dim+1)*4 bytes.
__global__ void update_partial_kernel(const float *__restrict__ X, i64 N, i64 ld,
                                      const int32_t *__restrict__ assign, i64 astride,
                                      const int32_t *__restrict__ subs,
                                      const int32_t *__restrict__ from,
                                      const int32_t *__restrict__ dims, int K, int dmax,
                                      float *__restrict__ part_sum, int32_t *__restrict__ part_cnt) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int W = blockDim.x >> 5;
  const int m = subs[blockIdx.y];
  const int f = from[m], dm = dims[m];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *acc = reinterpret_cast<float *>(smraw) + (size_t)w * K * dm;
  int *cnt = reinterpret_cast<int *>(reinterpret_cast<float *>(smraw) + (size_t)W * K * dm) +
             (size_t)w * K;
  for (int t = lane; t < K * dm; t += 32) acc[t] = 0.0f;
  for (int t = lane; t < K; t += 32) cnt[t] = 0;
  __syncwarp();
  const i64 per_cta = ceil_div(N, (i64)gridDim.x);
  const i64 c0 = (i64)blockIdx.x * per_cta;
  const i64 c1 = c0 + per_cta < N ? c0 + per_cta : N;
  const i64 per_w = ceil_div(per_cta, (i64)W);
  i64 r0 = c0 + (i64)w * per_w;
  i64 r1 = r0 + per_w < c1 ? r0 + per_w : c1;
  const int32_t *a = assign + (i64)m * astride;
  for (int jb = 0; jb < dm; jb += 32) {   // dims beyond 32 handled in lane blocks
    const int j = jb + lane;
    for (i64 r = r0; r < r1; r += 4) {
      int ai[4];
      float xv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const bool ok = r + u < r1;
        ai[u] = ok ? a[r + u] : -1;
        xv[u] = (ok && j < dm) ? X[(r + u) * ld + f + j] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (ai[u] >= 0) {
          if (j < dm) acc[ai[u] * dm + j] = __fadd_rn(acc[ai[u] * dm + j], xv[u]);
          if (jb == 0 && lane == 0) cnt[ai[u]] += 1;
        }
      }
    }
  }
  __syncthreads();
  // fold the W warps in order, write the CTA partial
  float *all = reinterpret_cast<float *>(smraw);
  int *allc = reinterpret_cast<int *>(all + (size_t)W * K * dm);
  const i64 slot = (i64)blockIdx.y * gridDim.x + blockIdx.x;
  for (int t = threadIdx.x; t < K * dm; t += blockDim.x) {
    float sacc = all[t];
    for (int ww = 1; ww < W; ww++) sacc = __fadd_rn(sacc, all[(size_t)ww * K * dm + t]);
    part_sum[slot * K * dmax + (t / dm) * dmax + (t % dm)] = sacc;
  }
  for (int t = threadIdx.x; t < K; t += blockDim.x) {
    int c = 0;
    for (int ww = 0; ww < W; ww++) c += allc[(size_t)ww * K + t];
    part_cnt[slot * K + t] = c;
  }
}

// sums[m][k][j] = sum over CTAs b (in order) of part_sum; counts likewise.  grid (K, nsub)
__global__ void update_reduce_kernel(const float *__restrict__ part_sum,
                                     const int32_t *__restrict__ part_cnt, int B,
                                     const int32_t *__restrict__ subs,
                                     const int32_t *__restrict__ dims, int K, int dmax,
                                     float *__restrict__ sums, int32_t *__restrict__ counts) {
  const int m = subs[blockIdx.y];
  const int k = blockIdx.x;
  const i64 base = (i64)blockIdx.y * B;
  for (int j = threadIdx.x; j < dmax; j += blockDim.x) {
    float s = 0.0f;
    if (j < dims[m])
      for (int b = 0; b < B; b++) s = __fadd_rn(s, part_sum[((base + b) * K + k) * dmax + j]);
    sums[((i64)m * K + k) * dmax + j] = s;
  }
  if (threadIdx.x == 0) {
    int c = 0;
    for (int b = 0; b < B; b++) c += part_cnt[(base + b) * K + k];
    counts[(i64)m * K + k] = c;
  }
}

// counts[subs[y]][k] = cnt[y][k]  (running-mean path keeps its counts per launch slot)
__global__ void scatter_counts_kernel(const int32_t *__restrict__ cnt,
                                      const int32_t *__restrict__ subs, int ns, int K,
                                      int32_t *__restrict__ counts) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (i64)ns * K) return;
  counts[(i64)subs[t / K] * K + t % K] = cnt[t];
}

// centroid = sum / count (empty cluster -> all-zero, as the reference leaves it). grid covers M*K*dmax
__global__ void update_finalize_kernel(const float *__restrict__ sums,
                                       const int32_t *__restrict__ counts,
                                       const int32_t *__restrict__ active, int M, int K, int dmax,
                                       float *__restrict__ cb) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (i64)M * K * dmax) return;
  const int m = (int)(t / ((i64)K * dmax));
  if (active && !active[m]) return;
  const int c = counts[t / dmax];
  cb[t] = c > 0 ? __fdiv_rn(sums[t], (float)c) : 0.0f;
}

// ---- centroid update, literal running mean (GULON_UPDATE_RUNNING_MEAN) --------------------------
// c <- p + (x - p)/n over the cluster's rows IN ROW ORDER is a sequential recurrence per
// (cluster, dimension) but the M*K*dim chains are independent.  Rows are first grouped by
// cluster with a stable counting sort (hist -> scan -> scatter), then one thread walks each chain.
constexpr int RM_TILE = 8192;  // rows per tile; one warp scatters a tile in order

// grid (tiles, nsub), block 256: tile_hist[(y*tiles + tile)*K + a]
__global__ void rm_hist_kernel(const int32_t *__restrict__ assign, i64 astride, i64 N,
                               const int32_t *__restrict__ subs, int K,
                               int32_t *__restrict__ tile_hist) {
  extern __shared__ int sh[];
  const int m = subs[blockIdx.y];
  for (int t = threadIdx.x; t < K; t += blockDim.x) sh[t] = 0;
  __syncthreads();
  const i64 r0 = (i64)blockIdx.x * RM_TILE;
  const i64 r1 = r0 + RM_TILE < N ? r0 + RM_TILE : N;
  for (i64 r = r0 + threadIdx.x; r < r1; r += blockDim.x)
    atomicAdd(&sh[assign[(i64)m * astride + r]], 1);
  __syncthreads();
  int32_t *dst = tile_hist + ((i64)blockIdx.y * gridDim.x + blockIdx.x) * K;
  for (int t = threadIdx.x; t < K; t += blockDim.x) dst[t] = sh[t];
}

// grid (nsub), block 256: per cluster, exclusive prefix over tiles; then cluster bases.
// tile_hist becomes the within-cluster offset of each tile; counts/base are [nsub][K].
__global__ void rm_scan_kernel(int32_t *__restrict__ tile_hist, int tiles, int K,
                               int32_t *__restrict__ counts, int32_t *__restrict__ base) {
  const int y = blockIdx.x;
  int32_t *th = tile_hist + (i64)y * tiles * K;
  for (int a = threadIdx.x; a < K; a += blockDim.x) {
    int run = 0;
    for (int t = 0; t < tiles; t++) {
      int v = th[(i64)t * K + a];
      th[(i64)t * K + a] = run;
      run += v;
    }
    counts[(i64)y * K + a] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int a = 0; a < K; a++) {
      base[(i64)y * K + a] = run;
      run += counts[(i64)y * K + a];
    }
  }
}

// grid (ceil(tiles/4), nsub), block 128 (4 warps, one tile each): stable scatter of row ids.
__global__ void rm_scatter_kernel(const int32_t *__restrict__ assign, i64 astride, i64 N,
                                  const int32_t *__restrict__ subs, int K, int tiles,
                                  const int32_t *__restrict__ tile_hist,
                                  const int32_t *__restrict__ base, int32_t *__restrict__ order,
                                  i64 ostride) {
  extern __shared__ int sh[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tile = blockIdx.x * 4 + w;
  if (tile >= tiles) return;
  const int y = blockIdx.y, m = subs[y];
  int *cnt = sh + w * K;
  const int32_t *th = tile_hist + ((i64)y * tiles + tile) * K;
  for (int t = lane; t < K; t += 32) cnt[t] = th[t] + base[(i64)y * K + t];
  __syncwarp();
  const i64 r0 = (i64)tile * RM_TILE;
  const i64 r1 = r0 + RM_TILE < N ? r0 + RM_TILE : N;
  for (i64 rb = r0; rb < r1; rb += 32) {
    const i64 r = rb + lane;
    const bool ok = r < r1;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int a = assign[(i64)m * astride + r];
      const unsigned mask = __match_any_sync(act, a);
      const int leader = __ffs(mask) - 1;
      const int rank = __popc(mask & ((1u << lane) - 1));
      int old = 0;
      if (lane == leader) {
        old = cnt[a];
        cnt[a] = old + __popc(mask);
      }
      old = __shfl_sync(mask, old, leader);
      order[(i64)y * ostride + old + rank] = (int32_t)r;
    }
    __syncwarp();
  }
}

// One thread per (cluster, dimension): walks the cluster's rows in order.
// grid (ceil(K*dim/128), nsub), block 128; thread t -> cluster t / dim, dimension t % dim.
__global__ void __launch_bounds__(128) rm_chain_kernel(
    const float *__restrict__ X, i64 ld, const int32_t *__restrict__ subs,
    const int32_t *__restrict__ from, const int32_t *__restrict__ dims, int K, int dmax,
    const int32_t *__restrict__ order, i64 ostride, const int32_t *__restrict__ counts,
    const int32_t *__restrict__ base, float *__restrict__ cb) {
  const int y = blockIdx.y, m = subs[y];
  const int dm = dims[m], f = from[m];
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= K * dm) return;
  const int a = t / dm, j = t % dm;
  const int n = counts[(i64)y * K + a];
  const int32_t *ord = order + (i64)y * ostride + base[(i64)y * K + a];
  const float *xcol = X + f + j;
  float p = 0.0f;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    int rr[8];
    float xv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) rr[u] = ord[i + u];
#pragma unroll
    for (int u = 0; u < 8; u++) xv[u] = xcol[(i64)rr[u] * ld];
#pragma unroll
    for (int u = 0; u < 8; u++)
      p = __fadd_rn(p, __fdiv_rn(__fsub_rn(xv[u], p), __int2float_rn(i + u + 1)));
  }
  for (; i < n; i++) {
    const float xv = xcol[(i64)ord[i] * ld];
    p = __fadd_rn(p, __fdiv_rn(__fsub_rn(xv, p), __int2float_rn(i + 1)));
  }
  cb[((i64)m * K + a) * dmax + j] = p;
}

// gather K rows' column windows into a codebook slot (KMeans.init). grid (K), block 32+
__global__ void gather_rows_kernel(const float *__restrict__ X, i64 ld, int from, int dim,
                                   const i64 *__restrict__ rows, i64 row_offset, i64 n_local,
                                   float *__restrict__ dst, int dmax) {
  const int k = blockIdx.x;
  const i64 r = rows[k] - row_offset;
  for (int j = threadIdx.x; j < dmax; j += blockDim.x) {
    float v = 0.0f;
    if (j < dim && r >= 0 && r < n_local) v = X[r * ld + from + j];
    dst[(i64)k * dmax + j] = v;
  }
}

// exact squared distance keys for Index.exactNearestNeighbours (G/Index.scala:209-229):
// sum_i fl(dx*dx), dx = fl(q_i - x_i), sequential.  grid (n_pad/128, nq), block 128; the query
// vector sits in shared memory, each thread walks one row.
__global__ void __launch_bounds__(128) exact_keys_kernel(const float *__restrict__ X, i64 ld,
                                                         int D, i64 from, i64 until,
                                                         const float *__restrict__ Q, i64 ldq,
                                                         u64 *__restrict__ keys, i64 n_pad) {
  extern __shared__ float sq[];
  const i64 q = blockIdx.y;
  for (int j = threadIdx.x; j < D; j += 128) sq[j] = Q[q * ldq + j];
  __syncthreads();
  const i64 t = (i64)blockIdx.x * 128 + threadIdx.x;
  const i64 row = from + t;
  u64 key = KEY_SENT;
  if (row < until) {
    const float *x = X + row * ld;
    float s = 0.0f;
    for (int j = 0; j < D; j++) {
      const float dx = __fsub_rn(sq[j], x[j]);
      s = __fadd_rn(s, __fmul_rn(dx, dx));
    }
    key = make_key(s, (uint32_t)row);
  }
  keys[q * n_pad + t] = key;
}

// exact squared distances of listed candidate rows (re-rank, G/Tests.scala:24-37 +
// G/MathUtils.scala:85-95).  grid (n_pad/128, nq), block 128; cand [nq][R], ids < 0 skipped.
__global__ void __launch_bounds__(128) rerank_keys_kernel(const float *__restrict__ X, i64 ld,
                                                          int D, i64 N,
                                                          const float *__restrict__ Q, i64 ldq,
                                                          const int32_t *__restrict__ cand, int R,
                                                          u64 *__restrict__ keys, i64 n_pad) {
  extern __shared__ float sq[];
  const i64 q = blockIdx.y;
  for (int j = threadIdx.x; j < D; j += 128) sq[j] = Q[q * ldq + j];
  __syncthreads();
  const i64 t = (i64)blockIdx.x * 128 + threadIdx.x;
  u64 key = KEY_SENT;
  if (t < R) {
    const int32_t id = cand[q * R + t];
    if (id >= 0 && id < N) {
      const float *x = X + (i64)id * ld;
      float s = 0.0f;
      for (int j = 0; j < D; j++) {
        const float dx = __fsub_rn(sq[j], x[j]);
        s = __fadd_rn(s, __fmul_rn(dx, dx));
      }
      key = make_key(s, (uint32_t)id);
    }
  }
  keys[q * n_pad + t] = key;
}

// Fused re-rank: exact distances of a query's R candidates and their k best, one CTA per query, no key
// materialisation in global memory.  cand [nq][R] holds GLOBAL row ids; this device owns rows
// [id_lo, id_lo + N) (X row 0 = global row id_lo); ids outside (other shards' rows, or -1) are skipped.
// The reference's distance is a SEQUENTIAL fp32 sum over the D coordinates (G/MathUtils.scala:85-95), so
// one thread must walk a whole row -- but rows are 4 KB apart, and a lane that reads its own row
// touches one sector per request.  Loads are therefore warp-cooperative: for each of its 32 candidates the
// warp reads a 128-byte stretch of the row (one coalesced request), parks it in a padded shared-memory
// tile, and then every lane sums ITS candidate's 32 coordinates from the tile in order.  (Round 2: the
// lane-per-row loads ran at 11 % of HBM.)  The keys are sorted in shared memory; the k smallest go
// out as (global id, distance), (distance, id) ascending.
// grid nq, block 128; dynamic shared memory: P u64 + D floats + 4 x 32 x 33 floats; P = power of two >= R, R <= 1024.
constexpr int RERANK_RMAX = 1024;
constexpr int RERANK_TILE = 32 * 33;
__global__ void __launch_bounds__(128) rerank_topk_kernel(const float *__restrict__ X, i64 ld, int D,
                                                          i64 N, i64 id_lo,
                                                          const float *__restrict__ Q, i64 ldq,
                                                          const int32_t *__restrict__ cand, int R, int P,
                                                          int k, int32_t *__restrict__ ids,
                                                          float *__restrict__ dists,
                                                          int32_t *__restrict__ sizes) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  u64 *keys = reinterpret_cast<u64 *>(rr_smem);
  float *sq = reinterpret_cast<float *>(rr_smem + (size_t)P * sizeof(u64));
  float *tiles = sq + ((D + 3) & ~3);
  const i64 q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *tile = tiles + warp * RERANK_TILE;
  for (int j = tid; j < D; j += 128) sq[j] = Q[q * ldq + j];
  for (int t = tid; t < P; t += 128) keys[t] = KEY_SENT;
  __syncthreads();
  for (int c0 = warp * 32; c0 < R; c0 += 128) {
    const int ci = c0 + lane;
    i64 gid = -1, row = -1;
    if (ci < R) {
      gid = cand[q * R + ci];
      row = gid - id_lo;
      if (gid < 0 || row < 0 || row >= N) row = -1;
    }
    if (!__any_sync(0xffffffffu, row >= 0)) continue;
    float s = 0.0f;
    for (int j0 = 0; j0 < D; j0 += 32) {
      // cooperative loads: candidate c's coordinates [j0, j0 + 32) -> tile[c][0..32); all 32 requests of
      // the warp are issued before the first one is stored (32 x 128 bytes in flight per warp)
      float v[32];
      const bool in = j0 + lane < D;
#pragma unroll
      for (int c = 0; c < 32; c++) {
        const i64 rc = __shfl_sync(0xffffffffu, row, c);
        v[c] = (in && rc >= 0) ? __ldg(X + rc * ld + j0 + lane) : 0.0f;
      }
#pragma unroll
      for (int c = 0; c < 32; c++) tile[c * 33 + lane] = v[c];
      __syncwarp();
      if (row >= 0) {
        const int w = D - j0 < 32 ? D - j0 : 32;
        const float *mine = tile + lane * 33;
        if (w == 32) {
#pragma unroll
          for (int t = 0; t < 32; t++) {
            const float dx = __fsub_rn(sq[j0 + t], mine[t]);
            s = __fadd_rn(s, __fmul_rn(dx, dx));
          }
        } else {
          for (int t = 0; t < w; t++) {
            const float dx = __fsub_rn(sq[j0 + t], mine[t]);
            s = __fadd_rn(s, __fmul_rn(dx, dx));
          }
        }
      }
      __syncwarp();
    }
    if (row >= 0) keys[ci] = make_key(s, (uint32_t)gid);
  }
  __syncthreads();
  block_bitonic_sort(keys, P, tid, 128);
  int cnt = 0;
  for (int i = tid; i < k; i += 128) {
    const u64 key = i < P ? keys[i] : KEY_SENT;
    const bool ok = key != KEY_SENT;
    ids[q * k + i] = ok ? (int32_t)(uint32_t)key : -1;
    dists[q * k + i] = ok ? ord2f((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
    cnt += ok;
  }
  if (sizes) {
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (tid == 0) sizes[q] = s_cnt;
  }
}

// codes [M][ps] -> out [N][ldo]: ProductQuantizer.decode, G/ProductQuantizer.scala:58-78
template <typename CodeT>
__global__ void decode_kernel(const CodeT *__restrict__ codes, i64 ps, i64 N,
                              const float *__restrict__ cb, const int32_t *__restrict__ from,
                              const int32_t *__restrict__ dims, int M, int K, int dmax,
                              float *__restrict__ out, i64 ldo) {
  const i64 row = (i64)blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= N) return;
  for (int m = 0; m < M; m++) {
    const int c = codes[(i64)m * ps + row];
    const float *src = cb + ((i64)m * K + c) * dmax;
    for (int j = threadIdx.x; j < dims[m]; j += blockDim.x) out[row * ldo + from[m] + j] = src[j];
  }
}

}  // namespace gulon
