// pscan.cuh -- pruned scan: an integer lower-bound pass over all rows, exact fp32 re-evaluation of
// the survivors only.  Results are identical to fscan::fused_scan_kernel (and to the reference:
// PQIndex.distances + batchQuery, G/Index.scala:393-440) because every distance that can enter a
// top-k list is still computed as (((0 + LUT[0][c0]) + LUT[1][c1]) + ...) in fp32, round-to-nearest.
//
// Why: the exact scan is bound by shared-memory bandwidth -- one 4-byte table read per (row, query,
// quantizer).  A row can be discarded without computing its distance if a LOWER BOUND of it already
// exceeds the query's current k-th best distance tau.  The bound used here is
//     LB = base + delta * sum_m q[m][code_m],   q = min(qmax, floor((LUT[m][c] - min_c LUT[m][c]) / delta)),
// base = sum_m min_c LUT[m][c].  The quantised entries are FB = 8 or 16 bits wide, so one 128-bit
// shared-memory read serves QT = 16 (FB = 8) or 8 (FB = 16) queries, accumulated with plain 32-bit
// integer adds on packed fields (fields cannot carry: bias + M * qmax <= 2^FB - 1).  With
// bias = FLAG - 1 - T, FLAG = 2^(FB-1), T = floor((tau * (1 + 2^-11) - base) / delta), the top bit of
// a field is set iff sum q > T, which implies fp32-sum > tau (the 2^-11 covers the fp32 summation
// error for M <= 2048 and the double rounding of the quantiser).  Rows whose QT fields are all
// flagged are dropped; the others ("survivors", 1e-5 .. 1e-3 of the pairs once tau is warm) are
// re-evaluated exactly by one warp each from the fp32 tables in global memory and offered to the
// top-k lists.  tau comes from an exact scan of the first rows of the range (the "boot" rows, done
// by the fused kernel) and tightens as a CTA's own list fills.
//
// FB = 8 halves the shared-memory bytes per (row, query, quantizer) again but is coarse
// (qmax = 127 / M levels per quantizer); the host uses it while qmax >= 3 and falls back to FB = 16.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "select.cuh"

namespace gulon {
namespace pscan {

constexpr int NT = 512;
constexpr int KMAX = 128;
constexpr int SORTN = 256;      // per-query sort area: k list entries + up to CAPQ candidates
constexpr int CAPQ = SORTN - KMAX;
constexpr int REGION_BYTES = 65536;  // both replicated slice buffers, interleaved (see the kernel)
constexpr int SLOTS = 256;           // items whose survivors may wait in the queue

// FB = bits per lower-bound field, W = 32-bit words per table entry (one LDS.32/64/128 per row).
//   queries per tile QT = W * 32 / FB; replicas of an entry = 128 B / (4 W) so that every lane of a
//   wavefront reads its own bank group; rows per thread RPT = 64 / W (64 accumulator registers).
// W = 4 serves the most queries per pass (best when many queries wait); W = 1 makes a pass 4x
// shorter: the code planes then stream from HBM at a large fraction of its bandwidth, the
// low-latency shape for small query batches.
template <int FB, int W>
struct Cfg {
  static constexpr int QT = W * 32 / FB;         // queries per tile
  static constexpr int FPW = 32 / FB;            // fields per 32-bit word
  static constexpr int RPT = 64 / W;             // rows per thread
  static constexpr int R = NT * RPT;             // rows per item
  static constexpr int NREP = 32 / W;            // replicas of a table entry in a slice
  static constexpr int FLAG = 1 << (FB - 1);
  static constexpr uint32_t FLAGMASK = FB == 8 ? 0x80808080u : 0x80008000u;
  static constexpr int QSH = QT == 16 ? 4 : QT == 8 ? 3 : 2;       // log2(QT)
  static constexpr int RSH = RPT == 16 ? 13 : RPT == 32 ? 14 : 15;  // log2(R)
  static constexpr int QCAP = NT * QT;           // survivor queue (one row per thread in the slow path)
  // queue length that triggers an exact re-evaluation pass (>= NT: a pass evaluates one per thread)
  static constexpr int DRAIN_AT = QCAP / 4 < 2048 ? QCAP / 4 : 2048;
  // up to 64 KB of alignment slack in front of the 64 KB-aligned slice region
  static constexpr int SMEM_BYTES = 65536 + REGION_BYTES + QT * SORTN * 8 + QCAP * 4;
  static_assert(QT >= 4 && QT % 4 == 0, "tiles are made of query groups of 4");
};
constexpr int R_MAX = NT * 64;

// quantisation units between base and the boot threshold
inline int t0_units(int FB, int ML) {
  if (FB == 16) return 2048;
  const int qmax = 127 / ML;
  return std::max(1, ML * qmax / 2);
}

struct QParam {
  double base;       // sum_m min_c LUT[m][c]
  double inv_delta;  // 0: nothing can be pruned (no threshold yet)
};

struct Params {
  const uint8_t *codes;
  i64 ps;
  const uint8_t *rowcodes;  // row-major copy [N][rcs] of the planes (survivor evaluation), or null
  i64 rcs;
  const float *sufmin;      // [T*QT][M+1]: sum_{i >= m} min_c LUT[i][c], rounded down
  i64 from, until;  // rows scanned by this kernel (after the boot rows)
  i64 split_len;
  const uint32_t *qlut;   // [T][ML][256][W] QT packed fields
  const int32_t *msel;    // [T][ML] quantizers of the lower bound, ascending
  const float4 *lutI;     // [T*QT/4][M][256] exact tables (scan.cuh layout)
  const QParam *qp;       // [T*QT]
  const u64 *boot_tail;   // [T*QT] key of the boot list tail (KEY_SENT: none)
  const u64 *floor;       // [T*QT] k-chunked scans: only keys > floor enter the lists (null: none)
  // WIDE indexes (256 < K <= 65536, 16-bit centroid ids): `codes` / `lutI` / `qlut` are the planes and
  // tables of the 8-bit GROUP ids (every centroid belongs to one of 256 groups; a group's table entry is
  // the minimum over its members: still a lower bound); survivors are evaluated with the real ids
  const uint16_t *rowcodes16;  // [N][rcs16] row-major 16-bit ids
  i64 rcs16;                   // elements per row (a multiple of 8)
  const float *lutW;           // [T*QT][M][K] exact tables
  int K;
  u64 *lists;             // [S][T*QT][k]
  unsigned long long *stats;  // [0] survivors, [1] list candidates, [2] slow-path items
  i64 nq;
  int M, ML, T, k, S, Bs;
};

// ---- per-query quantisation parameters ---------------------------------------------------------
// grid (G groups of 4 queries), block 256.  mins[(g*4+j)*M + m] = min_c LUT; qp / boot_tail per query.
__global__ void __launch_bounds__(256) qparams_kernel(const float4 *__restrict__ lutI, int M, int K,
                                                      i64 nq, const u64 *__restrict__ boot_keys,
                                                      i64 boot_stride, int k, int t0,
                                                      float *__restrict__ mins,
                                                      float *__restrict__ spread,
                                                      float *__restrict__ sufmin,
                                                      QParam *__restrict__ qp,
                                                      u64 *__restrict__ boot_tail) {
  __shared__ float red[8][4];
  __shared__ float reds[8][4];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float INF = __int_as_float(0x7f800000);
  double base = 0.0;
  for (int m = 0; m < M; m++) {
    float4 v = make_float4(INF, INF, INF, INF);
    float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < K) {
      v = lutI[((i64)g * M + m) * 256 + tid];
      sm = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      v.x = fminf(v.x, __shfl_xor_sync(0xffffffffu, v.x, o));
      v.y = fminf(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
      v.z = fminf(v.z, __shfl_xor_sync(0xffffffffu, v.z, o));
      v.w = fminf(v.w, __shfl_xor_sync(0xffffffffu, v.w, o));
      sm.x += __shfl_xor_sync(0xffffffffu, sm.x, o);
      sm.y += __shfl_xor_sync(0xffffffffu, sm.y, o);
      sm.z += __shfl_xor_sync(0xffffffffu, sm.z, o);
      sm.w += __shfl_xor_sync(0xffffffffu, sm.w, o);
    }
    __syncthreads();
    if (lane == 0) {
      red[warp][0] = v.x; red[warp][1] = v.y; red[warp][2] = v.z; red[warp][3] = v.w;
      reds[warp][0] = sm.x; reds[warp][1] = sm.y; reds[warp][2] = sm.z; reds[warp][3] = sm.w;
    }
    __syncthreads();
    if (tid < 4) {
      float mn = red[0][tid], su = reds[0][tid];
      for (int w = 1; w < 8; w++) {
        mn = fminf(mn, red[w][tid]);
        su += reds[w][tid];
      }
      mins[((i64)g * 4 + tid) * M + m] = mn;
      // mean_c LUT - min_c LUT: what a random row adds to the lower bound through this quantizer
      // (only ranks the quantizers for the subset choice; never enters a distance)
      spread[((i64)g * 4 + tid) * M + m] = su / (float)(K > 0 ? K : 1) - mn;
      base += (double)mn;
    }
  }
  if (tid < 4) {
    const i64 q = (i64)g * 4 + tid;
    // suffix sums of the minima (what the quantizers not yet summed add at least), rounded down
    double suf = 0.0;
    sufmin[q * (M + 1) + M] = 0.0f;
    for (int m = M - 1; m >= 0; m--) {
      suf += (double)mins[q * M + m];
      sufmin[q * (M + 1) + m] = __double2float_rd(suf);
    }
    QParam p;
    p.base = base;
    p.inv_delta = 0.0;
    u64 tk = KEY_SENT;
    if (q < nq) {
      tk = boot_keys[q * boot_stride + (k - 1)];
      if (tk != KEY_SENT) {
        const double tau0 = (double)ord2f((uint32_t)(tk >> 32));
        if (tau0 == tau0 && tau0 < 3.0e38 && base == base && base < 3.0e38) {
          double delta = (tau0 - base) / (double)t0;
          const double floor1 = tau0 * (1.0 / 1048576.0);
          if (!(delta > floor1)) delta = floor1;
          if (!(delta > 1e-30)) delta = 1e-30;
          p.inv_delta = 1.0 / delta;
        }
      }
    }
    qp[q] = p;
    boot_tail[q] = tk;
  }
}

// The lower bound may leave quantizers out (they then contribute their minimum, which `base` already
// holds): LB = base + delta * sum_{m in SEL} q[m][code_m] is still a lower bound of the distance, and
// the pass reads ML = |SEL| <= M code planes and table slices instead of M.  Each tile keeps the ML
// quantizers that add most, on average, to the bound of its queries, measured in units of each
// query's slack tau - base.  grid T, block 256; msel[t][ML] ascending.
constexpr int MSEL_MAX = 1024;
__global__ void __launch_bounds__(256) qselect_kernel(const float *__restrict__ spread,
                                                      const QParam *__restrict__ qp, int M, int ML,
                                                      int QT, int32_t *__restrict__ msel) {
  __shared__ float score[MSEL_MAX];
  __shared__ unsigned char pick[MSEL_MAX];
  const int t = blockIdx.x, tid = threadIdx.x;
  for (int m = tid; m < M; m += 256) {
    float sc = 0.f;
    for (int f = 0; f < QT; f++) {
      const i64 q = (i64)t * QT + f;
      const double w = qp[q].inv_delta;
      float a = spread[q * M + m] * (w > 0.0 ? (float)w : 0.f);
      if (!(a > 0.f)) a = 0.f;            // NaN / negative: no contribution
      if (a > 1e30f) a = 1e30f;
      sc += a;
    }
    score[m] = sc;
  }
  __syncthreads();
  for (int m = tid; m < M; m += 256) {
    const float sc = score[m];
    int rank = 0;
    for (int o = 0; o < M; o++) {
      const float so = score[o];
      rank += (so > sc || (so == sc && o < m)) ? 1 : 0;
    }
    pick[m] = rank < ML ? 1 : 0;
  }
  __syncthreads();
  for (int m = tid; m < M; m += 256) {
    if (pick[m]) {
      int pos = 0;
      for (int o = 0; o < m; o++) pos += pick[o];
      msel[(i64)t * ML + pos] = m;
    }
  }
}

// grid (T, ML), block 256 (thread = code): the QT quantised entries of a tile, field f = query f.
template <int FB, int W>
__global__ void __launch_bounds__(256) qlut_build_kernel(const float4 *__restrict__ lutI,
                                                         const float *__restrict__ mins,
                                                         const QParam *__restrict__ qp,
                                                         const int32_t *__restrict__ msel, int M,
                                                         int ML, int K,
                                                         uint32_t *__restrict__ qlut) {
  using C = Cfg<FB, W>;
  const int t = blockIdx.x, c = threadIdx.x;
  const int m = msel[(i64)t * ML + blockIdx.y];
  const int qmax = (C::FLAG - 1) / ML;
  uint32_t w[W];
#pragma unroll
  for (int i = 0; i < W; i++) w[i] = 0u;
#pragma unroll
  for (int h = 0; h < C::QT / 4; h++) {
    const int g = (C::QT / 4) * t + h;
    const float4 v = lutI[((i64)g * M + m) * 256 + c];
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const i64 q = (i64)g * 4 + j;
      const double d = ((double)vv[j] - (double)mins[q * M + m]) * qp[q].inv_delta;
      int qi = 0;
      if (c < K && d > 0.0) qi = d >= (double)qmax ? qmax : (int)floor(d);
      const int f = h * 4 + j;
      w[f / C::FPW] |= (uint32_t)qi << (FB * (f % C::FPW));
    }
  }
#pragma unroll
  for (int i = 0; i < W; i++) qlut[(((i64)t * ML + blockIdx.y) * 256 + c) * W + i] = w[i];
}

// boot list [rows][boot_stride] + split lists [S][rows][k] -> keys [rows][stride]
__global__ void gather_lists2_kernel(const u64 *__restrict__ lists, int S, i64 rows, int k,
                                     const u64 *__restrict__ boot, i64 boot_stride,
                                     u64 *__restrict__ keys, i64 stride) {
  const i64 q = blockIdx.y;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < stride;
       t += (i64)gridDim.x * blockDim.x) {
    u64 v = KEY_SENT;
    if (t < (i64)S * k) {
      const int s = (int)(t / k), i = (int)(t % k);
      v = lists[((i64)s * rows + q) * k + i];
    } else if (t < (i64)(S + 1) * k) {
      v = boot[q * boot_stride + (t - (i64)S * k)];
    }
    keys[q * stride + t] = v;
  }
}

// planes [M][ps] -> rows [N][rcs] (rcs = M rounded up to 16): a survivor's M codes are one or two
// 32-byte sectors instead of M sectors in M planes.  grid over rows, thread = row.
__global__ void __launch_bounds__(256) rowcodes_kernel(const uint8_t *__restrict__ codes, i64 ps,
                                                       i64 N, int M, i64 rcs,
                                                       uint8_t *__restrict__ rows) {
  const i64 r = (i64)blockIdx.x * 256 + threadIdx.x;
  if (r >= N) return;
  for (int m0 = 0; m0 < M; m0 += 16) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int u = 0; u < 16; u++)
      if (m0 + u < M) w[u >> 2] |= (uint32_t)codes[(i64)(m0 + u) * ps + r] << (8 * (u & 3));
    *reinterpret_cast<uint4 *>(rows + r * rcs + m0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- wide indexes: group ids ------------------------------------------------------------------------
// planes of 16-bit ids [M][ps16] -> planes of 8-bit group ids [M][ps8] through gmap [M][K]
__global__ void __launch_bounds__(256) group_codes_kernel(const uint16_t *__restrict__ codes16, i64 ps16, i64 N,
                                                          int M, int K, const uint8_t *__restrict__ gmap,
                                                          uint8_t *__restrict__ gcodes, i64 ps8) {
  const i64 r = (i64)blockIdx.x * 256 + threadIdx.x;
  const int m = blockIdx.y;
  if (r >= N) return;
  const int c = codes16[(i64)m * ps16 + r];
  gcodes[(i64)m * ps8 + r] = c < K ? gmap[(i64)m * K + c] : 0;
}
// planes [M][ps16] -> rows [N][rcs16]
__global__ void __launch_bounds__(256) rowcodes16_kernel(const uint16_t *__restrict__ codes16, i64 ps16, i64 N,
                                                         int M, i64 rcs16, uint16_t *__restrict__ rows) {
  const i64 r = (i64)blockIdx.x * 256 + threadIdx.x;
  if (r >= N) return;
  for (int m = 0; m < rcs16; m++) rows[r * rcs16 + m] = m < M ? codes16[(i64)m * ps16 + r] : (uint16_t)0;
}
// exact tables lutW [nq][M][K] -> group-minimum tables in the interleaved layout of the scan kernels:
// lutG[(g4 * M + m) * 256 + g] = float4 of queries 4 g4 .. 4 g4 + 3, entry = min over the members of group g
// (+inf for an empty group; 0 for queries past nq).  members [M][K] centroid ids sorted by group, start [M][257].
// grid (G4, M), block 256 (thread = group).
__global__ void __launch_bounds__(256) lut_group_min_kernel(const float *__restrict__ lutW, i64 nq, int M, int K,
                                                            const int32_t *__restrict__ members,
                                                            const int32_t *__restrict__ start,
                                                            float4 *__restrict__ lutG) {
  const int g4 = blockIdx.x, m = blockIdx.y, g = threadIdx.x;
  const int b = start[m * 257 + g], e = start[m * 257 + g + 1];
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const i64 q = (i64)g4 * 4 + j;
    float mn = 0.0f;
    if (q < nq) {
      mn = __int_as_float(0x7f800000);
      const float *l = lutW + (q * M + m) * K;
      for (int i = b; i < e; i++) {
        const float v = l[members[(i64)m * K + i]];
        // a NaN entry must not hide behind fminf: it keeps the group at the minimum of the others, which
        // is still a lower bound of every finite member and the NaN member is re-evaluated exactly anyway
        mn = fminf(mn, v);
      }
    }
    r[j] = mn;
  }
  lutG[((i64)g4 * M + m) * 256 + g] = make_float4(r[0], r[1], r[2], r[3]);
}

// The code planes stream through L2 once per pass: evict-first keeps them from displacing the exact
// tables (tens of MB, re-read at random by the survivor evaluation), which are loaded evict-last.
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ u64 l2_policy_evict_last() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// Same inputs as gather_lists2_kernel, but sorted on the spot: when the (S + 1) * k keys of a query
// fit MERGE_SMALL_MAX, one warp sorts them in shared memory and writes the k smallest -- no 4096-key
// selection pass for a handful of keys.  grid ceil(rows / 4), block 128 (warp = query).
constexpr int MERGE_SMALL_MAX = 1024;
__global__ void __launch_bounds__(128) merge_small_kernel(const u64 *__restrict__ lists, int S, i64 rows,
                                                          int k, const u64 *__restrict__ boot,
                                                          i64 boot_stride, u64 *__restrict__ out) {
  __shared__ u64 sb[4][MERGE_SMALL_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 q = (i64)blockIdx.x * 4 + warp;
  if (q >= rows) return;
  const int total = (S + 1) * k;
  int P = 2;
  while (P < total) P <<= 1;
  for (int t = lane; t < P; t += 32) {
    u64 v = KEY_SENT;
    if (t < S * k) {
      const int s = t / k, i = t % k;
      v = lists[((i64)s * rows + q) * k + i];
    } else if (t < total) {
      v = boot[q * boot_stride + (t - S * k)];
    }
    sb[warp][t] = v;
  }
  __syncwarp();
  warp_bitonic_sort(sb[warp], P, lane);
  for (int i = lane; i < k; i += 32) out[q * k + i] = sb[warp][i];
}

__device__ __forceinline__ uint4 ldg_stream_u4(const void *p, u64 pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_keep_f32(const float *p, u64 pol) {
  float r;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
template <int W>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t (&v)[W]) {
  if constexpr (W == 4) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(addr)
                 : "memory");
  } else if constexpr (W == 2) {
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr) : "memory");
  } else {
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[0]) : "r"(addr) : "memory");
  }
}
// table entry of (tile, quantizer, code) replicated into one 16-byte store value
template <int W>
__device__ __forceinline__ uint4 ldg_entry(const uint32_t *p) {
  if constexpr (W == 4) {
    return ldg_stream_u4(p);
  } else if constexpr (W == 2) {
    const uint2 e = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_uint4(e.x, e.y, e.x, e.y);
  } else {
    const uint32_t e = __ldg(p);
    return make_uint4(e, e, e, e);
  }
}

// Shared-memory layout (dynamic): a 64 KB-ALIGNED 64 KB region holds both slice buffers interleaved:
//   byte address = region | code << 8 | buffer << 7 | replica * 4W
// so that the address of a row's table entry is ONE byte-permute of (code word, per-lane base):
// PRMT puts the row's code byte into address bits 8..15.  After the region: per-query sort areas,
// then the survivor queue.
template <int FB, int W, bool WIDE = false>
__global__ void __launch_bounds__(NT, 1) pruned_scan_kernel(const Params p) {
  using C = Cfg<FB, W>;
  constexpr int QT = C::QT, RPT = C::RPT, R = C::R, QSH = C::QSH, RSH = C::RSH;
  constexpr int NV = RPT / 16;  // 16-byte code vectors per thread and quantizer
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t dyn0 = smem_u32(smem_raw);
  const uint32_t region = (dyn0 + 0xffffu) & ~0xffffu;
  unsigned char *region_g = smem_raw + (region - dyn0);
  u64 *sortbuf = reinterpret_cast<u64 *>(region_g + REGION_BYTES);
  uint32_t *surv = reinterpret_cast<uint32_t *>(region_g + REGION_BYTES + QT * SORTN * 8);
  __shared__ int s_cnt[QT];
  __shared__ u64 s_thr[QT];
  __shared__ uint32_t s_bias[W];
  __shared__ int s_nsurv, s_nbefore;
  __shared__ unsigned long long s_stat[3];
  __shared__ i64 s_slot_chunk[SLOTS];
  __shared__ int s_slot_tile[SLOTS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x / p.Bs, j = blockIdx.x % p.Bs;
  if (s >= p.S) return;
  const u64 pol_stream = l2_policy_evict_first();
  const i64 lo = p.from + (i64)s * p.split_len;
  i64 hi = lo + p.split_len;
  if (hi > p.until) hi = p.until;
  if (lo >= hi || j >= p.T) return;
  const i64 origin = lo & ~15LL;
  const i64 n_chunks = (hi - origin + R - 1) / R;
  const int n_my = (p.T - j + p.Bs - 1) / p.Bs;
  const i64 n_items = n_chunks * n_my;
  const int M = p.M, ML = p.ML, k = p.k;
  // survivors may wait in the queue across items only while the CTA stays on one tile (the sort
  // areas are per query of the current tile)
  const bool defer = n_my == 1;

  const int code_id = tid & 255, half = tid >> 8;
  int fill_off[4];  // byte offsets inside the region (buffer 0)
#pragma unroll
  for (int t = 0; t < 4; t++) fill_off[t] = code_id * 256 + ((lane + 4 * half + t) & 7) * 16;
  // buffer bit (0x80) toggles per quantizer
  uint32_t lane_base = region | ((uint32_t)(lane & (C::NREP - 1)) * (4u * W));

  if (tid < QT) s_cnt[tid] = 0;
  if (tid < W) s_bias[tid] = 0;
  if (tid == 0) s_nsurv = 0;
  if (tid < 3) s_stat[tid] = 0;

  int parity = 0;
  {
    const uint4 v = ldg_entry<W>(p.qlut + (((i64)j * ML) * 256 + code_id) * W);
#pragma unroll
    for (int t = 0; t < 4; t++) *reinterpret_cast<uint4 *>(region_g + fill_off[t]) = v;
  }
  uint4 ccur[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    ccur[v] = make_uint4(0, 0, 0, 0);
    const i64 r = origin + (i64)tid * RPT + 16 * v;
    if (r < hi) ccur[v] = ldg_stream_u4(p.codes + (i64)__ldg(p.msel + (i64)j * ML) * p.ps + r, pol_stream);
  }
  __syncthreads();

  int slot = 0;          // items queued since the last drain
  bool need_thr = true;  // thresholds change only with the tile or after a drain

  for (i64 it = 0; it < n_items; ++it) {
    const int t = j + (int)(it % n_my) * p.Bs;
    const i64 chunk0 = origin + (it / n_my) * R;
    const i64 row0 = chunk0 + (i64)tid * RPT;
    const bool has_next = it + 1 < n_items;
    const int tn = has_next ? j + (int)((it + 1) % n_my) * p.Bs : t;
    const i64 row0n = has_next ? origin + ((it + 1) / n_my) * R + (i64)tid * RPT : row0;
    const bool warp_live = chunk0 + (i64)(warp * 32) * RPT < hi;  // any row of this warp in range

    u64 *L0 = p.lists + ((i64)s * p.T * QT + (i64)t * QT) * k;
    // thresholds of the tile's queries (published by the barriers of the quantizer loop)
    if (need_thr && tid < QT) {
      const i64 q = (i64)t * QT + tid;
      const u64 tl = fscan::ldcg_u64(L0 + (i64)tid * k + (k - 1));
      const u64 bt = p.boot_tail[q];
      const u64 thr = tl < bt ? tl : bt;
      s_thr[tid] = thr;
      int T = -1;
      if (q < p.nq) {
        const QParam qp = p.qp[q];
        if (thr == KEY_SENT || qp.inv_delta == 0.0) {
          T = C::FLAG - 1;
        } else {
          const double tau = (double)ord2f((uint32_t)(thr >> 32));
          const double x = (tau * (1.0 + 1.0 / 2048.0) - qp.base) * qp.inv_delta;
          T = x < 0.0 ? -1 : (x >= (double)(C::FLAG - 1) ? C::FLAG - 1 : (int)floor(x));
        }
      }
      // fields of one word are owned by consecutive threads of the first warp: combine by shuffle
      uint32_t b = (uint32_t)(C::FLAG - 1 - T) << (FB * (tid % C::FPW));
#pragma unroll
      for (int o = 1; o < C::FPW; o <<= 1) b |= __shfl_xor_sync((1u << QT) - 1u, b, o);
      if (tid % C::FPW == 0) s_bias[tid / C::FPW] = b;
    }
    need_thr = !defer;
    if (tid == 0) {
      s_slot_chunk[slot] = chunk0;
      s_slot_tile[slot] = t;
      s_nbefore = s_nsurv;  // entries queued by earlier items (no pushes happen during the loop)
    }

    uint32_t acc[RPT][W];
#pragma unroll
    for (int i = 0; i < RPT; i++)
#pragma unroll
      for (int w = 0; w < W; w++) acc[i][w] = 0u;

    for (int m = 0; m < ML; ++m) {
      uint4 cnext[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) cnext[v] = make_uint4(0, 0, 0, 0);
      uint4 lnext = make_uint4(0, 0, 0, 0);
      bool do_fill = false;
      if (m + 1 < ML) {
        const uint8_t *plane = p.codes + (i64)__ldg(p.msel + (i64)t * ML + m + 1) * p.ps;
#pragma unroll
        for (int v = 0; v < NV; v++)
          if (row0 + 16 * v < hi) cnext[v] = ldg_stream_u4(plane + row0 + 16 * v, pol_stream);
        lnext = ldg_entry<W>(p.qlut + (((i64)t * ML + m + 1) * 256 + code_id) * W);
        do_fill = true;
      } else if (has_next) {
        const uint8_t *plane = p.codes + (i64)__ldg(p.msel + (i64)tn * ML) * p.ps;
#pragma unroll
        for (int v = 0; v < NV; v++)
          if (row0n + 16 * v < hi) cnext[v] = ldg_stream_u4(plane + row0n + 16 * v, pol_stream);
        lnext = ldg_entry<W>(p.qlut + (((i64)tn * ML) * 256 + code_id) * W);
        do_fill = true;
      }
      if (warp_live) {
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const uint4 &cv = ccur[i >> 4];
          const uint32_t cw = ((i >> 2) & 3) == 0 ? cv.x : ((i >> 2) & 3) == 1 ? cv.y
                              : ((i >> 2) & 3) == 2 ? cv.z : cv.w;
          // address = lane_base with the row's code byte in bits 8..15
          const uint32_t a = __byte_perm(cw, lane_base, 0x7604u | ((uint32_t)(i & 3) << 4));
          uint32_t v[W];
          lds_words<W>(a, v);
#pragma unroll
          for (int w = 0; w < W; w++) acc[i][w] += v[w];
        }
      }
      if (do_fill) {
        unsigned char *dst = region_g + ((parity ^ 1) << 7);
#pragma unroll
        for (int t4 = 0; t4 < 4; t4++) *reinterpret_cast<uint4 *>(dst + fill_off[t4]) = lnext;
      }
      __syncthreads();
      parity ^= 1;
      lane_base ^= 0x80u;
#pragma unroll
      for (int v = 0; v < NV; v++) ccur[v] = cnext[v];
    }

    // ---- flag test -------------------------------------------------------------------------
    const int n_before = s_nbefore;  // published by the barriers of the quantizer loop
    const int vlo = lo > row0 ? (int)(lo - row0 > RPT ? RPT : lo - row0) : 0;
    const int vhi = hi - row0 >= RPT ? RPT : (hi > row0 ? (int)(hi - row0) : 0);
    uint32_t bias[W];
#pragma unroll
    for (int w = 0; w < W; w++) bias[w] = s_bias[w];
    unsigned long long rm = 0;  // rows of this thread with at least one unflagged field
#pragma unroll
    for (int i = 0; i < RPT; i++) {
      uint32_t x = 0xffffffffu;
#pragma unroll
      for (int w = 0; w < W; w++) {
        acc[i][w] += bias[w];
        x &= acc[i][w];
      }
      if ((x & C::FLAGMASK) != C::FLAGMASK && i >= vlo && i < vhi) rm |= 1ull << i;
    }

    // survivors of the rows selected by `rowmask` -> queue; entry = slot | local row | field
    auto push = [&](unsigned long long rowmask, int slot_) {
      const unsigned long long todo = rm & rowmask;
      if (todo == 0) return;
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        if ((todo >> i) & 1ull) {
#pragma unroll
          for (int w = 0; w < W; w++) {
            uint32_t live = ~acc[i][w] & C::FLAGMASK;  // unflagged fields of this word
            while (live) {
              const int bit = __ffs(live) - 1;
              live &= live - 1;
              const int f = w * C::FPW + bit / FB;
              const int pos = atomicAdd(&s_nsurv, 1);
              if (pos < C::QCAP)
                surv[pos] = ((uint32_t)slot_ << (RSH + QSH)) | ((uint32_t)(tid * RPT + i) << QSH) |
                            (uint32_t)f;
            }
          }
        }
      }
    };
    // exact re-evaluation of queue entries [0, n): one THREAD per survivor, NT survivors per round.
    // The distance is summed exactly as the reference does -- (((0 + t_0) + t_1) + ...) in fp32, m
    // ascending -- eight table reads in flight at a time, and the walk stops as soon as the exact
    // prefix plus the minima of the quantizers still to come exceeds the query's threshold (with
    // the same 2^-11 margin as the integer bound: the finished sum would compare greater).  Most
    // survivors of a subset bound stop after a third of the quantizers.  Codes come from the
    // row-major copy (one or two sectors per survivor) when the index has one.  Keys that beat
    // their query's threshold go to the query's sort area; a key that finds the area full (CAPQ)
    // waits in its register for the next merge round and is re-tested against the new threshold.
    auto drain = [&](int n) {
      for (int b0s = 0; b0s < n; b0s += NT) {
        bool pending = false;
        u64 key = KEY_SENT, floor_q = 0ull;
        int q = 0;
        if (b0s + tid < n) {
          const uint32_t code = surv[b0s + tid];
          q = (int)(code & (uint32_t)(QT - 1));
          const int sl = (int)(code >> (RSH + QSH));
          const i64 row = s_slot_chunk[sl] + (i64)((code >> QSH) & (uint32_t)(R - 1));
          const int ts = s_slot_tile[sl];
          const float *lut = reinterpret_cast<const float *>(
                                 p.lutI + ((i64)ts * (QT / 4) + (q >> 2)) * M * 256) + (q & 3);
          const float *suf = p.sufmin + ((i64)ts * QT + q) * (M + 1);
          const u64 pol_keep = l2_policy_evict_last();
          floor_q = p.floor ? p.floor[(i64)ts * QT + q] : 0ull;
          const u64 thr = s_thr[q];
          double tau_hi = 1.0e300;   // no early exit without a finite threshold
          if (thr != KEY_SENT) {
            const float tau = ord2f((uint32_t)(thr >> 32));
            if (tau == tau && tau < 3.0e38f) tau_hi = (double)tau * (1.0 + 1.0 / 2048.0);
          }
          float d = 0.0f;
          bool out = false;
          if constexpr (WIDE) {
            // 16-bit ids from the row-major copy (8 per 16-byte load), exact tables [M][K] of the query
            const float *lw = p.lutW + ((i64)ts * QT + q) * M * p.K;
            for (int mb = 0; mb < M && !out; mb += 8) {
              const uint4 c4 = *reinterpret_cast<const uint4 *>(p.rowcodes16 + row * p.rcs16 + mb);
              const uint32_t cw8[4] = {c4.x, c4.y, c4.z, c4.w};
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; u++) {
                const int c = (int)((cw8[u >> 1] >> (16 * (u & 1))) & 0xffffu);
                v[u] = mb + u < M ? ldg_keep_f32(lw + (i64)(mb + u) * p.K + c, pol_keep) : 0.0f;
              }
#pragma unroll
              for (int u = 0; u < 8; u++)
                if (mb + u < M) d = __fadd_rn(d, v[u]);
              const int done = mb + 8 < M ? mb + 8 : M;
              if ((double)d + (double)suf[done] > tau_hi) out = true;
            }
          } else {
          for (int m0 = 0; m0 < M && !out; m0 += 16) {
            uint32_t cw[4];
            if (p.rowcodes) {
              const uint4 c4 = *reinterpret_cast<const uint4 *>(p.rowcodes + row * p.rcs + m0);
              cw[0] = c4.x; cw[1] = c4.y; cw[2] = c4.z; cw[3] = c4.w;
            } else {
#pragma unroll
              for (int u = 0; u < 4; u++) cw[u] = 0u;
#pragma unroll
              for (int u = 0; u < 16; u++)
                if (m0 + u < M)
                  cw[u >> 2] |= (uint32_t)p.codes[(i64)(m0 + u) * p.ps + row] << (8 * (u & 3));
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int mb = m0 + 8 * h;
              if (mb < M && !out) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                  const int c = (int)((cw[2 * h + (u >> 2)] >> (8 * (u & 3))) & 0xffu);
                  v[u] = mb + u < M ? ldg_keep_f32(lut + ((mb + u) * 256 + c) * 4, pol_keep) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; u++)
                  if (mb + u < M) d = __fadd_rn(d, v[u]);
                const int done = mb + 8 < M ? mb + 8 : M;
                if ((double)d + (double)suf[done] > tau_hi) out = true;
              }
            }
          }
          }
          if (!out) {
            key = make_key(d, (uint32_t)row);
            pending = key > floor_q;   // rows an earlier pass of a k-chunked scan already reported stay out
          }
        }
        for (;;) {
          if (pending) {
            if (key < s_thr[q]) {
              const int pos = atomicAdd(&s_cnt[q], 1);
              if (pos < CAPQ) {
                sortbuf[q * SORTN + k + pos] = key;
                pending = false;
              }
            } else {
              pending = false;
            }
          }
          const int more = __syncthreads_or(pending ? 1 : 0);
          if (warp < QT) {
            const int nc = s_cnt[warp] < CAPQ ? s_cnt[warp] : CAPQ;
            if (nc > 0) {
              u64 *L = L0 + (i64)warp * k;
              fscan::warp_merge(sortbuf + warp * SORTN, L, k, nc, lane);
              if (lane == 0) {
                const u64 tl = L[k - 1];
                const u64 bt = p.boot_tail[(i64)t * QT + warp];
                s_thr[warp] = tl < bt ? tl : bt;
                atomicAdd(&s_stat[1], (unsigned long long)nc);
              }
            }
            __syncwarp();
            if (lane == 0) s_cnt[warp] = 0;
          }
          __syncthreads();
          if (!more) break;
        }
      }
      if (tid == 0) s_stat[0] += (unsigned long long)n;
    };

    push(~0ull, slot);
    __syncthreads();
    const int ns = s_nsurv;
    if (ns > C::QCAP) {
      // queue overflow (adversarial orders only): finish the entries of earlier items, then redo
      // this item one row index per round so that a round never exceeds the queue
      drain(n_before);
      for (int r = 0; r < RPT; r++) {
        __syncthreads();
        if (tid == 0) s_nsurv = 0;
        __syncthreads();
        push(1ull << r, slot);
        __syncthreads();
        drain(s_nsurv);
      }
      __syncthreads();
      if (tid == 0) {
        s_nsurv = 0;
        s_stat[2] += 1;
        s_slot_chunk[0] = chunk0;
        s_slot_tile[0] = t;
      }
      slot = 0;
      need_thr = true;
    } else if (ns > 0 && (!defer || !has_next || ns >= C::DRAIN_AT || slot == SLOTS - 1)) {
      drain(ns);
      if (tid == 0) s_nsurv = 0;
      slot = 0;
      need_thr = true;
    } else if (ns > 0) {
      slot++;
    }
    // the quantizer loop of the next item has >= 1 barrier before its n_before read and pushes
  }
  __syncthreads();
  if (tid < 3 && p.stats && s_stat[tid]) atomicAdd(p.stats + tid, s_stat[tid]);
}

}  // namespace pscan
}  // namespace gulon
