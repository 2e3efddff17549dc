// tcassign.cuh -- KMeans.assign / ProductQuantizer.encode on the 5th-generation tensor cores.
//
// Reference: KMeans.assign, G/KMeans.scala:24-55,70-98: for every row the argmin over k of
//   s_k = fl(off_k - 2 * dot_seq(x, c_k)),  strict '<' in ascending k (lowest index on ties here).
// The exact CUDA-core kernel (kmeans.cuh) spends 2*K*dim unfusable FMUL/FADD per row and window.
// This kernel computes ALL K scores of 128 rows approximately with tcgen05.mma (norms-plus-cross-term
// as one dense contraction), uses them only to discard centroids that provably cannot be the
// argmin, and re-evaluates the few remaining candidates with the reference's literal arithmetic:
//
//   A[row][.]  = [ xh | xh | xl | 1 1 1 | 0.. ]            (bf16, K-major, 48 slots)
//   B[k][.]    = [ bh | bl | bh | o1 o2 o3 | 0.. ]          b = -2 c_k = bh + bl (+2^-16), O_k = o1+o2+o3
//   D = A B^T  = off_k - 2 x.c_k  up to  E <= 2^-13 |x| |c_k| + 2^-16 |off_k|   (fp32 accumulate in TMEM)
//
// Epilogue ("sweep"): thread = row = TMEM lane; the 256 columns are reduced to 32 chunk minima of
// 8 columns with FMNMX; every chunk whose minimum is within 2E of the row minimum is a candidate
// (this always includes the chunk of the true argmin and of every exact tie).  Candidate (row,
// chunk) pairs go through a shared-memory queue to the "exact" warps, which evaluate the 8
// centroids of the chunk with __fmul_rn/__fadd_rn in the reference's order and keep the
// (score, index)-lexicographic minimum per row.  On clustered data ~1.05 chunks per row survive,
// i.e. ~8.4 exact evaluations per row instead of 256.
//
// Warp roles (416 threads, 1 CTA/SM, persistent over work units):
//   warps 0-3   producers: load the rows' windows (fp32), split to bf16, write the A tile in the
//               canonical no-swizzle K-major UMMA layout, the fp32 rows and the 2E windows
//   warps 4-7   sweep: tcgen05.ld the accumulator (warp w owns TMEM lanes 32*(w%4)..), queue pairs
//   warps 8-11  exact evaluation of queued pairs, output
//   warp 12     TMEM allocation, TMA bulk loads of the per-window operand blobs, tcgen05.mma issue
// Pipelines: A tiles x2 (a_full/a_empty), TMEM accumulators x2 (t_full/t_empty), tile slots x3
// (s_ready/q_full/slot_free), all mbarriers.  A work unit = (group of <= 3 windows, 8192 rows): the
// windows of a group are adjacent in the row, so a CTA reads up to 120 contiguous bytes per row.
#pragma once
#include <cuda_bf16.h>
#include <float.h>

#include "common.cuh"

namespace gulon {
namespace tca {

constexpr int NT = 416;
constexpr int TM = 128;          // rows per tile (UMMA M)
constexpr int TN = 256;          // centroids per tile (UMMA N)
constexpr int KP = 48;           // padded contraction depth (3 x K16)
constexpr int GRP = 3;           // windows per work unit
constexpr int UNIT_ROWS = 8192;  // rows per work unit
constexpr int NS = 3;            // tile slots in flight (fp32 rows, queue, results)
constexpr int QCAP = 1024;       // queued (row, chunk) pairs per tile before the overflow path
constexpr int A_BYTES = TM * KP * 2;         // 12288
constexpr int B_BYTES = TN * KP * 2;         // 24576
constexpr int CB_LD = 17;                    // floats per centroid row in shared memory (odd: no bank conflicts)
constexpr int CB_BYTES = TN * CB_LD * 4;     // 17408
constexpr int OFF_BYTES = TN * 4;            // 1024
constexpr int META_BYTES = 16;               // cmax, omax, flags, pad
constexpr int BLOB_BYTES = B_BYTES + CB_BYTES + OFF_BYTES + META_BYTES;  // per window, 16 B multiple
constexpr int XS_LD = 17;                    // floats per fp32 row in a slot
constexpr int SLOT_BYTES = TM * XS_LD * 4 + TM * 4 + QCAP * 2 + TM * 8 + 16;  // 8-byte multiple
constexpr int SMEM_BYTES = GRP * BLOB_BYTES + 2 * A_BYTES + NS * SLOT_BYTES + 256 + 1024;
constexpr u64 RES_INIT = ((u64)0xFF7FFFFFu << 32);  // (ord(FLT_MAX), k = 0): "nothing accepted yet"

static_assert(BLOB_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

struct Params {
  const float *X;
  i64 N, ld;
  const unsigned char *blobs;   // [M][BLOB_BYTES] prepared operands (tc_prep_kernel)
  const int32_t *subs;          // windows to process (device list), nsub entries
  const int32_t *from;          // [M] first column of every window
  int nsub, K;
  void *out;                    // [M][out_stride] uint8 or int32
  i64 out_stride;
  unsigned long long *stats;    // [0] exact evaluations (pairs), [1] overflow tiles
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sa(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t *b) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(sa(b)) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(sa(b)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(sa(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(dst)),
               "l"(src), "r"(bytes), "r"(sa(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (base lane + t).
// The registers are valid only after tc_ld_wait().
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// The registers are operands of the wait so that no use of them can be scheduled above it.
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// minima of the 4 chunks of 8 columns held in r[32]
__device__ __forceinline__ void chunk_mins(const uint32_t (&r)[32], float *cm) {
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const float m01 = fminf(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]));
    const float m23 = fminf(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
    const float m45 = fminf(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
    const float m67 = fminf(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
    cm[c] = fminf(fminf(m01, m23), fminf(m45, m67));
  }
}
// shared-memory matrix descriptor: K-major, no swizzle; 8x(16 B) core matrices, LBO between the two
// K chunks of an instruction, SBO between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ u64 smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (u64)((addr >> 4) & 0x3FFFu) | ((u64)((lbo >> 4) & 0x3FFFu) << 16) | ((u64)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, K-major, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint16_t bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float bf16_val(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// ---- operand preparation: one blob per window ------------------------------------------------------
// grid (nsub), block 256 (thread = centroid).  cb [M][K][dmax] fp32, off [M][K]; blob of window m at m * BLOB_BYTES.
__global__ void __launch_bounds__(256) tc_prep_kernel(const float *__restrict__ cb, const float *__restrict__ off,
                                                      const int32_t *__restrict__ subs,
                                                      const int32_t *__restrict__ dims, int K, int dmax,
                                                      unsigned char *__restrict__ blobs) {
  const int m = subs[blockIdx.x], k = threadIdx.x, dim = dims[m];
  unsigned char *blob = blobs + (size_t)m * BLOB_BYTES;
  float *cbs = reinterpret_cast<float *>(blob + B_BYTES);
  float *offs = reinterpret_cast<float *>(blob + B_BYTES + CB_BYTES);
  float *meta = reinterpret_cast<float *>(blob + B_BYTES + CB_BYTES + OFF_BYTES);
  __shared__ float s_c[256], s_o[256];
  __shared__ int s_bad;
  if (k == 0) s_bad = 0;
  __syncthreads();
  uint16_t slots[KP];
#pragma unroll
  for (int i = 0; i < KP; i++) slots[i] = 0;
  float nrm2 = 0.0f, o = 0.0f;
  bool bad = false;
  if (k < K && 3 * dim + 3 <= KP) {
    const float *c = cb + ((size_t)m * K + k) * dmax;
    o = off[(size_t)m * K + k];
    for (int j = 0; j < dim; j++) {
      const float cj = c[j];
      const float b = -2.0f * cj;
      const uint16_t bh = bf16_bits(b);
      const uint16_t bl = bf16_bits(b - bf16_val(bh));
      slots[j] = bh;
      slots[dim + j] = bl;
      slots[2 * dim + j] = bh;
      nrm2 += cj * cj;
      cbs[k * CB_LD + j] = cj;
      if (!(fabsf(cj) < 1e18f)) bad = true;
    }
    for (int j = dim; j < CB_LD; j++) cbs[k * CB_LD + j] = 0.0f;
    const uint16_t o1 = bf16_bits(o);
    const float r1 = o - bf16_val(o1);
    const uint16_t o2 = bf16_bits(r1);
    const uint16_t o3 = bf16_bits(r1 - bf16_val(o2));
    slots[3 * dim] = o1;
    slots[3 * dim + 1] = o2;
    slots[3 * dim + 2] = o3;
    if (!(fabsf(o) < 1e36f)) bad = true;
  } else {
    // padding centroid: approximate score 3e38, never within 2E of a real minimum
    for (int j = 0; j < CB_LD; j++) cbs[k * CB_LD + j] = 0.0f;
    if (3 * dim < KP) slots[3 * dim] = bf16_bits(3.0e38f);  // times the row's 1.0 slot
    o = 0.0f;
  }
  offs[k] = o;
  // canonical layout: element (k, slot) at (k/8)*SBO + (slot/8)*128 + (k%8)*16 + (slot%8)*2
#pragma unroll
  for (int kc = 0; kc < KP / 8; kc++) {
    uint4 v;
    v.x = slots[kc * 8 + 0] | ((uint32_t)slots[kc * 8 + 1] << 16);
    v.y = slots[kc * 8 + 2] | ((uint32_t)slots[kc * 8 + 3] << 16);
    v.z = slots[kc * 8 + 4] | ((uint32_t)slots[kc * 8 + 5] << 16);
    v.w = slots[kc * 8 + 6] | ((uint32_t)slots[kc * 8 + 7] << 16);
    *reinterpret_cast<uint4 *>(blob + (k >> 3) * (KP / 8) * 128 + kc * 128 + (k & 7) * 16) = v;
  }
  s_c[k] = k < K ? sqrtf(nrm2) : 0.0f;
  s_o[k] = k < K ? fabsf(o) : 0.0f;
  if (bad) s_bad = 1;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (k < st) {
      s_c[k] = fmaxf(s_c[k], s_c[k + st]);
      s_o[k] = fmaxf(s_o[k], s_o[k + st]);
    }
    __syncthreads();
  }
  if (k == 0) {
    meta[0] = s_c[0] * 1.001f;  // max |c_k|
    meta[1] = s_o[0] * 1.001f;  // max |off_k|
    meta[2] = (s_bad || 3 * dim + 3 > KP) ? 1.0f : 0.0f;  // non-finite operands: every chunk is a candidate
    meta[3] = 0.0f;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------
template <int DIM, typename OutT>
__global__ void __launch_bounds__(NT, 1) tc_assign_kernel(const Params p) {
  static_assert(3 * DIM + 3 <= KP, "window too wide for the 48-slot contraction");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *blob_s = smem;                              // GRP blobs
  unsigned char *a_s = smem + GRP * BLOB_BYTES;              // 2 A tiles
  unsigned char *slot_s = a_s + 2 * A_BYTES;                 // NS slots
  uint64_t *bars = reinterpret_cast<uint64_t *>(slot_s + NS * SLOT_BYTES);
  uint64_t *b_full = bars;            // 1
  uint64_t *a_full = bars + 1;        // 2
  uint64_t *a_empty = bars + 3;       // 2
  uint64_t *t_full = bars + 5;        // 2
  uint64_t *t_empty = bars + 7;       // 2
  uint64_t *s_ready = bars + 9;       // NS
  uint64_t *q_full = bars + 12;       // NS
  uint64_t *slot_free = bars + 15;    // NS
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_groups = (p.nsub + GRP - 1) / GRP;
  const i64 n_ranges = (p.N + UNIT_ROWS - 1) / UNIT_ROWS;
  const i64 n_units = n_ranges * n_groups;

  auto slot_xs = [&](int s) { return reinterpret_cast<float *>(slot_s + s * SLOT_BYTES); };
  auto slot_win = [&](int s) { return reinterpret_cast<float *>(slot_s + s * SLOT_BYTES + TM * XS_LD * 4); };
  auto slot_q = [&](int s) { return reinterpret_cast<uint16_t *>(slot_s + s * SLOT_BYTES + TM * XS_LD * 4 + TM * 4); };
  auto slot_res = [&](int s) {
    return reinterpret_cast<u64 *>(slot_s + s * SLOT_BYTES + TM * XS_LD * 4 + TM * 4 + QCAP * 2);
  };
  auto slot_cnt = [&](int s) {
    return reinterpret_cast<int *>(slot_s + s * SLOT_BYTES + TM * XS_LD * 4 + TM * 4 + QCAP * 2 + TM * 8);
  };

  if (tid == 0) {
    mb_init(b_full, 1);
    for (int i = 0; i < 2; i++) {
      mb_init(a_full + i, 4);
      mb_init(a_empty + i, 1);
      mb_init(t_full + i, 1);
      mb_init(t_empty + i, 4);
    }
    for (int i = 0; i < NS; i++) {
      mb_init(s_ready + i, 4);
      mb_init(q_full + i, 4);
      mb_init(slot_free + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int s = 0; s < NS; s++) {
    if (tid < TM) slot_res(s)[tid] = RES_INIT;
    if (tid < 2) slot_cnt(s)[tid] = 0;
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  i64 tile = 0;        // tiles processed by this CTA so far (drives every ring index / phase)
  uint32_t b_phase = 0;

  for (i64 unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int g = (int)(unit % n_groups);
    const i64 r_begin = (unit / n_groups) * UNIT_ROWS;
    const i64 r_end = r_begin + UNIT_ROWS < p.N ? r_begin + UNIT_ROWS : p.N;
    const int nw = p.nsub - g * GRP < GRP ? p.nsub - g * GRP : GRP;   // windows in this group
    const int n_blocks = (int)((r_end - r_begin + TM - 1) / TM);
    const int n_tiles = n_blocks * nw;

    // operands of the group's windows -> shared memory (TMA bulk copies, one mbarrier)
    if (warp == 12 && lane == 0) {
      mb_expect_tx(b_full, (uint32_t)nw * BLOB_BYTES);
      for (int w = 0; w < nw; w++)
        bulk_g2s(blob_s + w * BLOB_BYTES, p.blobs + (size_t)p.subs[g * GRP + w] * BLOB_BYTES, BLOB_BYTES, b_full);
    }
    mb_wait(b_full, b_phase);
    b_phase ^= 1u;

    if (warp < 4) {
      // ================= producers: thread = row of the tile =================
      const int r = tid;
      float xn[DIM];
      auto load_row = [&](int t) {
        const int w = t % nw;
        const i64 row = r_begin + (i64)(t / nw) * TM + r;
        const int fr = p.from[p.subs[g * GRP + w]];
        if (row < r_end) {
          const float *src = p.X + row * p.ld + fr;
#pragma unroll
          for (int j = 0; j < DIM; j++) xn[j] = __ldg(src + j);
        } else {
#pragma unroll
          for (int j = 0; j < DIM; j++) xn[j] = 0.0f;
        }
      };
      if (n_tiles > 0) load_row(0);
      for (int t = 0; t < n_tiles; t++) {
        const i64 ti = tile + t;
        const int ab = (int)(ti & 1), sl = (int)(ti % NS);
        float x[DIM];
#pragma unroll
        for (int j = 0; j < DIM; j++) x[j] = xn[j];
        if (t + 1 < n_tiles) load_row(t + 1);
        const float *meta = reinterpret_cast<const float *>(blob_s + (t % nw) * BLOB_BYTES + B_BYTES + CB_BYTES + OFF_BYTES);
        // split and pack
        uint16_t slots[KP];
#pragma unroll
        for (int i = 0; i < KP; i++) slots[i] = 0;
        float n2 = 0.0f;
        bool bad = meta[2] != 0.0f;
#pragma unroll
        for (int j = 0; j < DIM; j++) {
          const uint16_t xh = bf16_bits(x[j]);
          const uint16_t xl = bf16_bits(x[j] - bf16_val(xh));
          slots[j] = xh;
          slots[DIM + j] = xh;
          slots[2 * DIM + j] = xl;
          n2 += x[j] * x[j];
          if (!(fabsf(x[j]) < 1e18f)) bad = true;
        }
        slots[3 * DIM] = slots[3 * DIM + 1] = slots[3 * DIM + 2] = 0x3F80;  // bf16 1.0
        // 2E window: E = 2^-12 |x| max|c| + 2^-15 max|off| (twice the derived bound, see header)
        const float xnorm = sqrtf(n2) * 1.001f;
        float win = 2.0f * (xnorm * meta[0] * (1.0f / 4096.0f) + meta[1] * (1.0f / 32768.0f)) + 1e-30f;
        if (bad) win = __int_as_float(0x7f800000);  // +inf: every chunk is a candidate
        mb_wait(slot_free + sl, (uint32_t)(((ti / NS) & 1) ^ 1));
        mb_wait(a_empty + ab, (uint32_t)(((ti >> 1) & 1) ^ 1));
        unsigned char *A = a_s + ab * A_BYTES + (r >> 3) * (KP / 8) * 128 + (r & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < KP / 8; kc++) {
          uint4 v;
          v.x = slots[kc * 8 + 0] | ((uint32_t)slots[kc * 8 + 1] << 16);
          v.y = slots[kc * 8 + 2] | ((uint32_t)slots[kc * 8 + 3] << 16);
          v.z = slots[kc * 8 + 4] | ((uint32_t)slots[kc * 8 + 5] << 16);
          v.w = slots[kc * 8 + 6] | ((uint32_t)slots[kc * 8 + 7] << 16);
          *reinterpret_cast<uint4 *>(A + kc * 128) = v;
        }
        float *xs = slot_xs(sl) + r * XS_LD;
#pragma unroll
        for (int j = 0; j < DIM; j++) xs[j] = x[j];
        slot_win(sl)[r] = win;
        fence_async_smem();  // A is read by the tensor core through the async proxy
        __syncwarp();
        if (lane == 0) {
          mb_arrive(a_full + ab);
          mb_arrive(s_ready + sl);
        }
      }
    } else if (warp < 8) {
      // ================= sweep: thread = row = TMEM lane =================
      const int q4 = warp & 3, r = q4 * 32 + lane;
      for (int t = 0; t < n_tiles; t++) {
        const i64 ti = tile + t;
        const int acc = (int)(ti & 1), sl = (int)(ti % NS);
        mb_wait(t_full + acc, (uint32_t)((ti >> 1) & 1));
        tc_fence_after();
        float cmin[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * TN);
        {
          // two register buffers: the next 32 columns are in flight while the previous are reduced
          uint32_t v0[32], v1[32];
          tc_ld32_issue(taddr, v0);
          tc_ld_wait(v0);
#pragma unroll
          for (int cb = 0; cb < 8; cb += 2) {
            tc_ld32_issue(taddr + (cb + 1) * 32, v1);
            chunk_mins(v0, cmin + cb * 4);
            tc_ld_wait(v1);
            if (cb + 2 < 8) tc_ld32_issue(taddr + (cb + 2) * 32, v0);
            chunk_mins(v1, cmin + (cb + 1) * 4);
            if (cb + 2 < 8) tc_ld_wait(v0);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mb_arrive(t_empty + acc);  // accumulator drained
        float rmin = cmin[0];
#pragma unroll
        for (int c = 1; c < 32; c++) rmin = fminf(rmin, cmin[c]);
        mb_wait(s_ready + sl, (uint32_t)((ti / NS) & 1));
        const float thr = rmin + slot_win(sl)[r];
        const bool all = !(thr < __int_as_float(0x7f800000));  // +inf or NaN: take everything
        const i64 row = r_begin + (i64)(t / nw) * TM + r;
        if (row < r_end) {
          int *cnt = slot_cnt(sl);
          uint16_t *q = slot_q(sl);
#pragma unroll
          for (int c = 0; c < 32; c++) {
            if (all || cmin[c] <= thr) {
              const int pos = atomicAdd(cnt, 1);
              if (pos < QCAP) q[pos] = (uint16_t)((r << 5) | c);
              else cnt[1] = 1;  // overflow: the exact warps take every (row, chunk) of the tile
            }
          }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(q_full + sl);
      }
    } else if (warp < 12) {
      // ================= exact evaluation of candidate chunks =================
      const int xt = tid - 256;        // 0..127
      const int pj = xt & 7, pp = xt >> 3;  // centroid inside the chunk, pair lane (16 pairs at a time)
      unsigned long long n_eval = 0, n_over = 0;
      for (int t = 0; t < n_tiles; t++) {
        const i64 ti = tile + t;
        const int sl = (int)(ti % NS), w = t % nw;
        mb_wait(q_full + sl, (uint32_t)((ti / NS) & 1));
        const float *cbs = reinterpret_cast<const float *>(blob_s + w * BLOB_BYTES + B_BYTES);
        const float *offs = reinterpret_cast<const float *>(blob_s + w * BLOB_BYTES + B_BYTES + CB_BYTES);
        const float *xs = slot_xs(sl);
        const uint16_t *q = slot_q(sl);
        u64 *res = slot_res(sl);
        int *cnt = slot_cnt(sl);
        const bool over = cnt[1] != 0;
        const int n = over ? TM * 32 : cnt[0];
        for (int p0 = 0; p0 < n; p0 += 16) {
          const int pi = p0 + pp;
          u64 key = ~0ull;
          int rr = 0;
          if (pi < n) {
            const int code = over ? pi : (int)q[pi];
            rr = code >> 5;
            const int k = ((code & 31) << 3) + pj;
            if (k < p.K) {
              const float *xr = xs + rr * XS_LD;
              const float *c = cbs + k * CB_LD;
              float d = 0.0f;
#pragma unroll
              for (int j = 0; j < DIM; j++) d = __fadd_rn(d, __fmul_rn(xr[j], c[j]));
              float s = __fsub_rn(offs[k], __fmul_rn(2.0f, d));
              if (s < FLT_MAX) {            // NaN and >= Float.MaxValue are never accepted
                s = s + 0.0f;               // -0.0 == +0.0 for the reference's '<'
                key = ((u64)f2ord(s) << 32) | (u64)k;
              }
            }
          }
          // minimum over the 8 centroids of the chunk (8 consecutive lanes)
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            const u64 other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
          }
          if (pj == 0 && pi < n && key != ~0ull) atomicMin(res + rr, key);
        }
        n_eval += (unsigned long long)n;
        n_over += over ? 1 : 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // all pairs of the tile are folded into res
        const i64 row = r_begin + (i64)(t / nw) * TM + xt;
        const u64 rk = res[xt];
        res[xt] = RES_INIT;
        if (row < r_end) {
          OutT *out = reinterpret_cast<OutT *>(p.out) + (i64)p.subs[g * GRP + w] * p.out_stride + row;
          *out = (OutT)(uint32_t)(rk & 0xffffffffu);
        }
        if (xt == 0) {
          cnt[0] = 0;
          cnt[1] = 0;
        }
        __syncwarp();
        if (lane == 0) mb_arrive(slot_free + sl);
      }
      if (p.stats && xt == 0) {
        atomicAdd(p.stats, n_eval);
        if (n_over) atomicAdd(p.stats + 1, n_over);
      }
    } else {
      // ================= MMA issue (one thread) =================
      if (lane == 0) {
        for (int t = 0; t < n_tiles; t++) {
          const i64 ti = tile + t;
          const int ab = (int)(ti & 1), acc = (int)(ti & 1), w = t % nw;
          mb_wait(a_full + ab, (uint32_t)((ti >> 1) & 1));
          mb_wait(t_empty + acc, (uint32_t)(((ti >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t a_addr = sa(a_s + ab * A_BYTES), b_addr = sa(blob_s + w * BLOB_BYTES);
#pragma unroll
          for (int ks = 0; ks < KP / 16; ks++) {
            const u64 ad = smem_desc(a_addr + ks * 256, 128, (KP / 8) * 128);
            const u64 bd = smem_desc(b_addr + ks * 256, 128, (KP / 8) * 128);
            tc_mma_bf16(tmem_base + (uint32_t)(acc * TN), ad, bd, IDESC, ks > 0 ? 1u : 0u);
          }
          tc_commit(a_empty + ab);   // A tile consumed
          tc_commit(t_full + acc);   // accumulator ready
        }
      }
      __syncwarp();
    }
    tile += n_tiles;
    // every role is done with the group's operands before the next unit overwrites them
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace tca
}  // namespace gulon
