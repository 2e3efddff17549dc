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
//   B[k][.]    = [ bh | bl | bh | o1 o2 o3 | 0.. ]          b = -2 c_k ~ bh + bl, off_k = o1 + o2 + o3
//   D = A B^T  = off_k - 2 x.c_k  up to  E <= 2^-13.4 |x| |c_k| + 2^-17 |off_k|   (fp32 accumulate in TMEM)
//
// (xh = x truncated to bf16, xl = bf16(x - xh); the dropped terms are xl'.bl, the rounding of xl and
// the third piece of b: 4.5 * 2^-16 |x||c| in total, plus the accumulation error of the tensor core.)
//
// Data path, per CTA (persistent over work units; unit = (group of <= 3 adjacent windows, row range)):
//   TMA     X[128 rows][32 floats] boxes (128-byte swizzle) straight from the row-major matrix into a
//           ring of raw tiles; one box serves all windows of the group.  The per-window operand
//           blobs (B tile, fp32 centroids, offsets) arrive by bulk copies once per unit.
//   warps 0-3   "split": thread = row; read the row's window from the raw tile, split to bf16, write
//               the A tile in the canonical no-swizzle K-major UMMA layout.
//   warp 12     one thread issues tcgen05.mma (M=128, N=256, 3 x K16) into one of two TMEM accumulators.
//   warps 4-11  two groups of 4 warps, one per accumulator; thread = row = TMEM lane.  "Sweep": the
//               256 scores are reduced with FMNMX3 to 32 minima of 8-centroid chunks; a chunk is a
//               candidate if its minimum is within 2E of the row minimum (this always includes the
//               chunk of the true argmin and of every exact tie).  "Exact": the same thread evaluates
//               the 8 centroids of every candidate chunk with __fmul_rn/__fadd_rn in the reference's
//               order (ascending k, strict '<') and writes the index.  On clustered data ~1.01 chunks
//               per row survive, i.e. ~8 exact evaluations per row instead of 256.
//   warp 13     one thread issues the TMA loads.
// All hand-offs are mbarriers; nothing but the setup / teardown __syncthreads is CTA-wide.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <float.h>

#include "common.cuh"

namespace gulon {
namespace tca {

constexpr int NWG = 4;            // sweep/exact warp groups (4 warps each); tile t belongs to group t % NWG
constexpr int NT = 32 * (4 + 4 * NWG + 2);
constexpr int TM = 128;          // rows per tile (UMMA M)
constexpr int TN = 256;          // centroids per tile (UMMA N)
constexpr int KP = 48;           // padded contraction depth (3 x K16)
constexpr int GRP_MAX = 3;       // windows per work unit (stride of the group table)
constexpr int BOX_COLS = 32;     // floats per row of a raw tile = the 128-byte swizzle span
constexpr int NRAW = 3;          // raw tiles in flight
constexpr int RAW_BYTES = TM * BOX_COLS * 4;  // 16384
constexpr int A_BYTES = TM * KP * 2;          // 12288
constexpr int B_BYTES = TN * KP * 2;          // 24576
constexpr int CH = 8;                         // centroids per chunk
constexpr int NCH = TN / CH;                  // 32 chunks = one mask word
constexpr int CB_LD = 20;                     // floats per centroid (odd number of 16-byte units)
constexpr int CHUNK_FLOATS = CH * CB_LD + 4;  // + 16 bytes: chunks start in different bank groups
constexpr int CB_BYTES = NCH * CHUNK_FLOATS * 4;  // 20992
constexpr int META_BYTES = 16;                    // max |c|, max |off|, flags, pad
constexpr int BLOB_BYTES = B_BYTES + CB_BYTES + META_BYTES;  // per window
// a centroid row holds its dim coordinates and, in the last float of the 16-byte unit after them,
// its offset |c|^2 (+inf for padding centroids), so that one run of LDS.128 fetches both
__host__ __device__ constexpr int off_slot(int dim) { return ((dim + 1 + 3) & ~3) - 1; }
static_assert(off_slot(15) < CB_LD, "offset slot outside the centroid row");
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = NRAW * RAW_BYTES + 2 * A_BYTES + GRP_MAX * BLOB_BYTES + BAR_BYTES;
// Fused Lloyd pass (assignment + fixed-point cluster sums + changed count, one read of the matrix):
// units hold at most FUSE_GRP windows, and the room of the third operand blob holds the shared-memory
// accumulators of kupdate.cuh: per window LO[256][DIM], HI[256][DIM], CNT[256], BAD[256] 32-bit words.
constexpr int FUSE_GRP = 2;
constexpr int FUSE_KMAX = 256;
constexpr int FUSE_FIX_BITS = 28;
__host__ __device__ constexpr int fuse_wstride(int dim) { return 2 * FUSE_KMAX * dim + 2 * FUSE_KMAX; }
__host__ __device__ constexpr int fuse_smem_bytes(int dim) {
  return NRAW * RAW_BYTES + 2 * A_BYTES + FUSE_GRP * BLOB_BYTES + FUSE_GRP * fuse_wstride(dim) * 4 + BAR_BYTES;
}
static_assert(fuse_smem_bytes(15) <= 232448, "the fused pass must fit the 227 KB of an sm_100 CTA");

static_assert(BLOB_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
static_assert(NCH == 32, "the candidate mask is one 32-bit word");


struct Params {
  i64 N;
  int unit_rows;                // rows per work unit, a multiple of TM
  const unsigned char *blobs;   // [M][BLOB_BYTES] prepared operands (tc_prep_kernel)
  const int32_t *groups;        // [n_groups][GRP_MAX] window ids, -1 padded; the windows of a group are adjacent
                                // and (first column % 4) + nw * DIM <= BOX_COLS
  const int32_t *from;          // [M] first column of every window
  int n_groups, K;
  void *out;                    // [M][out_stride] uint8 or int32
  i64 out_stride;
  unsigned long long *stats;    // [0] candidate (row, chunk) pairs
  volatile int *dbg;            // host-mapped breadcrumbs [warp][8] of CTA 0 (diagnostics), or null
  // fused Lloyd pass only (FUSE): `out` is int32 and holds the PREVIOUS assignments on entry
  const float *scale;           // [M] 2^S of the fixed-point sums (kupdate.cuh)
  unsigned long long *sums;     // [M][K][dmax] biased fixed-point sums (RED.ADD.64 at the end of a unit)
  int32_t *counts;              // [M][K]
  int32_t *bad;                 // [M][K]
  int32_t *diff;                // [M] rows whose assignment changed
  int dmax;
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
__device__ __forceinline__ bool mb_try(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(sa(b)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware, do not spin
      : "memory");
  return ok != 0;
}
// the same without the suspend-time hint (the instruction still blocks for a hardware-defined time);
// compile with -DGULON_TC_WAIT_NOHINT to use it: measured 731 M vectors/s on the c2 encode against 737 M
__device__ __forceinline__ bool mb_try_nohint(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(sa(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A hand-off that never arrives is a bug in the pipeline: trap (the launch fails with an error and
// the breadcrumbs) instead of hanging the device.  Legitimate waits are microseconds long.
//
// Polling discipline (measured, round 2).  The ncu source page of the encode kernel shows 44 % of the
// executed warp instructions in these wait loops (SYNCS / BRA / ISETP / VIADD / NANOSLEEP.SYNCS: a waiter
// suspended with the hardware hint is woken by every mbarrier event of the CTA and polls again).  Both
// alternatives were measured on the c2 encode and are slower: a fixed back-off (__nanosleep 32..256 ns
// between polls) 524 M vectors/s -- the hand-offs sit on the critical path, a late wake-up stalls the
// two-accumulator pipeline; a pure test_wait spin 484 M -- the spinners take the issue slots; try_wait
// without the hint 731 M.  The suspend-hint wait below gives 737 M.
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t parity) {
#ifdef GULON_TC_WAIT_NOHINT
  for (uint32_t spins = 0; !mb_try_nohint(b, parity); spins++)
    if (spins > (1u << 26)) __trap();
#else
  for (uint32_t spins = 0; !mb_try(b, parity); spins++)
    if (spins > (1u << 22)) __trap();
#endif
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(dst)),
               "l"(src), "r"(bytes), "r"(sa(bar))
               : "memory");
}
// one [TM rows][BOX_COLS floats] box of the matrix; rows / columns outside the matrix arrive as zeros
__device__ __forceinline__ void tma_box(void *dst, const CUtensorMap *map, int col, int row, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          sa(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(sa(bar)), "r"(col), "r"(row)
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
// 32 lanes x 16 consecutive columns (two chunks of 8)
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
// Two IEEE products / sums per instruction (FMUL2 / FADD2).  Each lane is rounded to nearest like
// __fmul_rn / __fadd_rn.  ptxas contracts a packed mul feeding a packed add into FFMA2, so the
// exact path may only pair a packed mul with SCALAR adds (and packed adds with no mul at all).  (A
// packed mul in the exact loop was tried: the operand pairs cost two MOVs each at 80 registers.)
__device__ __forceinline__ u64 mul2_rn(u64 a, u64 b) {  // operands / result: (lo float, hi float) in one b64
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  return (u64)__float_as_uint(lo) | ((u64)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo_of(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return *reinterpret_cast<float2 *>(&r);
}
// minima of the 2 chunks of 8 columns held in r[16]: four FMNMX3 / FMNMX per chunk
__device__ __forceinline__ void chunk_mins16(const uint32_t (&r)[16], float *cm) {
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const float t1 = min3(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]), __uint_as_float(r[c * 8 + 2]));
    const float t2 = min3(__uint_as_float(r[c * 8 + 3]), __uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
    const float t3 = min3(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]), t1);
    cm[c] = fminf(t2, t3);
  }
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
// two floats -> packed bf16x2 (lo in the low half), round to nearest
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- operand preparation: one blob per window ------------------------------------------------------
// grid (nsub), block 256 (thread = centroid).  cb [M][K][dmax] fp32, off [M][K]; blob of window m at m * BLOB_BYTES.
__global__ void __launch_bounds__(256) tc_prep_kernel(const float *__restrict__ cb, const float *__restrict__ off,
                                                      const int32_t *__restrict__ subs,
                                                      const int32_t *__restrict__ dims, int K, int dmax,
                                                      unsigned char *__restrict__ blobs) {
  const int m = subs[blockIdx.x], k = threadIdx.x, dim = dims[m];
  unsigned char *blob = blobs + (size_t)m * BLOB_BYTES;
  float *cbs = reinterpret_cast<float *>(blob + B_BYTES) + (k / CH) * CHUNK_FLOATS + (k % CH) * CB_LD;
  float *meta = reinterpret_cast<float *>(blob + B_BYTES + CB_BYTES);
  __shared__ float s_c[256], s_o[256];
  __shared__ int s_bad;
  if (k == 0) s_bad = 0;
  __syncthreads();
  uint16_t slots[KP];
#pragma unroll
  for (int i = 0; i < KP; i++) slots[i] = 0;
  float nrm2 = 0.0f, o = 0.0f;
  bool bad = false;
  for (int j = 0; j < CB_LD; j++) cbs[j] = 0.0f;
  if ((k % CH) == 0)
    for (int j = 0; j < 4; j++) cbs[CH * CB_LD + j] = 0.0f;
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
      cbs[j] = cj;
      if (!(fabsf(cj) < 1e18f)) bad = true;
    }
    const uint16_t o1 = bf16_bits(o);
    const float r1 = o - bf16_val(o1);
    const uint16_t o2 = bf16_bits(r1);
    const uint16_t o3 = bf16_bits(r1 - bf16_val(o2));
    slots[3 * dim] = o1;
    slots[3 * dim + 1] = o2;
    slots[3 * dim + 2] = o3;
    if (!(fabsf(o) < 1e36f)) bad = true;
    cbs[off_slot(dim)] = o;
  } else {
    // padding centroid: approximate score 3e38 (times the row's 1.0 slot), exact score +inf: never accepted
    if (3 * dim < KP) slots[3 * dim] = bf16_bits(3.0e38f);
    cbs[off_slot(dim < CB_LD - 1 ? dim : 0)] = __int_as_float(0x7f800000);
    o = 0.0f;
  }
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

// ---- reading a row's window out of a 128-byte-swizzled raw tile ---------------------------------------
// Row r of the tile starts at r * 128 bytes; its 16-byte unit u is stored at unit (u ^ (r & 7)).
// `off` = first float of the window inside the box.  The units holding the window are read with
// runtime addresses; the sub-unit shift (off & 3) selects one of four register renamings.
template <int DIM, int SH>
__device__ __forceinline__ void load_window_sh(const unsigned char *row_base, int sw, int u0, float (&x)[DIM]) {
  constexpr int NU = (SH + DIM + 3) / 4;
  float v[NU * 4];
#pragma unroll
  for (int u = 0; u < NU; u++) {
    const float4 q = *reinterpret_cast<const float4 *>(row_base + (((u0 + u) ^ sw) << 4));
    v[u * 4 + 0] = q.x;
    v[u * 4 + 1] = q.y;
    v[u * 4 + 2] = q.z;
    v[u * 4 + 3] = q.w;
  }
#pragma unroll
  for (int j = 0; j < DIM; j++) x[j] = v[SH + j];
}
template <int DIM>
__device__ __forceinline__ void load_window(const unsigned char *row_base, int sw, int off, float (&x)[DIM]) {
  const int u0 = off >> 2;
  switch (off & 3) {
    case 0: load_window_sh<DIM, 0>(row_base, sw, u0, x); break;
    case 1: load_window_sh<DIM, 1>(row_base, sw, u0, x); break;
    case 2: load_window_sh<DIM, 2>(row_base, sw, u0, x); break;
    default: load_window_sh<DIM, 3>(row_base, sw, u0, x); break;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------
// FUSE: one Lloyd pass.  Besides the assignment the sweep thread (which holds the row's window in
// registers) adds the row into the fixed-point sums of its new cluster -- the arithmetic of
// upd::update_fixed_kernel, bit for bit -- and counts the rows whose assignment differs from the one
// `out` held on entry (KMeans.computeClusters' `Arrays.equals(prev, next)`, G/KMeans.scala:144-154).
// Three passes over the matrix per iteration (assign, update, compare) become one.
template <int DIM, typename OutT, bool FUSE = false>
__global__ void __launch_bounds__(NT, 1) tc_assign_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  static_assert(3 * DIM + 3 <= KP, "window too wide for the 48-slot contraction");
  static_assert(DIM <= BOX_COLS - 3, "a window at any alignment must fit one box");
  static_assert(!FUSE || sizeof(OutT) == 4, "the fused pass works on int32 assignments");
  constexpr int DO = off_slot(DIM) + 1;  // floats of a centroid row the exact loop reads
  constexpr int W_MMA = 4 + 4 * NWG, W_TMA = W_MMA + 1;
  constexpr int NBLOB = FUSE ? FUSE_GRP : GRP_MAX;
  constexpr int WSTRIDE = fuse_wstride(DIM);
  extern __shared__ __align__(1024) unsigned char smem[];  // the 128-byte swizzle needs 1024-byte aligned tiles
  unsigned char *raw_s = smem;                                // NRAW raw tiles, 1024-byte aligned
  unsigned char *a_s = raw_s + NRAW * RAW_BYTES;              // 2 A tiles
  unsigned char *blob_s = a_s + 2 * A_BYTES;                  // NBLOB blobs
  unsigned int *acc_s = reinterpret_cast<unsigned int *>(blob_s + NBLOB * BLOB_BYTES);  // FUSE: accumulators
  uint64_t *bars = reinterpret_cast<uint64_t *>(blob_s + NBLOB * BLOB_BYTES + (FUSE ? FUSE_GRP * WSTRIDE * 4 : 0));
  uint64_t *blob_full = bars;         // 1
  uint64_t *blob_empty = bars + 1;    // 1
  uint64_t *raw_full = bars + 2;      // NRAW
  uint64_t *raw_empty = bars + 5;     // NRAW
  uint64_t *a_full = bars + 8;        // 2
  uint64_t *a_empty = bars + 10;      // 2
  uint64_t *t_empty = bars + 12;      // 2: accumulator (tile & 1) swept
  // accumulator of tile t ready, one barrier per sweep group (t % NWG): a barrier with a single
  // waiting group can never be polled two phases ahead
  uint64_t *t_full = bars + 14;       // NWG
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 14 + NWG);
  static_assert((14 + NWG) * 8 + 4 <= BAR_BYTES, "barrier area too small");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#define TC_DBG(slot, val)                                      \
  do {                                                         \
    if (p.dbg && blockIdx.x == 0 && lane == 0) {               \
      p.dbg[warp * 8 + (slot)] = (int)(val);                   \
      __threadfence_system();                                  \
    }                                                          \
  } while (0)
  const i64 n_ranges = (p.N + p.unit_rows - 1) / p.unit_rows;
  const i64 n_units = n_ranges * p.n_groups;

  if (tid == 0) {
    mb_init(blob_full, 1);
    mb_init(blob_empty, 4 * NWG + 1);   // the sweep warps + the MMA commit
    for (int i = 0; i < NRAW; i++) {
      mb_init(raw_full + i, 1);
      mb_init(raw_empty + i, 4 + 4 * NWG);  // 4 split warps + the sweep warps
    }
    for (int i = 0; i < 2; i++) {
      mb_init(a_full + i, 4);
      mb_init(a_empty + i, 1);
      mb_init(t_empty + i, 4);
    }
    for (int i = 0; i < NWG; i++) mb_init(t_full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  TC_DBG(0, 1);
  if constexpr (FUSE)
    for (int t = tid; t < FUSE_GRP * WSTRIDE; t += NT) acc_s[t] = 0;
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  TC_DBG(0, 2);
  TC_DBG(1, tmem_base);

  i64 tile = 0;     // tiles of earlier units (drives the A / accumulator ring indices and phases)
  i64 blk = 0;      // row blocks of earlier units (drives the raw ring)
  uint32_t ui = 0;  // units done (blob phase)

  for (i64 unit = blockIdx.x; unit < n_units; unit += gridDim.x, ui++) {
    const int g = (int)(unit % p.n_groups);
    const i64 r_begin = (unit / p.n_groups) * p.unit_rows;
    const i64 r_end = r_begin + p.unit_rows < p.N ? r_begin + p.unit_rows : p.N;
    const int32_t *gw = p.groups + g * GRP_MAX;
    const int nw = 1 + (gw[1] >= 0 ? 1 : 0) + (gw[2] >= 0 ? 1 : 0);
    // TMA needs a 16-byte aligned box start: the box begins at the 4-float boundary below the
    // group's first column, the windows sit at m0 + w * DIM inside it
    const int from0 = p.from[gw[0]];
    const int col0 = from0 & ~3, m0 = from0 & 3;
    const int n_blocks = (int)((r_end - r_begin + TM - 1) / TM);
    const int n_tiles = n_blocks * nw;

    if (warp == W_TMA) {
      // ================= TMA issue (one thread) =================
      if (lane == 0) {
        auto issue_raw = [&](int b) {
          const i64 bi = blk + b;
          const int rb = (int)(bi % NRAW);
          mb_wait(raw_empty + rb, (uint32_t)(((bi / NRAW) & 1) ^ 1));
          TC_DBG(2, 100 + b);
          mb_expect_tx(raw_full + rb, RAW_BYTES);
          tma_box(raw_s + rb * RAW_BYTES, &tmap, col0, (int)(r_begin + (i64)b * TM), raw_full + rb);
          TC_DBG(2, 200 + b);
        };
        // the first raw tiles only need ring slots of the previous unit, not its operands
        int b = 0;
        for (; b < n_blocks && b < NRAW - 1; b++) issue_raw(b);
        mb_wait(blob_empty, (ui & 1u) ^ 1u);
        TC_DBG(3, 1 + ui);
        mb_expect_tx(blob_full, (uint32_t)nw * BLOB_BYTES);
        for (int w = 0; w < nw; w++)
          bulk_g2s(blob_s + w * BLOB_BYTES, p.blobs + (size_t)gw[w] * BLOB_BYTES, BLOB_BYTES, blob_full);
        for (; b < n_blocks; b++) issue_raw(b);
      }
      __syncwarp();
    } else if (warp == W_MMA) {
      // ================= MMA issue (one thread) =================
      if (lane == 0) {
        mb_wait(blob_full, ui & 1u);
        int w = 0;
        for (int t = 0; t < n_tiles; t++) {
          const i64 ti = tile + t;
          const int ab = (int)(ti & 1);
          mb_wait(a_full + ab, (uint32_t)((ti >> 1) & 1));
          mb_wait(t_empty + ab, (uint32_t)(((ti >> 1) & 1) ^ 1));
          tc_fence_after();
          TC_DBG(2, 100 + t);
          const uint32_t a_addr = sa(a_s + ab * A_BYTES), b_addr = sa(blob_s + w * BLOB_BYTES);
#pragma unroll
          for (int ks = 0; ks < KP / 16; ks++) {
            const u64 ad = smem_desc(a_addr + ks * 256, 128, (KP / 8) * 128);
            const u64 bd = smem_desc(b_addr + ks * 256, 128, (KP / 8) * 128);
            tc_mma_bf16(tmem_base + (uint32_t)(ab * TN), ad, bd, IDESC, ks > 0 ? 1u : 0u);
          }
          tc_commit(a_empty + ab);   // A tile consumed
          tc_commit(t_full + (int)(ti % NWG));  // accumulator ready
          TC_DBG(2, 200 + t);
          if (++w == nw) w = 0;
        }
        tc_commit(blob_empty);       // the unit's B operands are no longer read
      }
      __syncwarp();
    } else if (warp < 4) {
      // ================= split: thread = row of the tile =================
      const int r = tid;
      const unsigned char *row_off = raw_s + r * 128;
      const int sw = r & 7;
      unsigned char *a_row = a_s + (r >> 3) * (KP / 8) * 128 + (r & 7) * 16;
      i64 ti = tile;
      for (int b = 0; b < n_blocks; b++) {
        const i64 bi = blk + b;
        const int rb = (int)(bi % NRAW);
        mb_wait(raw_full + rb, (uint32_t)((bi / NRAW) & 1));
        TC_DBG(2, 100 + b);
        for (int w = 0; w < nw; w++, ti++) {
          float x[DIM];
          load_window<DIM>(row_off + rb * RAW_BYTES, sw, m0 + w * DIM, x);
          // xh = x truncated to bf16 (exactly representable), xl = bf16(x - xh)
          float xh[DIM], xl[DIM];
#pragma unroll
          for (int j = 0; j < DIM; j++) {
            xh[j] = __uint_as_float(__float_as_uint(x[j]) & 0xFFFF0000u);
            xl[j] = x[j] - xh[j];
          }
          auto slot = [&](int s) -> float {
            return s < DIM ? xh[s]
                           : s < 2 * DIM ? xh[s - DIM] : s < 3 * DIM ? xl[s - 2 * DIM] : s < 3 * DIM + 3 ? 1.0f : 0.0f;
          };
          uint4 v[KP / 8];
#pragma unroll
          for (int kc = 0; kc < KP / 8; kc++) {
            v[kc].x = pack_bf16x2(slot(kc * 8 + 0), slot(kc * 8 + 1));
            v[kc].y = pack_bf16x2(slot(kc * 8 + 2), slot(kc * 8 + 3));
            v[kc].z = pack_bf16x2(slot(kc * 8 + 4), slot(kc * 8 + 5));
            v[kc].w = pack_bf16x2(slot(kc * 8 + 6), slot(kc * 8 + 7));
          }
          const int ab = (int)(ti & 1);
          mb_wait(a_empty + ab, (uint32_t)(((ti >> 1) & 1) ^ 1));
#pragma unroll
          for (int kc = 0; kc < KP / 8; kc++) *reinterpret_cast<uint4 *>(a_row + ab * A_BYTES + kc * 128) = v[kc];
          fence_async_smem();  // A is read by the tensor core through the async proxy
          __syncwarp();
          if (lane == 0) mb_arrive(a_full + ab);
          TC_DBG(3, 100 + (int)(ti - tile));
        }
        __syncwarp();
        if (lane == 0) mb_arrive(raw_empty + rb);
      }
    } else {
      // ===== sweep + exact: thread = row = TMEM lane; group wg takes the tiles with index % NWG == wg =====
      const int wg = (warp - 4) >> 2, q4 = warp & 3, r = q4 * 32 + lane;
      const unsigned char *row_off = raw_s + r * 128;
      const int sw = r & 7;
      mb_wait(blob_full, ui & 1u);
      // "This warp is done with row block bi."  The arrival must land in the ring slot's phase OF THAT
      // BLOCK: a warp that skips blocks (tiles of other groups) could otherwise arrive while the slot
      // still holds an earlier block and complete that block's phase early.  Waiting for the block's
      // load first pins the phase (and keeps this warp at most one phase ahead of the barrier).
      auto release_block = [&](i64 bi) {
        const int rb = (int)(bi % NRAW);
        mb_wait(raw_full + rb, (uint32_t)((bi / NRAW) & 1));
        mb_arrive(raw_empty + rb);
      };
      unsigned n_pairs = 0;
      int chg[FUSE_GRP] = {0, 0};   // FUSE: rows of this thread whose assignment changed, per window slot
      float scl[FUSE_GRP] = {0.0f, 0.0f};
      if constexpr (FUSE) {
#pragma unroll
        for (int ws = 0; ws < FUSE_GRP; ws++) scl[ws] = ws < nw ? p.scale[gw[ws]] : 0.0f;
      }
      int rel = 0;  // next row block this warp has not released yet
      int t = (int)(((wg - tile) % NWG + NWG) % NWG);
      int b = t / nw, w = t - b * nw;
      for (; t < n_tiles; t += NWG) {
        const i64 ti = tile + t;
        const int acc = (int)(ti & 1);
        if (rel < b) {
          __syncwarp();
          if (lane == 0)
            for (int j = rel; j < b; j++) release_block(blk + j);
          rel = b;
        }
        // ---- sweep ----
        __syncwarp();  // the exact loop of the previous tile diverges; tcgen05.ld is warp-collective
        mb_wait(t_full + wg, (uint32_t)((ti / NWG) & 1));
        tc_fence_after();
        float cmin[NCH];
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * TN);
        {
          // two register buffers: the next 16 columns are in flight while the previous are reduced
          uint32_t v0[16], v1[16];
          tc_ld16_issue(taddr, v0);
          tc_ld_wait16(v0);
#pragma unroll
          for (int cb = 0; cb < 16; cb += 2) {
            tc_ld16_issue(taddr + (cb + 1) * 16, v1);
            chunk_mins16(v0, cmin + cb * 2);
            tc_ld_wait16(v1);
            if (cb + 2 < 16) tc_ld16_issue(taddr + (cb + 2) * 16, v0);
            chunk_mins16(v1, cmin + (cb + 1) * 2);
            if (cb + 2 < 16) tc_ld_wait16(v0);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mb_arrive(t_empty + acc);  // accumulator drained
        // row minimum: four independent FMNMX3 chains
        float rm[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float a = min3(cmin[q * 8 + 0], cmin[q * 8 + 1], cmin[q * 8 + 2]);
          const float bq = min3(cmin[q * 8 + 3], cmin[q * 8 + 4], cmin[q * 8 + 5]);
          rm[q] = fminf(min3(cmin[q * 8 + 6], cmin[q * 8 + 7], a), bq);
        }
        const float rmin = fminf(fminf(rm[0], rm[1]), fminf(rm[2], rm[3]));
        // ---- the row, its 2E window, the candidate mask ----
        const i64 bi = blk + b;
        const int rb = (int)(bi % NRAW);
        mb_wait(raw_full + rb, (uint32_t)((bi / NRAW) & 1));
        float x[DIM];
        load_window<DIM>(row_off + rb * RAW_BYTES, sw, m0 + w * DIM, x);
        const unsigned char *blob = blob_s + w * BLOB_BYTES;
        const float *meta = reinterpret_cast<const float *>(blob + B_BYTES + CB_BYTES);
        float n2 = 0.0f;
#pragma unroll
        for (int j = 0; j < DIM; j++) n2 = fmaf(x[j], x[j], n2);
        // E = 2^-12 |x| max|c| + 2^-15 max|off| (more than twice the derived bound, see header); window = 2E
        const float win =
            2.0f * (sqrtf(n2) * 1.001f * meta[0] * (1.0f / 4096.0f) + meta[1] * (1.0f / 32768.0f)) + 1e-30f;
        const float thr = rmin + win;
        // non-finite / huge rows or operands, +inf or NaN thresholds: every chunk is a candidate
        const bool all = !(n2 < 1e36f) || meta[2] != 0.0f || !(thr < __int_as_float(0x7f800000));
        // bit c = sign of (cmin[c] - thrU), thrU just above thr so that cmin == thr counts as well
        const float thr_u = thr + (fabsf(thr) * 2.4e-7f + 1e-37f);
        uint32_t mask = 0;
        const float2 nthr = make_float2(-thr_u, -thr_u);
#pragma unroll
        for (int c = NCH - 2; c >= 0; c -= 2) {
          const float2 dlt = add2_rn(make_float2(cmin[c], cmin[c + 1]), nthr);
          mask = __funnelshift_l(__float_as_uint(dlt.y), mask, 1);
          mask = __funnelshift_l(__float_as_uint(dlt.x), mask, 1);
        }
        if (all) mask = 0xffffffffu;
        const i64 row = r_begin + (i64)b * TM + r;
        if (row >= r_end) mask = 0;
        n_pairs += __popc(mask);
        int prev_a = 0;   // FUSE: the row's previous assignment (loaded now, compared after the recheck)
        if constexpr (FUSE) {
          if (row < r_end) prev_a = reinterpret_cast<const int32_t *>(p.out)[(i64)gw[w] * p.out_stride + row];
        }
        // ---- exact evaluation of the candidate chunks ----
        // Reference rule: ascending k, strict '<' from Float.MaxValue = the (score, k)-lexicographic
        // minimum over accepted scores.  The 8 centroids of a chunk are visited in a per-lane rotated
        // order, i = (step + rot) & 7 with rot = 5 (lane - c) mod 8, which puts the 16-byte units the 8
        // lanes of a quarter warp read at one step into 8 different bank groups whatever their chunks
        // are (unit index = 41 c + 5 i + j4 = lane + 5 step + j4 mod 8): conflict-free LDS.128.
        const float *cbs = reinterpret_cast<const float *>(blob + B_BYTES);
        float best = FLT_MAX;
        int idx = 0;
        // One candidate chunk (99 % of the rows): a second, FUSED fp32 filter before the literal
        // arithmetic.  With u = 2^-24 and A_k = sum |x_j c_kj| <= |x||c_k|, the reference's unfused score is
        // within 2 (DIM + 1) u A_k + u |s_k| of the real value and s~_k = fma(-2, fma-chain(x, c_k), off_k)
        // within 2 DIM u A_k + u |s_k|, so the two evaluations of ONE centroid disagree by at most
        //   D = (4 DIM + 2) u |x| max|c| + 2 u max|s|  <=  2^-18 |x| max|c| + 2^-23 max|s|     (DIM <= 15).
        // If exactly one centroid k* of the chunk has s~ within tol >= 2 D of the chunk minimum, then for
        // every other j: s_j >= s~_j - D > s~_k* + tol - D >= s_k* + tol - 2 D >= s_k*, strictly: k* IS the
        // reference's argmin and no exact tie exists.  tol = 2^-16 |x| max|c| + 2^-20 (|min| + max|off|)
        // is more than 4 D.  88 fused operations replace 176 unfused ones.  Anything else -- several
        // centroids within tol (duplicates, near ties), NaN / huge scores, several candidate chunks --
        // takes the literal path below.
        if (!all && (mask & (mask - 1)) == 0 && mask != 0) {
          const int c = __ffs(mask) - 1;
          const float *cc = cbs + c * CHUNK_FLOATS;
          const int rot = (5 * (lane - c)) & 7;
          float sf[CH];
#pragma unroll
          for (int step = 0; step < CH; step++) {
            const int i = (step + rot) & 7;
            float cv[DO];
#pragma unroll
            for (int j4 = 0; j4 < DO / 4; j4++) {
              const float4 q = *reinterpret_cast<const float4 *>(cc + i * CB_LD + 4 * j4);
              cv[4 * j4 + 0] = q.x; cv[4 * j4 + 1] = q.y; cv[4 * j4 + 2] = q.z; cv[4 * j4 + 3] = q.w;
            }
            float d = 0.0f;
#pragma unroll
            for (int j = 0; j < DIM; j++) d = fmaf(x[j], cv[j], d);
            sf[step] = fmaf(-2.0f, d, cv[DO - 1]);
          }
          float m = min3(sf[0], sf[1], sf[2]);
          m = min3(m, sf[3], sf[4]);
          m = min3(m, sf[5], sf[6]);
          m = fminf(m, sf[7]);
          const float tol = sqrtf(n2) * 1.001f * meta[0] * (1.0f / 65536.0f) +
                            (fabsf(m) + meta[1]) * (1.0f / 1048576.0f) + 1e-30f;
          const float lim = m + tol;
          int near = 0, at = 0;
#pragma unroll
          for (int step = 0; step < CH; step++) {
            const bool in = sf[step] <= lim;
            near += in ? 1 : 0;
            at = in ? step : at;
          }
          if (near == 1 && m < 1.0e37f && m > -1.0e37f) {
            idx = c * CH + ((at + rot) & 7);
            mask = 0;   // decided
          }
        }
        while (mask) {
          const int c = __ffs(mask) - 1;
          mask &= mask - 1;
          const float *cc = cbs + c * CHUNK_FLOATS;
          const int rot = (5 * (lane - c)) & 7;
          float sc[CH];
          int kk[CH];
#pragma unroll
          for (int step = 0; step < CH; step++) {
            const int i = (step + rot) & 7;
            float cv[DO];
#pragma unroll
            for (int j4 = 0; j4 < DO / 4; j4++) {
              const float4 q = *reinterpret_cast<const float4 *>(cc + i * CB_LD + 4 * j4);
              cv[4 * j4 + 0] = q.x; cv[4 * j4 + 1] = q.y; cv[4 * j4 + 2] = q.z; cv[4 * j4 + 3] = q.w;
            }
            float d = 0.0f;
#pragma unroll
            for (int j = 0; j < DIM; j++) d = __fadd_rn(d, __fmul_rn(x[j], cv[j]));
            sc[step] = __fsub_rn(cv[DO - 1], __fmul_rn(2.0f, d));
            kk[step] = c * CH + i;
          }
          // the chunk's (score, k)-lexicographic minimum; NaN scores never win (fminf drops them)
          float m = min3(sc[0], sc[1], sc[2]);
          m = min3(m, sc[3], sc[4]);
          m = min3(m, sc[5], sc[6]);
          m = fminf(m, sc[7]);
          int km = 0x7fffffff;
#pragma unroll
          for (int step = 0; step < CH; step++) km = min(km, sc[step] == m ? kk[step] : 0x7fffffff);
          // accepted only below Float.MaxValue (the reference's initial minimum); chunks ascend, so on
          // equal scores the earlier chunk (lower k) stays
          if (m < best) {
            best = m;
            idx = km;
          }
        }
        if (row < r_end) reinterpret_cast<OutT *>(p.out)[(i64)gw[w] * p.out_stride + row] = (OutT)idx;
        if constexpr (FUSE) {
          if (row < r_end) {
            // upd::update_fixed_kernel's accumulation: v = round(x 2^S) + 2^28 into 64-bit sums kept as two
            // 32-bit words; all low-word atomics first (their returns in flight together), carries after
            unsigned int *lo = acc_s + w * WSTRIDE + idx * DIM;
            unsigned int *hi = lo + FUSE_KMAX * DIM;
            const float sc_w = w == 0 ? scl[0] : scl[1];
            bool bad = false;
            unsigned int v[DIM], old[DIM];
#pragma unroll
            for (int j = 0; j < DIM; j++) {
              const float sx = x[j] * sc_w;                // exact: a power of two
              const bool ok = fabsf(sx) <= 268435456.0f;    // finite and within 2^28 (always, for finite x)
              bad = bad || !ok;
              v[j] = ok ? (unsigned int)(__float2int_rn(sx) + (1 << FUSE_FIX_BITS)) : 0u;
            }
#pragma unroll
            for (int j = 0; j < DIM; j++) old[j] = atomicAdd(lo + j, v[j]);
#pragma unroll
            for (int j = 0; j < DIM; j++)
              if (old[j] + v[j] < old[j]) atomicAdd(hi + j, 1u);  // carry out of the low word
            atomicAdd(acc_s + w * WSTRIDE + 2 * FUSE_KMAX * DIM + idx, 1u);
            if (bad) atomicOr(acc_s + w * WSTRIDE + 2 * FUSE_KMAX * DIM + FUSE_KMAX + idx, 1u);
            if (w == 0) chg[0] += prev_a != idx; else chg[1] += prev_a != idx;
          }
        }
        TC_DBG(4, 100 + t);
        // next tile of this group
        w += NWG;
        while (w >= nw) {
          w -= nw;
          b++;
        }
      }
      // release the remaining row blocks and the unit's operands
      __syncwarp();
      if (lane == 0) {
        for (int j = rel; j < n_blocks; j++) release_block(blk + j);
        mb_arrive(blob_empty);
      }
      if (p.stats) {
        n_pairs = __reduce_add_sync(0xffffffffu, n_pairs);
        if (lane == 0) atomicAdd(p.stats, (unsigned long long)n_pairs);
      }
      if constexpr (FUSE) {
        // flush the unit's sums (and clear them for the next unit): the 16 sweep warps only
        constexpr int NSW = 4 * NWG * 32;
        const int st = tid - 128;
#pragma unroll
        for (int ws = 0; ws < FUSE_GRP; ws++) {
          const int c = __reduce_add_sync(0xffffffffu, chg[ws]);
          if (lane == 0 && c && ws < nw) atomicAdd(p.diff + gw[ws], c);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NSW) : "memory");
        for (int ws = 0; ws < nw; ws++) {
          unsigned int *lo = acc_s + ws * WSTRIDE, *hi = lo + FUSE_KMAX * DIM, *cnt = hi + FUSE_KMAX * DIM,
                       *bd = cnt + FUSE_KMAX;
          const int m = gw[ws];
          for (int t2 = st; t2 < p.K * DIM; t2 += NSW) {
            const unsigned long long v = ((unsigned long long)hi[t2] << 32) | lo[t2];
            if (v) {
              atomicAdd(p.sums + ((i64)m * p.K + t2 / DIM) * p.dmax + t2 % DIM, v);
              lo[t2] = 0;
              hi[t2] = 0;
            }
          }
          for (int t2 = st; t2 < p.K; t2 += NSW) {
            if (cnt[t2]) {
              atomicAdd(p.counts + (i64)m * p.K + t2, (int)cnt[t2]);
              cnt[t2] = 0;
            }
            if (bd[t2]) {
              atomicOr(p.bad + (i64)m * p.K + t2, 1);
              bd[t2] = 0;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NSW) : "memory");
      }
    }
    tile += n_tiles;
    blk += n_blocks;
  }

  TC_DBG(0, 3);
  tc_fence_before();
  __syncthreads();
  TC_DBG(0, 4);
  if (warp == W_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  TC_DBG(0, 5);
#undef TC_DBG
}

}  // namespace tca
}  // namespace gulon
