// tscan.cuh -- PQIndex.batchQuery for large query batches: a tcgen05 lower-bound filter over the DECODED rows,
// exact re-evaluation of the survivors.
//
// Reference: PQIndex.distances / batchQuery, G/Index.scala:393-440: for every (query, row) the fp32 sum
//   d = (((0 + T[0][c_0]) + T[1][c_1]) + ...),  T[m][c] = sequential fp32 sum of (q_i - centroid_i)^2,
// and the k smallest (distance, id) per query (TopKHeap, G/TopKHeap.scala:44-94).
//
// d is, up to fp32 rounding, the squared distance between the query and the row's RECONSTRUCTION x^ (the
// concatenation of its centroids):  d* = |q|^2 + |x^|^2 - 2 q.x^.  The pruned scan (pscan.cuh) bounds d
// from below with 8-bit table gathers from shared memory -- one byte through the 128 B/clk crossbar per
// (row, query, quantizer), the roof it sits on.  This path bounds it with ONE dense contraction instead:
//
//   A[row][.]   = [ bf16(x^_0 .. x^_{D-1}) | a1 a2 | nb | 1 | 0.. ]      (rows of the index, built once)
//   B[query][.] = [ -2 bf16(q_0 .. q_{D-1}) | 1  1 | -e | c | 0.. ]      (per stage: holds tau)
//   D = A B^T   =  (|x^|^2)'' - 2 qb.xb - e nb + (|q|^2 - tau')''        (fp32 accumulate in TMEM)
//
// a1+a2 <= |x^|^2 (1 - EPS_ACC), nb >= |x^|, e >= COEF |q|, c <= C - 2 EPS_ACC |C| (every piece rounded
// towards the safe side) with C = |q|^2 - tau' and tau' = tau (1 + 2^-11) (the fp32 summation error of the
// reference, as in pscan.cuh).  Four extra columns: D = 300 contracts over KP = 304 = 19 K16 steps.
// With the operand rounding |q.x^ - qb.xb| <= (2^-8 + 2^-18) |q||x^| and an accumulation error of the
// tensor core of at most EPS_ACC times the sum of the absolute products (EPS_ACC = eps_acc(KP) = 2^-12 up to
// KP = 320; measured < 2^-16, tests/test_gpu_tscan.py), COEF = 2^-7 + 2^-17 + 2.03 EPS_ACC gives
//
//   D_computed <= d* - tau'     for every (row, query),
//
// so a row whose reference distance is <= tau has D <= 0: **no row that can enter a top-k list is ever
// dropped**; rows with D <= 0 ("survivors", ~1e-5 of the pairs) are re-evaluated with the reference's
// literal fp32 table sum and merged into the lists, which makes the result bit-identical to the exact
// kernels and to the oracle.  tau is the k-th best distance known so far: the first 256 rows are evaluated
// exactly (boot_kernel; for large k the first 8192 through fscan), then stages of geometrically growing length
// run qprep -> filter -> eval -> merge.
//
// Filter kernels (persistent; item = (block of 256 queries, row split), items ordered split-major so that all
// CTAs stream the same rows from L2):
//   filter_kernel            single CTAs: the whole B operand (<= 5 chunks of [256 queries][64 bf16]) resident, a
//                            3-slot ring of A chunks ([128 rows][64 bf16]), tcgen05.mma.cta_group::1 M128 N256 K16
//   filter2_kernel<false>    CTA pairs (the default): half of B per CTA, 9-slot ring, cta_group::2 M256 N256 K16
//   filter2_kernel<true>     CTA pairs, D > 316: B chunks streamed with the A chunks (7 slots of 32 KB)
// warp 0 TMA, warp 1 (pair kernel: and warp 19, alternate tiles) MMA issue, warps 2-17 epilogue (thread = row =
// TMEM lane, 64 of the 256 columns each: two tcgen05.ld.32x32b.x32, accumulator handed back, then FMNMX tree and
// the survivors appended with one atomic per warp and 32 columns), warp 18 optional sentinel.
// Roofline: the tensor pipe -- 2 * 128 * 256 * KP flop per tile; bytes are secondary (DRAM < 1 % of the HBM peak).
#pragma once
#include "common.cuh"
#include "select.cuh"
#include "tcassign.cuh"

namespace gulon {
namespace tscan {

using tca::mb_arrive;
using tca::mb_expect_tx;
using tca::mb_init;
using tca::mb_wait;
using tca::sa;
using tca::tc_commit;
using tca::tc_fence_after;
using tca::tc_fence_before;
using tca::tc_ld32_issue;
using tca::tc_ld_wait;
using tca::tc_mma_bf16;
using tca::tma_box;

constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int TN = 256;            // queries per block (UMMA N)
constexpr int KC = 64;             // bf16 per K chunk = one 128-byte swizzle span
constexpr int NKC_MAX = 5;         // resident B chunks: KP <= 320
constexpr int KP_MAX = NKC_MAX * KC;
constexpr int NEXTRA = 4;          // a1 a2 | nb | 1   /   1 1 | -e | c   (D = 300 -> KP = 304: 19 K16 steps)
constexpr int NSTAGE = 3;          // A ring
constexpr int A_BYTES = TM * KC * 2;   // 16384
constexpr int B_BYTES = TN * KC * 2;   // 32768
constexpr int NEPI = 16;           // epilogue warps: 4 per TMEM lane quadrant, 64 columns each
constexpr int NT = 32 * (2 + NEPI + 2);   // TMA, MMA, epilogue warps, sentinel, second MMA warp (pair kernel)
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = NKC_MAX * B_BYTES + NSTAGE * A_BYTES + BAR_BYTES;
static_assert(SMEM_BYTES <= 232448, "the filter must fit the 227 KB of an sm_100 CTA");

// allowance for the tensor core's fp32 accumulation, per unit of the sum of |products|: 2^-12 up to KP = 320,
// growing with the depth of the contraction beyond (KP 2^-20: 2^-10 at KP = 1008); measured errors are
// more than 16x smaller (tests/test_gpu_tscan.py)
__host__ __device__ inline double eps_acc(int KP) { return KP <= 320 ? 1.0 / 4096.0 : (double)KP / 1048576.0; }
__host__ __device__ inline double bound_coef(int KP) { return 1.0 / 128.0 + 1.0 / 131072.0 + 2.03 * eps_acc(KP); }
constexpr int KP_STREAM_MAX = 1024;        // deepest contraction of the streamed-B form of the pair kernel
constexpr double TAU_SLACK = 1.0 / 2048.0;  // the reference's fp32 summation error (pscan.cuh uses the same)
constexpr float FINITE_MAX = 1e30f;         // larger norms go to the fallback path (products must not overflow)

__host__ __device__ inline int padded_k(int D) { return (D + NEXTRA + 15) / 16 * 16; }

// ---- bf16 helpers (bit patterns) ------------------------------------------------------------------
__device__ __forceinline__ uint16_t bf_rn(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float bf_val(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }
// toward zero: drop the low 16 bits
__device__ __forceinline__ uint16_t bf_rz(float f) { return (uint16_t)(__float_as_uint(f) >> 16); }
// toward -inf / +inf of a finite float
__device__ __forceinline__ uint16_t bf_rd(float f) {
  const uint32_t u = __float_as_uint(f);
  uint16_t b = (uint16_t)(u >> 16);
  if ((u & 0xFFFFu) && (u & 0x80000000u)) b++;   // negative with dropped bits: one step away from zero
  return b;
}
__device__ __forceinline__ uint16_t bf_ru(float f) {
  const uint32_t u = __float_as_uint(f);
  uint16_t b = (uint16_t)(u >> 16);
  if ((u & 0xFFFFu) && !(u & 0x80000000u)) b++;  // positive with dropped bits: one step up
  return b;
}
// v as two bf16 pieces with p1 + p2 <= v and v - (p1 + p2) <= 2^-15 |v| (the first truncates, the second rounds down)
__device__ __forceinline__ void split2_down(double v, uint16_t &p1, uint16_t &p2) {
  p1 = bf_rz(__double2float_rz(v));
  const double r1 = v - (double)bf_val(p1);
  p2 = bf_rd(__double2float_rd(r1));
}

// ---- A operand: the decoded rows, built once per index ---------------------------------------------
// One warp per row.  rowcodes [N][rcs] (the row's M centroid ids side by side), cb [M][K][dmax] fp32,
// colm / colt [D]: sub-quantizer and coordinate of every column.  xb [N][KP] bf16.
// bad[0] is set when a row's norm is not finite or too large for the filter (the index then keeps the
// pruned scan).
__global__ void __launch_bounds__(256) decode_rows_kernel(const uint8_t *__restrict__ rowcodes, i64 rcs, i64 N,
                                                          const float *__restrict__ cb, int K, int dmax,
                                                          const int16_t *__restrict__ colm,
                                                          const int16_t *__restrict__ colt, int D, int KP,
                                                          uint16_t *__restrict__ xb, int *__restrict__ bad) {
  const i64 row = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const uint8_t *rc = rowcodes + row * rcs;
  uint16_t *out = xb + row * KP;
  double nrm = 0.0;
  for (int j = lane; j < D; j += 32) {
    const int m = colm[j], t = colt[j];
    const float v = __ldg(cb + ((i64)m * K + rc[m]) * dmax + t);
    out[j] = bf_rn(v);
    nrm += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  if (lane == 0) {
    uint16_t a1 = 0, a2 = 0, nb = 0;
    if (!(nrm <= (double)FINITE_MAX)) {
      atomicExch(bad, 1);
    } else {
      split2_down(nrm * (1.0 - eps_acc(KP) - 1e-9), a1, a2);
      nb = bf_ru(__double2float_ru(sqrt(nrm) * (1.0 + 1e-9)));
    }
    out[D + 0] = a1;
    out[D + 1] = a2;
    out[D + 2] = nb;
    out[D + 3] = 0x3F80;   // 1.0
  }
  for (int j = D + NEXTRA + lane; j < KP; j += 32) out[j] = 0;
}

// ---- B operand: one row per query, rebuilt for every stage (it holds the threshold) -----------------
// One warp per query slot (nslots = query blocks * 256).  cur [nq][stride]: the sorted best keys so far.
// A query the filter cannot serve (non-finite coordinates, huge norm, no finite threshold yet) sets
// bit 0 of flag[0] and bad[q]; its row -- and the rows of the padding slots -- let nothing survive.
__global__ void __launch_bounds__(256) qprep_kernel(const float *__restrict__ Q, i64 ldq, i64 nq, i64 nslots, int D,
                                                    int KP, const u64 *__restrict__ cur, i64 stride, int k,
                                                    uint16_t *__restrict__ qb, int *__restrict__ flag,
                                                    uint8_t *__restrict__ bad) {
  const i64 q = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nslots) return;
  uint16_t *out = qb + q * KP;
  bool ok = q < nq;
  double nrm = 0.0;
  if (ok) {
    const float *qv = Q + q * ldq;
    for (int j = lane; j < D; j += 32) {
      const float v = qv[j];
      if (!(fabsf(v) <= FINITE_MAX)) ok = false;
      nrm += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    ok = __all_sync(0xffffffffu, ok) && nrm <= (double)FINITE_MAX;
    float tau = 0.0f;
    if (ok) {
      const u64 key = cur[q * stride + (k - 1)];
      tau = ord2f((uint32_t)(key >> 32));
      if (key == KEY_SENT || !(tau >= 0.0f) || !(tau <= FINITE_MAX)) ok = false;
    }
    if (!ok && lane == 0) {
      atomicOr(flag, 1);
      if (bad) bad[q] = 1;   // this query is answered by the pruned scan afterwards (the rest of the batch stays)
    }
    if (ok) {
      for (int j = lane; j < D; j += 32) out[j] = bf_rn(-2.0f * qv[j]);
      if (lane == 0) {
        const double taup = (double)tau * (1.0 + TAU_SLACK);
        const double c = nrm * (1.0 - 1e-9) - taup;
        out[D + 0] = 0x3F80;
        out[D + 1] = 0x3F80;
        out[D + 2] = bf_ru(__double2float_ru(bound_coef(KP) * sqrt(nrm) * (1.0 + 1e-9))) | 0x8000u;   // -e
        out[D + 3] = bf_rd(__double2float_rd(c - 2.0 * eps_acc(KP) * fabs(c) - 1e-300));       // one piece, rounded down
      }
    }
  }
  if (!ok) {
    // c = 2^126 (column D + 3 against the rows' 1.0), zeros elsewhere: D = 2^126 > 0 for every row, nothing survives
    for (int j = lane; j < D + NEXTRA; j += 32) out[j] = j == D + 3 ? (uint16_t)0x7E80 : (uint16_t)0;
  }
  for (int j = D + NEXTRA + lane; j < KP; j += 32) out[j] = 0;
}

// ---- the filter -------------------------------------------------------------------------------------
struct FParams {
  i64 sfrom, suntil;     // rows of this stage
  i64 split_len;         // rows per split
  int S;                 // row splits
  int NB;                // query blocks
  int nkc;               // 64-column chunks of the operands (1 .. NKC_MAX)
  int ksteps;            // KP / 16
  u64 *surv;             // [NB][capb] (query in block << 32 | row)
  unsigned *scount;      // [NB]
  unsigned capb;
  int *flag;             // [0] |= 2 on overflow
  float *dump;           // diagnostics: D of every (row - sfrom, query slot) [rows][NB * 256], or null
  unsigned long long *stats;  // [0] tiles, [1] warp slow paths
  int epi_wait;          // see mb_wait_epi
};

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ u64 desc_sw128(uint32_t addr) {
  return (u64)((addr >> 4) & 0x3FFFu) | ((u64)1 << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Issue forms for a CONVERGED warp: every lane executes the statement, one elected lane issues the
// instruction.  With warp-uniform operands the compiler keeps descriptors and addresses in uniform
// registers (UTCHMMA straight from UR), where a `if (lane == 0)` block costs an ELECT + R2UR waterfall of
// ~20 instructions per MMA -- more than the 128 cycles an M128 N256 K16 instruction runs.
__device__ __forceinline__ u64 make_desc(uint32_t lo, uint32_t hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
constexpr uint32_t DESC_HI = (uint32_t)(((1024u >> 4)) | (1u << 14) | (2u << 29));   // SBO 1024, version 1, 128-byte swizzle
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void mma1_elect(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(make_desc(alo, DESC_HI)), "l"(make_desc(blo, DESC_HI)), "r"(tca::IDESC), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit1_elect(uint64_t *bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(sa(bar))
      : "memory");
}

// Hand-offs on the critical path (accumulator full / empty, ring slot full) are polled without the
// suspend hint: a suspended waiter wakes up ~1 us after the barrier completes, and with two accumulators
// every tile pays two such wake-ups in series (ncu, round 2: the MMA warp waited for `tempty`, the
// epilogue warps for `tfull`, the tensor pipe was 59 % active).  The CTA has 18 warps on four schedulers
// and little else to issue, so spinning costs nothing that matters.
__device__ __forceinline__ void mb_spin(uint64_t *b, uint32_t parity);
// how the epilogue warps wait for an accumulator (FParams::epi_wait): 0 spin, 1 try_wait without the
// suspend hint, 2 try_wait with it, 3 (default) a named barrier released by the sentinel warp
__device__ __forceinline__ void mb_wait_epi(uint64_t *b, uint32_t parity, int mode) {
  if (mode == 0) {
    mb_spin(b, parity);
  } else if (mode == 1) {
    for (uint32_t spins = 0; !tca::mb_try_nohint(b, parity); spins++)
      if (spins > (1u << 26)) __trap();
  } else {
    mb_wait(b, parity);
  }
}
__device__ __forceinline__ void mb_spin(uint64_t *b, uint32_t parity) {
  const uint32_t addr = sa(b);
  for (uint32_t spins = 0;; spins++) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 28)) __trap();
  }
}

// Waking the epilogue warps (FParams::epi_wait == 3, the default): a suspended mbarrier waiter wakes up
// ~1 us after the phase completes, and with two accumulators that latency sits on the critical path of
// every other tile (M + wake-up + sweep must fit into 2 M; at KP = 144 it did not: tensor pipe 62 %).
// Spinning in all 16 epilogue warps takes the issue slots the MMA warp needs.  So ONE extra warp (the
// sentinel) spins on `tfull` and releases the epilogue warps through a named hardware barrier, which wakes
// its waiters within tens of cycles and costs no issue slots while they wait.
__device__ __forceinline__ void epi_bar_arrive(uint32_t id) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"((NEPI + 1) * 32) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(uint32_t id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"((NEPI + 1) * 32) : "memory");
}

// One epilogue warp, one accumulator tile: thread = row = TMEM lane, COLS = 64 columns (queries) starting
// at column c0 of the query block.  sweep_load brings the columns into registers -- after it the
// accumulator can be handed back to the MMA warp -- and sweep_scan works on the registers: per 32
// columns the minimum; a warp that saw D <= 0 appends its survivors (one atomic per warp and 32 columns).
template <int COLS>
__device__ __forceinline__ void sweep_load(uint32_t taddr, uint32_t (&v)[COLS / 32][32]) {
#pragma unroll
  for (int j = 0; j < COLS / 32; j++) tc_ld32_issue(taddr + j * 32, v[j]);
#pragma unroll
  for (int j = 0; j < COLS / 32; j++) tc_ld_wait(v[j]);
}
// minimum of 32 accumulator columns: four independent chains of three-input minima (FMNMX3): 16 + 3
// instructions where a tree of two-input minima takes 31 -- at KP = 144 the epilogue warps were 56 % busy,
// half of it here
__device__ __forceinline__ float min32(const uint32_t (&x)[32]) {
  float m0 = __uint_as_float(x[0]), m1 = __uint_as_float(x[1]), m2 = __uint_as_float(x[2]), m3 = __uint_as_float(x[3]);
#pragma unroll
  for (int i = 4; i < 28; i += 8) {
    m0 = tca::min3(m0, __uint_as_float(x[i]), __uint_as_float(x[i + 1]));
    m1 = tca::min3(m1, __uint_as_float(x[i + 2]), __uint_as_float(x[i + 3]));
    m2 = tca::min3(m2, __uint_as_float(x[i + 4]), __uint_as_float(x[i + 5]));
    m3 = tca::min3(m3, __uint_as_float(x[i + 6]), __uint_as_float(x[i + 7]));
  }
  m0 = tca::min3(m0, __uint_as_float(x[28]), __uint_as_float(x[29]));
  m1 = tca::min3(m1, __uint_as_float(x[30]), __uint_as_float(x[31]));
  return fminf(tca::min3(m0, m1, m2), m3);
}
template <int COLS>
__device__ __forceinline__ void sweep_scan(const FParams &p, const uint32_t (&v)[COLS / 32][32], i64 row, i64 r1, int qb,
                                           int c0, int lane) {
  float mn[COLS / 32];
  float all = 1.0f;
#pragma unroll
  for (int j = 0; j < COLS / 32; j++) {
    mn[j] = min32(v[j]);
    all = fminf(all, mn[j]);
  }
  // one vote per tile and warp; the rest only runs for a warp that holds a survivor (or under the diagnostics dump)
  if (!(__any_sync(0xffffffffu, all <= 0.0f && row < r1) || p.dump)) return;
#pragma unroll
  for (int j = 0; j < COLS / 32; j++) {
    const uint32_t(&x)[32] = v[j];
    const bool hit = mn[j] <= 0.0f && row < r1;
    if (__any_sync(0xffffffffu, hit) || p.dump) {
      if (p.dump && row < r1) {
        float *dst = p.dump + (row - p.sfrom) * ((i64)p.NB * TN) + (i64)qb * TN + c0 + j * 32;
#pragma unroll
        for (int c = 0; c < 32; c++) dst[c] = __uint_as_float(x[c]);
      }
      uint32_t mask = 0;
      if (hit) {
#pragma unroll
        for (int c = 0; c < 32; c++) mask |= (__uint_as_float(x[c]) <= 0.0f ? 1u : 0u) << c;
      }
      // the warp's entries go to consecutive slots: exclusive prefix of the lanes' counts, one atomic
      const uint32_t n = __popc(mask);
      uint32_t incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      uint32_t base = 0;
      if (lane == 0) {
        base = atomicAdd(p.scount + qb, total);
        if (p.stats) atomicAdd(p.stats + 1, 1ull);
      }
      base = __shfl_sync(0xffffffffu, base, 0) + incl - n;
      while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        if (base < p.capb)
          p.surv[(size_t)qb * p.capb + base] = ((u64)(uint32_t)(c0 + j * 32 + c) << 32) | (u64)(uint32_t)row;
        else
          atomicOr(p.flag, 2);
        base++;
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 1) filter_kernel(const __grid_constant__ CUtensorMap mapA,
                                                       const __grid_constant__ CUtensorMap mapB, const FParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *b_s = smem;                                // nkc chunks of [256][128 B]
  unsigned char *a_s = smem + NKC_MAX * B_BYTES;            // NSTAGE tiles of [128][128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(a_s + NSTAGE * A_BYTES);
  uint64_t *full = bars;                  // NSTAGE
  uint64_t *empty = bars + NSTAGE;        // NSTAGE
  uint64_t *bfull = bars + 2 * NSTAGE;    // 1
  uint64_t *bempty = bfull + 1;           // 1
  uint64_t *tfull = bempty + 1;           // 2
  uint64_t *tempty = tfull + 2;           // 2
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(tempty + 2);
  static_assert((2 * NSTAGE + 6) * 8 + 4 <= BAR_BYTES, "barrier area too small");

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform in the compiler's eyes too
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; i++) {
      mb_init(full + i, 1);
      mb_init(empty + i, 1);
    }
    mb_init(bfull, 1);
    mb_init(bempty, 1);
    for (int i = 0; i < 2; i++) {
      mb_init(tfull + i, 1);
      mb_init(tempty + i, NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  const i64 n_items = (i64)p.NB * p.S;
  // (32-bit counters carried across items; see filter2_kernel)
  uint32_t it = 0, rs = 0, rph = 0, tcnt = 0;
  const uint32_t nkc = (uint32_t)p.nkc;
  const uint32_t a_base = sa(a_s), b_base = sa(b_s);
  const uint32_t tail_ks = (uint32_t)p.ksteps - (nkc - 1) * (KC / 16);

  for (i64 item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
    const int qb = (int)(item % p.NB);
    const i64 sp = item / p.NB;
    const i64 r0 = p.sfrom + sp * p.split_len;
    const i64 r1 = r0 + p.split_len < p.suntil ? r0 + p.split_len : p.suntil;
    const uint32_t n_tiles = (uint32_t)((r1 - r0 + TM - 1) / TM);

    if (warp == 0) {
      // ================= TMA issue (one thread) =================
      if (lane == 0) {
        const uint32_t n_chunks = n_tiles * nkc;
        const uint32_t n_pre = n_chunks < (uint32_t)NSTAGE ? n_chunks : (uint32_t)NSTAGE;
        auto load_b = [&]() {
          mb_wait(bempty, (it & 1u) ^ 1u);
          mb_expect_tx(bfull, nkc * B_BYTES);
          for (uint32_t kc = 0; kc < nkc; kc++) tma_box(b_s + kc * B_BYTES, &mapB, (int)(kc * KC), qb * TN, bfull);
        };
        int row = (int)r0;
        uint32_t kc = 0;
        // the first A chunks only need ring slots of the previous item, not its B operand
        for (uint32_t ci = 0; ci < n_chunks; ci++) {
          if (ci == n_pre) load_b();
          if (p.epi_wait & 4) mb_wait(empty + rs, rph ^ 1u); else mb_spin(empty + rs, rph ^ 1u);
          mb_expect_tx(full + rs, A_BYTES);
          tma_box(a_s + rs * A_BYTES, &mapA, (int)(kc * KC), row, full + rs);
          if (++kc == nkc) {
            kc = 0;
            row += TM;
          }
          if (++rs == (uint32_t)NSTAGE) {
            rs = 0;
            rph ^= 1u;
          }
        }
        if (n_pre == n_chunks) load_b();
      }
      __syncwarp();
    } else if (warp == 1) {
      // ================= MMA issue (one elected lane) =================
      mb_wait(bfull, it & 1u);
      for (uint32_t t = 0; t < n_tiles; t++, tcnt++) {
        const uint32_t acc = tcnt & 1u;
        mb_spin(tempty + acc, ((tcnt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * TN;
        for (uint32_t kc = 0; kc < nkc; kc++) {
          mb_spin(full + rs, rph);
          tc_fence_after();
          const uint32_t alo = desc_lo(a_base + rs * A_BYTES), blo = desc_lo(b_base + kc * B_BYTES);
          if (kc + 1 < nkc) {
            mma1_elect(d_addr, alo, blo, kc);
            mma1_elect(d_addr, alo + 2, blo + 2, 1u);
            mma1_elect(d_addr, alo + 4, blo + 4, 1u);
            mma1_elect(d_addr, alo + 6, blo + 6, 1u);
            commit1_elect(empty + rs);      // the ring slot is free once these MMAs have read it
          } else {
            mma1_elect(d_addr, alo, blo, kc);
            if (tail_ks > 1) mma1_elect(d_addr, alo + 2, blo + 2, 1u);
            if (tail_ks > 2) mma1_elect(d_addr, alo + 4, blo + 4, 1u);
            if (tail_ks > 3) mma1_elect(d_addr, alo + 6, blo + 6, 1u);
            commit1_elect(empty + rs);
            commit1_elect(tfull + acc);     // accumulator complete
          }
          if (++rs == (uint32_t)NSTAGE) {
            rs = 0;
            rph ^= 1u;
          }
        }
      }
      commit1_elect(bempty);   // the item's B operand is no longer read
    } else if (warp == 2 + NEPI) {
      // ================= sentinel: spins on `tfull`, releases the epilogue warps through a named barrier =================
      if ((p.epi_wait & 3) == 3) {
        for (uint32_t t = 0; t < n_tiles; t++, tcnt++) {
          const uint32_t acc = tcnt & 1u;
          mb_spin(tfull + acc, (tcnt >> 1) & 1u);
          epi_bar_arrive(1u + acc);
        }
      }
    } else if (warp < 2 + NEPI) {
      // ================= epilogue: thread = row = TMEM lane, 64 of the 256 columns =================
      constexpr int COLS = TN / (NEPI / 4);
      const int ew = warp - 2, quad = warp & 3, c0 = (ew >> 2) * COLS;
      i64 row = r0 + quad * 32 + lane;
      for (uint32_t t = 0; t < n_tiles; t++, tcnt++, row += TM) {
        const uint32_t acc = tcnt & 1u;
        if ((p.epi_wait & 3) == 3)
          epi_bar_sync(1u + acc);
        else
          mb_wait_epi(tfull + acc, (tcnt >> 1) & 1u, p.epi_wait & 3);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN + (uint32_t)c0;
        uint32_t v[COLS / 32][32];
        sweep_load<COLS>(taddr, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mb_arrive(tempty + acc);   // the columns are in registers: the accumulator is free
        sweep_scan<COLS>(p, v, row, r1, qb, c0, lane);
      }
      if (ew == 0 && lane == 0 && p.stats) atomicAdd(p.stats, (unsigned long long)n_tiles);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- the filter on CTA pairs (cta_group::2) -----------------------------------------------------------
// The single-CTA kernel above keeps the whole B operand (256 queries, 160 KB at KP = 320) in its shared
// memory, which leaves three 16 KB ring slots for A: 48 KB in flight against a load-to-free latency of
// ~2700 cycles is 18 B/clk where the tensor pipe wants 32 (ncu: tensor pipe 56 % active).  Here two CTAs
// of a cluster share one item: each holds HALF of B (its 128 queries ... of the N dimension) and its own
// 128-row A tiles; the leader issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 from the leader's
// shared memory, 128-255 from the peer's; N = 256: the two halves of B), the accumulator rows land in each
// CTA's own TMEM.  Per CTA: 80 KB of B, a ring of NSTAGE2 = 9 A slots (144 KB in flight), and the tensor
// core reads 8 KB instead of 12 KB of shared memory per instruction.
// Barriers: full / bfull live in the leader (both CTAs' TMA loads complete_tx on them, cp.async.bulk.tensor
// .cta_group::2); empty / bempty / tfull are signalled in BOTH CTAs by multicast tcgen05.commit; tempty
// lives in the leader and collects the epilogue warps of both CTAs (remote mbarrier.arrive).
constexpr int NSTAGE2 = 9;
constexpr int BH_BYTES = (TN / 2) * KC * 2;   // 16384: one chunk of a CTA's half of B
constexpr int SMEM2_BYTES = NKC_MAX * BH_BYTES + NSTAGE2 * A_BYTES + BAR_BYTES;
static_assert(SMEM2_BYTES <= 232448, "the pair filter must fit the 227 KB of an sm_100 CTA");
// instruction descriptor: D = F32, A = B = BF16, K-major, N = 256, M = 256 (cta_group::2)
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `p` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(sa(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mb_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA box whose completion is counted by an mbarrier given as a shared::cluster address (the leader's)
__device__ __forceinline__ void tma_box2(void *dst, const CUtensorMap *map, int col, int row, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          sa(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(col), "r"(row)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t *bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(sa(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void mma2_elect(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(make_desc(alo, DESC_HI)), "l"(make_desc(blo, DESC_HI)), "r"(IDESC2), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit2_elect(uint64_t *bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(sa(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// SB ("streamed B", KP > 320: c5's D = 1000): the query block does not fit next to the ring, so every ring slot
// carries the A chunk AND this CTA's half of the B chunk (32 KB, 7 slots); the B block (256 queries x KP) stays
// hot in L2, at twice the L2 -> shared-memory traffic of the resident form.
template <bool SB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1)
    filter2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const FParams p) {
  constexpr int NSTAGE2 = SB ? 7 : tscan::NSTAGE2;
  constexpr int SLOT = SB ? A_BYTES + BH_BYTES : A_BYTES;
  static_assert(NSTAGE2 * SLOT + (SB ? 0 : NKC_MAX * BH_BYTES) + BAR_BYTES <= SMEM2_BYTES, "shared memory layout");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *b_s = smem;                                // !SB: nkc chunks of [128][128 B]: this CTA's half of B
  unsigned char *a_s = smem + (SB ? 0 : NKC_MAX * BH_BYTES);  // NSTAGE2 slots: [128][128 B] of A (SB: + [128][128 B] of B)
  uint64_t *bars = reinterpret_cast<uint64_t *>(a_s + NSTAGE2 * SLOT);
  uint64_t *full = bars;                   // NSTAGE2 (leader's are used)
  uint64_t *empty = bars + NSTAGE2;        // NSTAGE2
  uint64_t *bfull = bars + 2 * NSTAGE2;    // 1 (leader's)
  uint64_t *bempty = bfull + 1;            // 1
  uint64_t *tfull = bempty + 1;            // 2
  uint64_t *tempty = tfull + 2;            // 2 (leader's)
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(tempty + 2);
  static_assert((2 * NSTAGE2 + 6) * 8 + 4 <= BAR_BYTES, "barrier area too small");

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform in the compiler's eyes too
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE2; i++) {
      mb_init(full + i, 1);
      mb_init(empty + i, 1);
    }
    mb_init(bfull, 1);
    mb_init(bempty, 2);   // both issuing warps commit it
    for (int i = 0; i < 2; i++) {
      mb_init(tfull + i, 1);
      mb_init(tempty + i, 2 * NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  const i64 n_items = (i64)p.NB * p.S;
  const i64 n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  // Ring / accumulator positions are plain 32-bit counters carried across items: the issue loops of the
  // TMA and MMA threads are single-thread instruction streams, every instruction in them is latency.
  uint32_t it = 0;                 // items done by this pair
  uint32_t rs = 0, rph = 0;        // ring slot of the next chunk and its phase parity (TMA and MMA warps, each its own copy)
  uint32_t tcnt = 0;               // tiles done (MMA and epilogue warps): accumulator tcnt & 1, phase (tcnt >> 1) & 1
  const uint32_t nkc = (uint32_t)p.nkc;
  const uint32_t full0 = map_to_cta(full, 0), bfull0 = map_to_cta(bfull, 0), tempty0 = map_to_cta(tempty, 0);
  const uint32_t a_base = sa(a_s), b_base = sa(b_s);
  const uint32_t tail_ks = (uint32_t)p.ksteps - (nkc - 1) * (KC / 16);   // K16 steps of the last chunk (1..4)

  for (i64 item = pair; item < n_items; item += n_pairs, it++) {
    const int qb = (int)(item % p.NB);
    const i64 sp = item / p.NB;
    const i64 r0 = p.sfrom + sp * p.split_len;
    const i64 r1 = r0 + p.split_len < p.suntil ? r0 + p.split_len : p.suntil;
    const uint32_t n_tiles = (uint32_t)((r1 - r0 + 2 * TM - 1) / (2 * TM));   // tiles of 256 rows, 128 per CTA
    const i64 my0 = r0 + (i64)rank * TM;

    if (warp == 0) {
      // ================= TMA issue (one thread per CTA) =================
      if (lane == 0) {
        const uint32_t n_chunks = n_tiles * nkc;
        const uint32_t n_pre = n_chunks < (uint32_t)NSTAGE2 ? n_chunks : (uint32_t)NSTAGE2;
        auto load_b = [&]() {
          if constexpr (SB) return;
          mb_wait(bempty, (it & 1u) ^ 1u);
          if (leader) mb_expect_tx(bfull, 2u * nkc * BH_BYTES);
          for (uint32_t kc = 0; kc < nkc; kc++)
            tma_box2(b_s + kc * BH_BYTES, &mapB, (int)(kc * KC), qb * TN + (int)rank * (TN / 2), bfull0);
        };
        int row = (int)my0;
        uint32_t kc = 0;
        // the first A chunks only need ring slots of the previous item, not its B operand
        for (uint32_t ci = 0; ci < n_chunks; ci++) {
          if (ci == n_pre) load_b();
          if (p.epi_wait & 4) mb_wait(empty + rs, rph ^ 1u); else mb_spin(empty + rs, rph ^ 1u);
          if (leader) mb_expect_tx(full + rs, 2 * SLOT);   // both CTAs' boxes
          tma_box2(a_s + rs * SLOT, &mapA, (int)(kc * KC), row, full0 + rs * 8u);
          if constexpr (SB)
            tma_box2(a_s + rs * SLOT + A_BYTES, &mapB, (int)(kc * KC), qb * TN + (int)rank * (TN / 2), full0 + rs * 8u);
          if (++kc == nkc) {
            kc = 0;
            row += 2 * TM;
          }
          if (++rs == (uint32_t)NSTAGE2) {
            rs = 0;
            rph ^= 1u;
          }
        }
        if (n_pre == n_chunks) load_b();
      }
      __syncwarp();
    } else if (warp == 1 || warp == 3 + NEPI) {
      // ================= MMA issue (the leader's warps 1 and 3 + NEPI, one elected lane each) =================
      // Two issuing warps on alternate tiles (warp 1: accumulator 0, the other: accumulator 1).  At KP = 144
      // a tile is 9 MMAs = 1152 cycles of tensor time, and one thread's instruction stream (~15 per MMA,
      // ~40 per ring hand-off) took ~1535 (timing experiment in profiles/README.md).  tcgen05.commit tracks
      // the MMAs of the committing thread, so each warp frees its own ring slots and publishes its own tiles;
      // both commit `bempty` (count 2).
      // (SB: a tile is up to 16 chunks but the ring has 7 slots, so a warp that skips a tile would look at a
      // slot two laps ahead and a parity wait cannot tell laps apart: warp 1 issues every tile there -- with
      // >= 21 MMAs per tile the issue stream is not the limiter anyway)
      if (leader && (!SB || warp == 1)) {
        const uint32_t mine = warp == 1 ? 0u : 1u;
        if constexpr (!SB) mb_wait(bfull, it & 1u);
        for (uint32_t t = 0; t < n_tiles; t++, tcnt++) {
          const uint32_t acc = tcnt & 1u;
          if (!SB && acc != mine) {   // the other warp's tile: skip its ring slots
            rs += nkc;
            while (rs >= (uint32_t)NSTAGE2) {
              rs -= (uint32_t)NSTAGE2;
              rph ^= 1u;
            }
            continue;
          }
          mb_spin(tempty + acc, ((tcnt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_addr = tmem_base + acc * TN;
          for (uint32_t kc = 0; kc < nkc; kc++) {
            mb_spin(full + rs, rph);
            tc_fence_after();
            const uint32_t alo = desc_lo(a_base + rs * SLOT);
            const uint32_t blo = SB ? desc_lo(a_base + rs * SLOT + A_BYTES) : desc_lo(b_base + kc * BH_BYTES);
            if (kc + 1 < nkc) {
              mma2_elect(d_addr, alo, blo, kc);
              mma2_elect(d_addr, alo + 2, blo + 2, 1u);
              mma2_elect(d_addr, alo + 4, blo + 4, 1u);
              mma2_elect(d_addr, alo + 6, blo + 6, 1u);
              commit2_elect(empty + rs);
            } else {
              mma2_elect(d_addr, alo, blo, kc);
              if (tail_ks > 1) mma2_elect(d_addr, alo + 2, blo + 2, 1u);
              if (tail_ks > 2) mma2_elect(d_addr, alo + 4, blo + 4, 1u);
              if (tail_ks > 3) mma2_elect(d_addr, alo + 6, blo + 6, 1u);
              commit2_elect(empty + rs);
              commit2_elect(tfull + acc);
            }
            if (++rs == (uint32_t)NSTAGE2) {
              rs = 0;
              rph ^= 1u;
            }
          }
        }
        if constexpr (!SB) {
          commit2_elect(bempty);
          // the last item's multicast arrivals must have landed before either CTA may exit
          if (warp == 1 && item + n_pairs >= n_items) mb_wait(bempty, it & 1u);
        }
      }
    } else if (warp == 2 + NEPI) {
      // ================= sentinel (both CTAs): spins on `tfull`, releases the epilogue warps =================
      if ((p.epi_wait & 3) == 3) {
        for (uint32_t t = 0; t < n_tiles; t++, tcnt++) {
          const uint32_t acc = tcnt & 1u;
          mb_spin(tfull + acc, (tcnt >> 1) & 1u);
          epi_bar_arrive(1u + acc);
        }
      }
    } else if (warp < 2 + NEPI) {
      // ================= epilogue: this CTA's 128 rows of every tile =================
      constexpr int COLS = TN / (NEPI / 4);
      const int ew = warp - 2, quad = warp & 3, c0 = (ew >> 2) * COLS;
      i64 row = my0 + quad * 32 + lane;
      for (uint32_t t = 0; t < n_tiles; t++, tcnt++, row += 2 * TM) {
        const uint32_t acc = tcnt & 1u;
        if ((p.epi_wait & 3) == 3)
          epi_bar_sync(1u + acc);
        else
          mb_wait_epi(tfull + acc, (tcnt >> 1) & 1u, p.epi_wait & 3);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN + (uint32_t)c0;
        uint32_t v[COLS / 32][32];
        sweep_load<COLS>(taddr, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mb_arrive_cluster(tempty0 + acc * 8u);   // the columns are in registers: the accumulator is free
        sweep_scan<COLS>(p, v, row, r1, qb, c0, lane);
      }
      if (ew == 0 && lane == 0 && p.stats) atomicAdd(p.stats, (unsigned long long)n_tiles);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- queries the filter handed back: gather them for the pruned scan, scatter its answers ------------
__global__ void gather_queries_kernel(const float *__restrict__ Q, i64 ldq, const int32_t *__restrict__ idx, int D,
                                      float *__restrict__ out) {
  const float *src = Q + (i64)idx[blockIdx.x] * ldq;
  for (int j = threadIdx.x; j < D; j += blockDim.x) out[(i64)blockIdx.x * D + j] = src[j];
}
__global__ void scatter_results_kernel(const int32_t *__restrict__ idx, int k, const int32_t *__restrict__ ids,
                                       const float *__restrict__ dists, const int32_t *__restrict__ sizes,
                                       int32_t *__restrict__ d_ids, float *__restrict__ d_dists,
                                       int32_t *__restrict__ d_sizes) {
  const i64 q = idx[blockIdx.x];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    d_ids[q * k + j] = ids[(i64)blockIdx.x * k + j];
    d_dists[q * k + j] = dists[(i64)blockIdx.x * k + j];
  }
  if (threadIdx.x == 0 && d_sizes) d_sizes[q] = sizes[blockIdx.x];
}

// ---- the first rows, exactly ---------------------------------------------------------------------------
// PQIndex.distances (G/Index.scala:393-409) of rows [from, from + nrows), nrows <= BOOT_ROWS, for the 4 queries
// of a table group, and their k best: the lists and thresholds the first filter stage starts from.  (The
// exact scan kernel pays 30 shared-memory table fills per (8192-row chunk, query group) and starts from empty
// lists -- 20-26 ms for 100 000 queries whatever the row count; this is one pass over the group's tables.)
// grid (G), block BOOT_ROWS (thread = row: the code planes are read coalesced), out [G * 4][k].
constexpr int BOOT_ROWS = 256;
__global__ void __launch_bounds__(BOOT_ROWS) boot_kernel(const uint8_t *__restrict__ codes, i64 ps, i64 from, int nrows,
                                                         const float4 *__restrict__ lutI, int M, int k,
                                                         u64 *__restrict__ out) {
  __shared__ u64 keys[4][BOOT_ROWS];
  const int g = blockIdx.x, tid = threadIdx.x;
  const i64 row = from + tid;
  const bool valid = tid < nrows;
  float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
  if (valid) {
    const float4 *lut = lutI + (i64)g * M * 256;
    for (int m = 0; m < M; m++) {
      const float4 t = __ldg(lut + m * 256 + codes[(i64)m * ps + row]);
      d0 = __fadd_rn(d0, t.x);
      d1 = __fadd_rn(d1, t.y);
      d2 = __fadd_rn(d2, t.z);
      d3 = __fadd_rn(d3, t.w);
    }
  }
  keys[0][tid] = valid ? make_key(d0, (uint32_t)row) : KEY_SENT;
  keys[1][tid] = valid ? make_key(d1, (uint32_t)row) : KEY_SENT;
  keys[2][tid] = valid ? make_key(d2, (uint32_t)row) : KEY_SENT;
  keys[3][tid] = valid ? make_key(d3, (uint32_t)row) : KEY_SENT;
  // one warp pair per query would do; the block-wide sort is simpler and this kernel is ~1 % of a batch
  for (int j = 0; j < 4; j++) block_bitonic_sort(keys[j], BOOT_ROWS, tid, BOOT_ROWS);
  for (int t = tid; t < 4 * k; t += BOOT_ROWS) out[((i64)g * 4 + t / k) * k + t % k] = keys[t / k][t % k];
}

// ---- survivors -> exact distances -> per-query candidate lists --------------------------------------
// grid (chunks, NB).  PQIndex.distances for one (query, row): the fp32 table sum in the reference's order.
struct EParams {
  const u64 *surv;
  const unsigned *scount;
  unsigned capb;
  const uint8_t *rowcodes;
  i64 rcs;
  const float4 *lutI;     // [G][M][256] interleaved tables (scan.cuh)
  int M;
  i64 nq;
  const u64 *cur;         // [nq][stride] best keys so far (sorted)
  i64 stride;
  int k;
  u64 *cand;              // [nq][capq]
  unsigned *ccount;       // [nq]
  unsigned capq;
  int *flag;              // |= 4 on overflow
  unsigned long long *stats;  // [2] survivors evaluated, [3] candidates kept
};

__global__ void __launch_bounds__(256) eval_kernel(const EParams p) {
  const int b = blockIdx.y;
  unsigned n = p.scount[b];
  if (n > p.capb) n = p.capb;
  unsigned kept = 0;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
    const u64 e = p.surv[(size_t)b * p.capb + i];
    const uint32_t row = (uint32_t)e;
    const i64 q = (i64)b * TN + (i64)(e >> 32);
    if (q >= p.nq) continue;
    const uint8_t *rc = p.rowcodes + (i64)row * p.rcs;
    const float *lut = reinterpret_cast<const float *>(p.lutI + (q >> 2) * p.M * 256) + (q & 3);
    float d = 0.0f;
    for (int m0 = 0; m0 < p.M; m0 += 16) {
      const uint4 w = *reinterpret_cast<const uint4 *>(rc + m0);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      const int mm = p.M - m0 < 16 ? p.M - m0 : 16;
      float tv[16];
#pragma unroll
      for (int j = 0; j < 16; j++)
        if (j < mm) tv[j] = __ldg(lut + ((i64)(m0 + j) * 256 + ((ww[j >> 2] >> ((j & 3) * 8)) & 255u)) * 4);
#pragma unroll
      for (int j = 0; j < 16; j++)
        if (j < mm) d = __fadd_rn(d, tv[j]);
    }
    const u64 key = make_key(d, row);
    if (key < p.cur[q * p.stride + (p.k - 1)]) {
      const unsigned slot = atomicAdd(p.ccount + q, 1u);
      if (slot < p.capq)
        p.cand[q * (i64)p.capq + slot] = key;
      else
        atomicOr(p.flag, 4);
      kept++;
    }
  }
  if (p.stats) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 2, (unsigned long long)n);
    if (kept) atomicAdd(p.stats + 3, (unsigned long long)kept);
  }
}

// ---- candidates + best list -> best list: one warp per query ----------------------------------------
constexpr int MERGE_WARPS = 4;
constexpr int MERGE_SORTN = 2048;   // k + capq <= MERGE_SORTN
__global__ void __launch_bounds__(32 * MERGE_WARPS) merge_kernel(const u64 *__restrict__ cur, i64 stride, i64 nq, int k,
                                                                 const u64 *__restrict__ cand,
                                                                 unsigned *__restrict__ ccount, unsigned capq,
                                                                 u64 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char msm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u64 *sb = reinterpret_cast<u64 *>(msm) + (size_t)w * MERGE_SORTN;
  const i64 q = (i64)blockIdx.x * MERGE_WARPS + w;
  if (q >= nq) return;
  unsigned n = ccount[q];
  if (n > capq) n = capq;
  if (n == 0) {
    for (int i = lane; i < k; i += 32) out[q * k + i] = cur[q * stride + i];
    return;
  }
  const int total = k + (int)n;
  int P = 64;
  while (P < total) P <<= 1;
  for (int i = lane; i < k; i += 32) sb[i] = cur[q * stride + i];
  for (int i = lane; i < (int)n; i += 32) sb[k + i] = cand[q * (i64)capq + i];
  for (int i = total + lane; i < P; i += 32) sb[i] = KEY_SENT;
  warp_bitonic_sort(sb, P, lane);
  for (int i = lane; i < k; i += 32) out[q * k + i] = sb[i];
  __syncwarp();
  if (lane == 0) ccount[q] = 0;
}

}  // namespace tscan
}  // namespace gulon
