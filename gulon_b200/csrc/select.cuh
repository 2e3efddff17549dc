// select.cuh -- exact (distance, id) top-k selection over arrays of u64 keys.
//
// Replaces the semantics of G/TopKHeap.scala (bounded heap fed in ascending row order) with a
// deterministic lexicographic (distance asc, id asc) selection: the order T/TopKHeapSpec.scala:16-31
// asserts.  Used by the simple scan path, the bootstrap of the fused scan, the shard merge
// (TopKHeap#merge, G/TopKHeap.scala:44-53) and the exact kNN (G/Index.scala:209-229).
//
// One pass sorts chunks of 4096 keys in shared memory (bitonic) and keeps the first kk of each
// chunk; passes repeat until one chunk is left.  Integer work, HBM-bound: 8 B read per key.
#pragma once
#include "common.cuh"

namespace gulon {

constexpr int SEL_CHUNK = 4096;
constexpr int SEL_NT = 512;

// Sorts n (power of two, <= 4096) keys in shared memory, ascending; all nt threads of the block.
__device__ inline void block_bitonic_sort(u64 *s, int n, int tid, int nt) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (n >> 1); t += nt) {
        int pos = 2 * t - (t & (stride - 1));
        u64 a = s[pos], b = s[pos + stride];
        bool up = (pos & size) == 0;
        if ((a > b) == up) {
          s[pos] = b;
          s[pos + stride] = a;
        }
      }
    }
  }
  __syncthreads();
}

// Same, for one warp (n <= 512), synchronised with __syncwarp.
__device__ inline void warp_bitonic_sort(u64 *s, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < (n >> 1); t += 32) {
        int pos = 2 * t - (t & (stride - 1));
        u64 a = s[pos], b = s[pos + stride];
        bool up = (pos & size) == 0;
        if ((a > b) == up) {
          s[pos] = b;
          s[pos + stride] = a;
        }
      }
    }
  }
  __syncwarp();
}

// in : [rows][in_stride] keys, in_stride a multiple of SEL_CHUNK (padding = KEY_SENT)
// out: [rows][out_stride]; chunk c of a row writes its kk smallest keys at c*kk.
__global__ void __launch_bounds__(SEL_NT) select_pass_kernel(const u64 *__restrict__ in,
                                                             i64 in_stride, u64 *__restrict__ out,
                                                             i64 out_stride, int kk) {
  __shared__ u64 s[SEL_CHUNK];
  const i64 row = blockIdx.y;
  const i64 c = blockIdx.x;
  const u64 *src = in + row * in_stride + c * SEL_CHUNK;
  for (int t = threadIdx.x; t < SEL_CHUNK; t += SEL_NT) s[t] = src[t];
  block_bitonic_sort(s, SEL_CHUNK, threadIdx.x, SEL_NT);
  u64 *dst = out + row * out_stride + c * kk;
  for (int t = threadIdx.x; t < kk; t += SEL_NT) dst[t] = s[t];
}

// keys [rows][stride] (first k of each row are the answer) -> ids / dists / sizes
__global__ void unpack_keys_kernel(const u64 *__restrict__ keys, i64 stride, i64 rows, int k,
                                   i64 id_offset, int32_t *__restrict__ ids,
                                   float *__restrict__ dists, int32_t *__restrict__ sizes) {
  i64 q = (i64)blockIdx.x * blockDim.y + threadIdx.y;
  if (q >= rows) return;
  int cnt = 0;
  for (int i = threadIdx.x; i < k; i += 32) {
    u64 key = i < stride ? keys[q * stride + i] : KEY_SENT;   // k may exceed the keys a row has
    bool ok = key != KEY_SENT;
    if (ids) ids[q * k + i] = ok ? (int32_t)((i64)(uint32_t)key + id_offset) : -1;
    if (dists) dists[q * k + i] = ok ? ord2f((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
    cnt += ok;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (sizes && threadIdx.x == 0) sizes[q] = cnt;
}

// lists [S][rows][k] -> keys [rows][stride] with row-major [s*k + i]; tail padded with KEY_SENT
__global__ void gather_lists_kernel(const u64 *__restrict__ lists, int S, i64 rows, int k,
                                    u64 *__restrict__ keys, i64 stride) {
  i64 q = blockIdx.y;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < stride;
       t += (i64)gridDim.x * blockDim.x) {
    u64 v = KEY_SENT;
    if (t < (i64)S * k) {
      int s = (int)(t / k), i = (int)(t % k);
      v = lists[((i64)s * rows + q) * k + i];
    }
    keys[q * stride + t] = v;
  }
}

// (ids, dists) of S shards, shard s at element offset s * shard_stride, each [rows][k] -> keys
// [.][stride] for result rows q_base + blockIdx.y; id < 0 marks an empty slot.  shard_stride =
// rows * k for plain [S][rows][k] arrays; the all-gathered [ids | dists] blocks of the sharded query
// use 2 * rows * k.
__global__ void pack_results_kernel(const int32_t *__restrict__ ids,
                                    const float *__restrict__ dists, int S, i64 shard_stride, int k,
                                    i64 q_base, u64 *__restrict__ keys, i64 stride) {
  const i64 q = q_base + blockIdx.y;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < stride;
       t += (i64)gridDim.x * blockDim.x) {
    u64 v = KEY_SENT;
    if (t < (i64)S * k) {
      int s = (int)(t / k), i = (int)(t % k);
      i64 at = (i64)s * shard_stride + q * k + i;
      int32_t id = ids[at];
      if (id >= 0) v = make_key(dists[at], (uint32_t)id);
    }
    keys[(i64)blockIdx.y * stride + t] = v;
  }
}

// Shard merge for short lists: S * k <= 1024 keys per query (8 shards x k = 10 is 80).  One warp per
// query packs the S candidate lists into keys in shared memory, sorts them and writes the k best --
// no 4096-key selection chunk per query (that path moved 32 KB per query for 160 useful bytes and took
// 19.7 ms for 100k queries x 2 shards).  grid ceil(nq / 4), block 128, dynamic smem 4 * P * 8 bytes.
constexpr int MERGE_WARP_MAX = 1024;
__global__ void __launch_bounds__(128) merge_results_small_kernel(const int32_t *__restrict__ ids,
                                                                  const float *__restrict__ dists, int S,
                                                                  i64 shard_stride, int k, i64 nq, int P,
                                                                  int32_t *__restrict__ out_ids,
                                                                  float *__restrict__ out_dists,
                                                                  int32_t *__restrict__ out_sizes) {
  extern __shared__ __align__(16) unsigned char mrs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u64 *sb = reinterpret_cast<u64 *>(mrs_smem) + (size_t)warp * P;
  const i64 q = (i64)blockIdx.x * 4 + warp;
  if (q >= nq) return;
  for (int t = lane; t < P; t += 32) {
    u64 v = KEY_SENT;
    if (t < S * k) {
      const int s = t / k, i = t % k;
      const i64 at = (i64)s * shard_stride + q * k + i;
      const int32_t id = ids[at];
      if (id >= 0) v = make_key(dists[at], (uint32_t)id);
    }
    sb[t] = v;
  }
  __syncwarp();
  warp_bitonic_sort(sb, P, lane);
  int cnt = 0;
  for (int i = lane; i < k; i += 32) {
    const u64 key = i < P ? sb[i] : KEY_SENT;
    const bool ok = key != KEY_SENT;
    out_ids[q * k + i] = ok ? (int32_t)(uint32_t)key : -1;
    out_dists[q * k + i] = ok ? ord2f((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
    cnt += ok;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (out_sizes && lane == 0) out_sizes[q] = cnt;
}

// k-chunked scans: the kk keys of a pass (rows of `cur`, ascending) go behind the `done` keys already in
// acc [rows][k]; the last key becomes the next pass's floor (KEY_SENT when the pass came back short: no
// row is left).  grid ceil(rows / 8), block (32, 8).
__global__ void append_pass_kernel(const u64 *__restrict__ cur, i64 cur_stride, i64 rows, int k, int done,
                                   int kk, u64 *__restrict__ acc, u64 *__restrict__ floor) {
  const i64 q = (i64)blockIdx.x * blockDim.y + threadIdx.y;
  if (q >= rows) return;
  for (int i = threadIdx.x; i < kk; i += 32) acc[q * k + done + i] = cur[q * cur_stride + i];
  if (threadIdx.x == 0) floor[q] = cur[q * cur_stride + kk - 1];
}

// rows of P <= 1024 keys (a power of two) sorted in place, one warp per row.  grid ceil(rows / 4), block 128,
// dynamic shared memory 4 * P * 8 bytes.
__global__ void __launch_bounds__(128) sort_rows_small_kernel(u64 *__restrict__ keys, int P, i64 rows) {
  extern __shared__ __align__(16) unsigned char srs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u64 *sb = reinterpret_cast<u64 *>(srs_smem) + (size_t)warp * P;
  const i64 q = (i64)blockIdx.x * 4 + warp;
  if (q >= rows) return;
  for (int t = lane; t < P; t += 32) sb[t] = keys[q * P + t];
  __syncwarp();
  warp_bitonic_sort(sb, P, lane);
  for (int t = lane; t < P; t += 32) keys[q * P + t] = sb[t];
}

// empty result rows: id -1, distance +inf, size 0
__global__ void fill_empty_kernel(i64 nq, int k, int32_t *__restrict__ ids,
                                  float *__restrict__ dists, int32_t *__restrict__ sizes) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nq * k) {
    if (ids) ids[t] = -1;
    if (dists) dists[t] = __int_as_float(0x7f800000);
  }
  if (sizes && t < nq) sizes[t] = 0;
}

// ---- any k: full merge sort of a row's keys ------------------------------------------------------
// The chunked selection above keeps k <= SEL_CHUNK / 2 keys per chunk.  The reference's TopKHeap takes
// any k (k > N returns every row, sorted: G/TopKHeap.scala:69-79); for k beyond the chunk limit the keys
// of a row are sorted completely: 4096-key chunks by the bitonic kernel (kk = SEL_CHUNK keeps a whole
// chunk), then log2(chunks) merge passes in which every key finds its place by one binary search in
// the partner run (lower bound from the left run, upper bound from the right: stable, a bijection
// even among equal sentinel keys).  8 B x (1 + log2) per key: a correctness path, not a fast one.
__global__ void __launch_bounds__(256) merge_pass_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
                                                         i64 stride, i64 run) {
  const i64 row = blockIdx.y;
  const u64 *src = in + row * stride;
  u64 *dst = out + row * stride;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < stride; i += (i64)gridDim.x * blockDim.x) {
    const i64 r = i / run;
    const i64 base = (r & ~1LL) * run;
    const bool left = (r & 1) == 0;
    const i64 o0 = left ? base + run : base;                       // partner run [o0, o1)
    i64 o1 = o0 + run;
    if (o0 > stride) o1 = o0;
    if (o1 > stride) o1 = stride;
    const u64 key = src[i];
    i64 lo = o0, hi = o1 > o0 ? o1 : o0;
    while (lo < hi) {                                              // left: #partner < key; right: #partner <= key
      const i64 mid = (lo + hi) >> 1;
      const u64 v = src[mid];
      if (left ? v < key : v <= key) lo = mid + 1; else hi = mid;
    }
    const i64 own0 = left ? base : base + run;
    dst[base + (i - own0) + (lo - o0)] = key;
  }
}

// Runs selection passes until every row holds its k smallest keys at the front of the result.
// keys_a: [rows][stride_a] input (stride_a multiple of SEL_CHUNK); scratch buffers ping-pong.
// On return *result / *result_stride describe where the answer lives.
struct Selector {
  DevBuf ping, pong;
  int run(u64 *keys_a, i64 stride_a, i64 rows, int k, cudaStream_t st, u64 **result,
          i64 *result_stride, u64 *final_out = nullptr, i64 final_stride = 0) {
    GREQUIRE(k >= 1, "k must be >= 1");
    GREQUIRE(stride_a % SEL_CHUNK == 0, "internal: key stride not a multiple of %d", SEL_CHUNK);
    GREQUIRE(rows <= 65535 * 64LL, "internal: too many selection rows");
    if (k > SEL_CHUNK / 2 && stride_a > SEL_CHUNK) return run_sort(keys_a, stride_a, rows, st, result, result_stride);
    int kk = k < SEL_CHUNK ? k : SEL_CHUNK;
    u64 *in = keys_a;
    i64 in_stride = stride_a;
    int phase = 0;
    for (;;) {
      i64 chunks = in_stride / SEL_CHUNK;
      bool last = chunks == 1;
      i64 out_stride = last ? (final_out ? final_stride : SEL_CHUNK)
                            : round_up(chunks * kk, SEL_CHUNK);
      u64 *out;
      if (last && final_out) {
        out = final_out;
      } else {
        DevBuf &b = (phase & 1) ? pong : ping;
        GCHECK(b.ensure((size_t)rows * (size_t)out_stride * sizeof(u64)));
        out = b.as<u64>();
        if (!last)
          GCU(cudaMemsetAsync(out, 0xFF, (size_t)rows * (size_t)out_stride * sizeof(u64), st));
      }
      // grid.y is limited to 65535: tile rows
      for (i64 r0 = 0; r0 < rows; r0 += 65535) {
        i64 nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((unsigned)chunks, (unsigned)nr);
        GLAUNCH(select_pass_kernel, grid, SEL_NT, 0, st, in + r0 * in_stride, in_stride,
                out + r0 * out_stride, out_stride, kk);
      }
      if (last) {
        *result = out;
        *result_stride = out_stride;
        return GULON_OK;
      }
      in = out;
      in_stride = out_stride;
      phase++;
    }
  }
  // full sort of every row (any k): see merge_pass_kernel
  int run_sort(u64 *keys_a, i64 stride, i64 rows, cudaStream_t st, u64 **result, i64 *result_stride) {
    GCHECK(ping.ensure((size_t)rows * (size_t)stride * sizeof(u64)));
    GCHECK(pong.ensure((size_t)rows * (size_t)stride * sizeof(u64)));
    const i64 chunks = stride / SEL_CHUNK;
    for (i64 r0 = 0; r0 < rows; r0 += 65535) {
      const i64 nr = rows - r0 < 65535 ? rows - r0 : 65535;
      dim3 grid((unsigned)chunks, (unsigned)nr);
      GLAUNCH(select_pass_kernel, grid, SEL_NT, 0, st, keys_a + r0 * stride, stride,
              ping.as<u64>() + r0 * stride, stride, SEL_CHUNK);
    }
    u64 *in = ping.as<u64>(), *out = pong.as<u64>();
    for (i64 run = SEL_CHUNK; run < stride; run <<= 1) {
      for (i64 r0 = 0; r0 < rows; r0 += 65535) {
        const i64 nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((unsigned)(ceil_div(stride, 256) < 4096 ? ceil_div(stride, 256) : 4096), (unsigned)nr);
        GLAUNCH(merge_pass_kernel, grid, 256, 0, st, in + r0 * stride, out + r0 * stride, stride, run);
      }
      u64 *t = in;
      in = out;
      out = t;
    }
    *result = in;
    *result_stride = stride;
    return GULON_OK;
  }
  void release() {
    ping.release();
    pong.release();
  }
};

}  // namespace gulon
