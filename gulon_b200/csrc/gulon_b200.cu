// gulon_b200.cu -- C ABI (include/gulon_b200.h) over the sm_100a kernels.
//
// Host logic only: argument checks (the reference's require(...)), handle ownership, chunking of
// host buffers through HBM, and the driver loops of KMeans.computeClusters (G/KMeans.scala:134-157)
// and ProductQuantizer.apply (G/ProductQuantizer.scala:121-153).  All arithmetic of the path runs
// in the kernels of kmeans.cuh / scan.cuh / select.cuh; there is NO CPU fallback -- without a CUDA
// device every compute entry point returns GULON_ENODEVICE.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <memory>
#include <set>

#include "common.cuh"
#include "mlctl.h"
#include "kmeans.cuh"
#include "pscan.cuh"
#include "scan.cuh"
#include "select.cuh"
#include "tcassign.cuh"
#include "tscan.cuh"
#include "kupdate.cuh"
#include "synth.cuh"

using namespace gulon;

// Scratch that lives in a handle is reused by the next call, which may arrive on another stream (the
// mutex that guards it is released when the work is ENQUEUED, not when it finishes).  Every user
// brackets its work with enter()/leave(): a caller on a different stream first waits, on the device,
// for the previous user's last kernel.  Calls on one stream are ordered by the stream itself.
struct StreamChain {
  cudaEvent_t ev = nullptr;
  cudaStream_t last = nullptr;
  bool used = false;
  int enter(cudaStream_t st) {
    if (used && st != last) GCU(cudaStreamWaitEvent(st, ev, 0));
    return GULON_OK;
  }
  int leave(cudaStream_t st) {
    if (!ev) GCU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GCU(cudaEventRecord(ev, st));
    last = st;
    used = true;
    return GULON_OK;
  }
  ~StreamChain() {
    if (ev) cudaEventDestroy(ev);
  }
};

// ---- handles --------------------------------------------------------------------------------
struct gulon_points_s {
  float *d = nullptr;
  i64 N = 0;
  int D = 0;
  i64 ld = 0;
  bool owned = false;
};

struct gulon_codebook_s {
  int D = 0, M = 0, K = 0, dmax = 0;
  std::vector<int32_t> from, dim;
  DevBuf cb, off, dfrom, ddim;
  // sub-quantizers grouped by window width (the split rule yields at most two widths)
  std::map<int, std::vector<int32_t>> by_dim;
  std::map<int, DevBuf> d_by_dim;
  std::mutex mu;
  DevBuf scratch_q, scratch_lut;  // gulon_prepare_query staging
  // gulon_pq_encode staging, kept across calls (no cudaMalloc / cudaFree / stream creation per call)
  cudaStream_t enc_st[2] = {nullptr, nullptr};
  DevBuf enc_x[2], enc_c[2];
  DevBuf tc;                      // tensor-core assignment operands (tcassign.cuh)
  std::map<int, DevBuf> tc_groups;  // per window width: groups of adjacent windows
  std::map<int, int> tc_ngroups;
  ~gulon_codebook_s() {
    tc.release();
    for (auto &kv : tc_groups) kv.second.release();
    cb.release(); off.release(); dfrom.release(); ddim.release();
    for (auto &kv : d_by_dim) kv.second.release();
    scratch_q.release(); scratch_lut.release();
    for (int i = 0; i < 2; i++) {
      enc_x[i].release();
      enc_c[i].release();
      if (enc_st[i]) cudaStreamDestroy(enc_st[i]);
    }
  }
};

struct gulon_index_s {
  gulon_codebook_t cb = nullptr;
  const uint8_t *codes = nullptr;
  const uint16_t *codes16 = nullptr;  // wide index (256 < K <= 65536): 16-bit ids, `codes` is null
  i64 N = 0, ps = 0;
  bool owned = false;
  DevBuf lutW;
  std::mutex mu;  // guards the scratch below while a query batch is being enqueued
  StreamChain chain;   // ... and orders batches that arrive on different streams (see StreamChain)
  std::mutex mu_host;  // guards the host-call staging buffers
  std::mutex mu_sh;    // guards the scratch of the sharded query (gulon_pq_query_sharded*)
  StreamChain chain_sh;
  DevBuf sh_send, sh_recv, sh_slice, sh_all, sh_q, sh_cand;
  DevBuf h_q, h_ids, h_dists, h_sizes;
  DevBuf lutI, keys, lists, qbuf, ids, dists, sizes, merged;
  DevBuf qlut, mins, qp, boot_tail, plists, pstats, boot_keys, spread, msel, merged2, sufmin;
  DevBuf rowcodes;            // row-major copy of the planes, built by the first pruned scan
  DevBuf kacc, kfloor;        // k-chunked scans: accumulated keys [Q4][k], last key of the previous pass [Q4]
  // tensor scan (tscan.cuh): xb = the decoded rows as bf16 operand rows [N][tensor_kp], built by the first
  // tensor scan; the rest is per-batch scratch
  DevBuf xb, tcol, tqb, tsurv, tscount, tcand, tccount, tcur0, tcur1, tflag, tstats, tbad, tbq, tbi, tbd, tbs, tbx;
  int tensor_state = 0;       // 0 not tried, 1 ready, -1 unavailable (shape, memory, non-finite rows)
  int tensor_kp = 0;
  // wide indexes (16-bit ids): 8-bit GROUP planes for the lower-bound scan (pscan.cuh), built by the first
  // pruned scan: gmap [M][K] group of every centroid, members / start = the groups' member lists
  DevBuf gmap, gmembers, gstart, gcodes, rowcodes16;
  i64 gps = 0, rcs16 = 0;
  int wide_state = 0;         // 0 not tried, 1 ready, -1 unavailable
  i64 rcs = 0;
  int rowcodes_state = 0;     // 0 not tried, 1 ready, -1 unavailable (no memory): planes are used
  Selector sel, sel_boot, sel2;
  // lower-bound subset size learned from the survivor rate of earlier launches (0: none yet)
  // size of the lower-bound subset, steered by the measured time per (row, query) pair of earlier
  // main-stage launches on this index (hill climbing; see scan_batch)
  gulon::MlController mlc;
  cudaEvent_t tm_ev0 = nullptr, tm_ev1 = nullptr;
  bool tm_pending = false;
  int tm_ml = 0;           // subset size of the launch being timed
  double tm_pairs = 0;     // its (row, query) pairs
  long long tm_shape = 0;  // its (tiles, rows)
  ~gulon_index_s() {
    if (owned && codes) cudaFree((void *)codes);
    if (owned && codes16) cudaFree((void *)codes16);
    lutW.release();
    lutI.release(); keys.release(); lists.release(); qbuf.release();
    ids.release(); dists.release(); sizes.release(); merged.release();
    qlut.release(); mins.release(); qp.release(); boot_tail.release(); plists.release();
    pstats.release(); boot_keys.release(); spread.release(); msel.release(); merged2.release();
    sel2.release(); sufmin.release(); rowcodes.release(); kacc.release(); kfloor.release();
    gmap.release(); gmembers.release(); gstart.release(); gcodes.release(); rowcodes16.release();
    xb.release(); tcol.release(); tqb.release(); tsurv.release(); tscount.release(); tcand.release();
    tccount.release(); tcur0.release(); tcur1.release(); tflag.release(); tstats.release();
    tbad.release(); tbq.release(); tbi.release(); tbd.release(); tbs.release(); tbx.release();
    if (tm_ev0) cudaEventDestroy(tm_ev0);
    if (tm_ev1) cudaEventDestroy(tm_ev1);
    h_q.release(); h_ids.release(); h_dists.release(); h_sizes.release();
    sel.release(); sel_boot.release();
    sh_send.release(); sh_recv.release(); sh_slice.release(); sh_all.release(); sh_q.release();
    sh_cand.release();
  }
};

namespace {

// ---- options / device -----------------------------------------------------------------------
std::atomic<long long> g_scan_impl{GULON_SCAN_AUTO};
std::atomic<long long> g_query_batch{0};      // 0 = auto (multiple of 16 * #SM)
std::atomic<long long> g_simple_scratch{1LL << 30};
std::atomic<long long> g_encode_chunk{1 << 18};  // rows per H2D chunk in gulon_pq_encode (two chunks in flight)
std::atomic<long long> g_fused_min_rows{16384};
std::atomic<long long> g_boot_rows{0};           // rows scanned exactly to seed the pruned scan; 0 = range / 64 in [8192, 32768], whole 8192-row chunks
std::atomic<long long> g_pruned_min_rows{1 << 18};
std::atomic<long long> g_pruned_bits{0};         // 0 = auto, 8 or 16: width of the lower-bound fields
std::atomic<long long> g_pruned_words{0};        // 0 = auto, 1/2/4: 32-bit words per table entry
std::atomic<long long> g_pruned_lb{0};           // 0 = auto (feedback), else quantizers in the lower bound
std::atomic<long long> g_pruned_stage_div{32};   // first stage = range / div rows with the full bound; 0 = one stage
std::atomic<long long> g_pruned_rowcodes{1};     // keep a row-major copy of the codes for the survivor evaluation
std::atomic<long long> g_tensor_min_rows{1 << 16};    // GULON_SCAN_AUTO: shorter ranges keep the pruned / exact scan
std::atomic<long long> g_tensor_min_pairs{1LL << 27}; // ... and batches with fewer (row, query) pairs
std::atomic<long long> g_tensor_min_queries{256};    // ... and smaller batches
std::atomic<long long> g_tensor_query_batch{0};       // queries per pass of the tensor scan; 0 = auto
std::atomic<long long> g_tensor_ratio{0};             // a stage scans ratio x the rows seen so far; 0 = auto from k
std::atomic<long long> g_tensor_boot{0};              // rows scanned exactly first; 0 = 8192
std::atomic<long long> g_tensor_max_bytes{64LL << 30};  // largest decoded copy of an index
std::atomic<long long> g_tensor_eval_blocks{16};          // CTAs per query block in tscan::eval_kernel
std::atomic<long long> g_tensor_epi_wait{2};              // tscan::mb_wait_epi
std::atomic<long long> g_tensor_pair{1};                  // the filter on CTA pairs (cta_group::2) or on single CTAs
std::atomic<long long> g_tensor_chunk_bytes{16LL << 20};  // operand rows of one row split (L2 working set)
std::atomic<unsigned long long> g_tstats[8];          // tiles, slow paths, survivors, candidates, pairs, fallbacks, batches, stages
std::atomic<long long> g_tscan_bad_queries{0};   // queries the tensor scan handed to the pruned scan one by one
std::atomic<long long> g_last_scan{0};           // GULON_SCAN_* that answered the last batch
std::atomic<long long> g_last_ml{0};             // quantizers in the lower bound of the last main stage
std::atomic<long long> g_assign_impl{GULON_ASSIGN_AUTO};  // exact CUDA-core kernel or tcgen05 filter + exact recheck
std::atomic<long long> g_assign_tc_min_rows{4096};
std::atomic<long long> g_update_fixed{1};        // GULON_UPDATE_SUM as the exact fixed-point sum when possible
std::atomic<long long> g_train_fused{1};         // sum mode: assignment + sums + changed count in one pass over the matrix
std::atomic<unsigned long long> g_tc_stats[3];   // candidate (row, chunk) pairs, overflow tiles, (row, window) pairs
std::atomic<long long> g_last_qt{0};             // queries per tile of the last pruned launch
std::atomic<long long> g_train_updates{0};       // last training: centroid updates of the longest-running window
std::atomic<long long> g_train_window_passes{0}; // last training: (window, assignment pass) pairs
std::atomic<unsigned long long> g_ppairs_main{0}; // (row, query) pairs of main-stage pruned launches
std::atomic<unsigned long long> g_pstats[3];     // survivors, list candidates, slow-path items
std::atomic<unsigned long long> g_ppairs{0};     // (row, query) pairs offered to the pruned kernel

// ---- optional per-launch timing of the dominant kernels (bench.py's roofline leg) -------------
// With option "profile" = 1 every fused-scan / assign launch is bracketed by CUDA events on the
// launching stream; gulon_get_counter("scan_kernel_ns" | "assign_kernel_ns" | ..._launches)
// synchronises the pending events and returns the running totals.
std::atomic<long long> g_profile{0};
struct KernelTimer {
  std::mutex mu;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  double ns = 0;
  long long launches = 0;
  cudaEvent_t begin(cudaStream_t st) {
    if (!g_profile.load()) return nullptr;
    cudaEvent_t e0;
    if (cudaEventCreate(&e0) != cudaSuccess) return nullptr;
    cudaEventRecord(e0, st);
    return e0;
  }
  void end(cudaEvent_t e0, cudaStream_t st) {
    if (!e0) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) != cudaSuccess) {
      cudaEventDestroy(e0);
      return;
    }
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lock(mu);
    pending.emplace_back(e0, e1);
  }
  void drain() {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &pr : pending) {
      float ms = 0;
      if (cudaEventSynchronize(pr.second) == cudaSuccess &&
          cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
        ns += (double)ms * 1e6;
        launches += 1;
      }
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    pending.clear();
  }
  void reset() {
    drain();
    std::lock_guard<std::mutex> lock(mu);
    ns = 0;
    launches = 0;
  }
};
KernelTimer g_t_tscan;  // tscan::filter_kernel
KernelTimer g_t_scan, g_t_assign, g_t_pscan, g_t_pscan_first;  // pscan: main stage; first: the short full-bound stage

int need_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(GULON_ENODEVICE,
                "no CUDA device visible (%s): libgulon_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  return GULON_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is PER DEVICE (and per function): a
// host that drives several GPUs from one process (gulon_set_device per thread) needs it on each.
int smem_optin(const void *func, size_t bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void *, int>> done;
  int dev = 0;
  GCU(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({func, dev})) return GULON_OK;
  GCU(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done.insert({func, dev});
  return GULON_OK;
}
#define GOPTIN(kern, bytes) GCHECK(smem_optin(reinterpret_cast<const void *>(kern), (bytes)))

// cudaMemcpy2D[Async] with rows that are contiguous on both sides is one linear copy; the runtime does
// not always collapse it (1M rows of 1200 B took 1.1 s through the 2-D path: one DMA descriptor per row).
inline cudaError_t copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                          size_t height, cudaMemcpyKind kind, cudaStream_t st, bool async) {
  if (height == 0 || width == 0) return cudaSuccess;
  if (dpitch == width && spitch == width)
    return async ? cudaMemcpyAsync(dst, src, width * height, kind, st) : cudaMemcpy(dst, src, width * height, kind);
  return async ? cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, st)
               : cudaMemcpy2D(dst, dpitch, src, spitch, width, height, kind);
}

// ---- java.util.Random (JDK javadoc algorithm) -- seeds KMeans.init, G/KMeans.scala:188-196 ----
struct JRandom {
  uint64_t s;
  explicit JRandom(int64_t seed) : s(((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
  int32_t next(int bits) {
    s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(int64_t)(s >> (48 - bits));
  }
  int32_t next_int(int32_t bound) {
    int32_t r = next(31);
    const int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r;; u = next(31)) {
      r = u % bound;
      if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) return r;
    }
  }
};

int split_rule(int D, int M, int32_t *from, int32_t *dim) {
  const int ideal = (D + M - 1) / M;
  const int full = M - (ideal * M - D);
  for (int i = 0; i < M; i++) {
    if (i < full) {
      from[i] = i * ideal;
      dim[i] = ideal;
    } else {
      from[i] = full * ideal + (i - full) * (ideal - 1);
      dim[i] = ideal - 1;
    }
  }
  return ideal;
}

template <typename T>
int upload(DevBuf &b, const std::vector<T> &v, cudaStream_t st) {
  GCHECK(b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) {
    GCU(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    GCU(cudaStreamSynchronize(st));  // v may be a temporary
  }
  return GULON_OK;
}

// ---- assignment launchers -------------------------------------------------------------------
template <int DIM, typename OutT>
int launch_assign_dim(const float *dX, i64 N, i64 ld, const float *cb, const float *off, int K,
                      int dmax, const int32_t *dsubs, int nsub, const int32_t *dfrom, OutT *out,
                      i64 out_stride, cudaStream_t st) {
  constexpr int DP = (DIM + 3) & ~3;
  int KC = K;
  const int kc_max = (46 * 1024) / (4 * (DP + 1));
  if (KC > kc_max) KC = kc_max;
  const size_t smem = (size_t)KC * (DP + 1) * sizeof(float);
  dim3 grid((unsigned)ceil_div(N, ASSIGN_TILE), (unsigned)nsub);
  auto kern = assign_exact_kernel<DIM, OutT>;
  cudaEvent_t ev = g_t_assign.begin(st);
  GLAUNCH(kern, grid, ASSIGN_NT, smem, st, dX, N, ld, cb, off, K, dmax, dsubs, dfrom, out,
          out_stride, KC);
  g_t_assign.end(ev, st);
  return GULON_OK;
}

// ---- tensor-core assignment (tcassign.cuh) ---------------------------------------------------
// What the tensor path needs besides the exact path's arguments.
struct TcCtx {
  DevBuf *blobs = nullptr;        // 256-byte header + one operand blob per window
  int n_windows = 0;
  bool prep = false;              // run tc_prep_kernel for the launch's windows first (centroids changed)
  const int32_t *hsubs = nullptr; // host copy of the launch's window list
  const int32_t *hfrom = nullptr; // host copy of the from[] table
  DevBuf *groups = nullptr;       // scratch for the group table when d_groups is null
  const int32_t *d_groups = nullptr;  // prepared group table (codebooks)
  int n_groups = 0;
  // fused Lloyd pass (tca::tc_assign_kernel<.., true>): fixed-point sums, counts, changed rows
  bool fuse = false;
  const float *scale = nullptr;
  unsigned long long *sums = nullptr;
  int32_t *counts = nullptr, *bad = nullptr, *diff = nullptr;
};

// Groups of up to GRP_MAX windows that are adjacent in the row, so that one TMA box (32 floats from
// the 16-byte boundary below the first window) serves them all.
std::vector<int32_t> build_tc_groups(const int32_t *subs, int n, const int32_t *from, int dim,
                                     int grp_max = tca::GRP_MAX) {
  std::vector<int32_t> g;
  int i = 0;
  while (i < n) {
    const int m0 = from[subs[i]] & 3;
    int cnt = 1;
    while (cnt < grp_max && i + cnt < n && from[subs[i + cnt]] == from[subs[i + cnt - 1]] + dim &&
           m0 + (cnt + 1) * dim <= tca::BOX_COLS)
      cnt++;
    for (int j = 0; j < tca::GRP_MAX; j++) g.push_back(j < cnt ? subs[i + j] : -1);
    i += cnt;
  }
  return g;
}

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                         const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                         const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tensor_map_encode_fn tensor_map_encoder() {
  static tensor_map_encode_fn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return reinterpret_cast<tensor_map_encode_fn>(f);
  }();
  return fn;
}

// the row-major matrix as a TMA tensor: boxes of [box_rows][32 floats], 128-byte swizzle, zero fill
int make_row_map(const float *dX, i64 N, i64 ld, int ncols, int box_rows, CUtensorMap *map) {
  tensor_map_encode_fn enc = tensor_map_encoder();
  GREQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)ncols, (cuuint64_t)N};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)tca::BOX_COLS, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(dX), gdim, gstride, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(GULON_ECUDA, "cuTensorMapEncodeTiled failed (%d) for N=%lld ld=%lld cols=%d", (int)r,
                (long long)N, (long long)ld, ncols);
  return GULON_OK;
}

bool tc_eligible(const float *dX, i64 N, i64 ld, int K, int dim) {
  return K <= tca::TN && 3 * dim + 3 <= tca::KP && (ld % 4) == 0 && ld * 4 < (1LL << 40) &&
         (reinterpret_cast<uintptr_t>(dX) & 15) == 0 && N < (1LL << 31) && tensor_map_encoder() != nullptr;
}

// Approximate scores on tcgen05, exact recheck of the candidate chunks.
template <int DIM, typename OutT, bool FUSE = false>
int launch_assign_tc_dim(const float *dX, i64 N, i64 ld, const float *cb, const float *off, int K,
                         int dmax, const int32_t *dsubs, int nsub, const int32_t *dfrom,
                         const int32_t *ddim, TcCtx &tc, OutT *out, i64 out_stride,
                         cudaStream_t st) {
  DevBuf &scratch = *tc.blobs;
  if (tc.prep) GCHECK(scratch.ensure(256 + (size_t)tc.n_windows * tca::BLOB_BYTES));
  unsigned char *blobs = scratch.as<unsigned char>() + 256;
  int ncols = 0;
  for (int i = 0; i < nsub; i++) ncols = std::max(ncols, tc.hfrom[tc.hsubs[i]] + DIM);
  const int32_t *d_groups = tc.d_groups;
  int n_groups = tc.n_groups;
  if (!d_groups) {
    std::vector<int32_t> g = build_tc_groups(tc.hsubs, nsub, tc.hfrom, DIM, FUSE ? tca::FUSE_GRP : tca::GRP_MAX);
    n_groups = (int)(g.size() / tca::GRP_MAX);
    GCHECK(upload(*tc.groups, g, st));
    d_groups = tc.groups->as<int32_t>();
  }
  unsigned long long *stats = nullptr;
  const bool prof = g_profile.load() != 0;
  DevBuf stats_buf;
  struct Guard {
    DevBuf &b;
    ~Guard() { b.release(); }
  } guard{stats_buf};
  if (prof) {
    GCHECK(stats_buf.ensure(16));
    stats = stats_buf.as<unsigned long long>();
    GCU(cudaMemsetAsync(stats, 0, 16, st));
  }
  if (tc.prep)
    GLAUNCH(tca::tc_prep_kernel, (unsigned)nsub, 256, 0, st, cb, off, dsubs, ddim, K, dmax, blobs);
  auto kern = tca::tc_assign_kernel<DIM, OutT, FUSE>;
  constexpr int kSmem = FUSE ? tca::fuse_smem_bytes(DIM) : tca::SMEM_BYTES;
  GOPTIN(kern, kSmem);
  CUtensorMap map;
  GCHECK(make_row_map(dX, N, ld, ncols, tca::TM, &map));
  tca::Params p;
  p.scale = tc.scale;
  p.sums = tc.sums;
  p.counts = tc.counts;
  p.bad = tc.bad;
  p.diff = tc.diff;
  p.dmax = dmax;
  p.N = N;
  // rows per work unit: at least ~16 units per CTA, at most 8192 rows
  const int sms = sm_count();
  i64 ur = 8192;
  while (ur > 1024 && ceil_div(N, ur) * n_groups < 16LL * sms) ur >>= 1;
  p.unit_rows = (int)ur;
  p.blobs = blobs;
  p.groups = d_groups;
  p.from = dfrom;
  p.n_groups = n_groups;
  p.K = K;
  p.out = out;
  p.out_stride = out_stride;
  p.stats = stats;
  // GULON_TC_DEBUG=1: breadcrumbs of CTA 0 in host-mapped memory, printed when the launch fails
  static int *dbg_host = [] {
    int *h = nullptr;
    const char *e = getenv("GULON_TC_DEBUG");
    if (e && *e == '1' && cudaHostAlloc(&h, 32 * 8 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) h = nullptr;
    return h;
  }();
  p.dbg = nullptr;
  if (dbg_host) {
    memset(dbg_host, 0, 32 * 8 * sizeof(int));
    int *d = nullptr;
    if (cudaHostGetDevicePointer(&d, dbg_host, 0) == cudaSuccess) p.dbg = d;
  }
  const i64 units = ceil_div(N, ur) * n_groups;
  const unsigned grid = (unsigned)std::min<i64>(units, sms);
  cudaEvent_t ev = g_t_assign.begin(st);
  GLAUNCH(kern, grid, tca::NT, kSmem, st, map, p);
  g_t_assign.end(ev, st);
  if (dbg_host) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      std::string m;
      for (int w = 0; w < tca::NT / 32; w++) {
        char b[160];
        snprintf(b, sizeof(b), " w%d[%d %d %d %d %d]", w, dbg_host[w * 8], dbg_host[w * 8 + 1],
                 dbg_host[w * 8 + 2], dbg_host[w * 8 + 3], dbg_host[w * 8 + 4]);
        m += b;
      }
      return fail(GULON_ECUDA, "tc_assign_kernel<%d>: %s; N=%lld groups=%d unit_rows=%d grid=%u;%s", DIM,
                  cudaGetErrorString(e), (long long)N, n_groups, p.unit_rows, grid, m.c_str());
    }
  }
  if (prof) {
    unsigned long long h[2] = {0, 0};
    GCU(cudaMemcpyAsync(h, stats, 16, cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    g_tc_stats[0] += h[0];
    g_tc_stats[2] += (unsigned long long)N * (unsigned long long)nsub;
  }
  return GULON_OK;
}

// One launch for `nsub` sub-quantizers that share the window width `dim`.
template <typename OutT>
int launch_assign(const float *dX, i64 N, i64 ld, const float *cb, const float *off, int K,
                  int dmax, const int32_t *dsubs, int nsub, const int32_t *dfrom,
                  const int32_t *ddim, int dim, TcCtx *tc, OutT *out, i64 out_stride,
                  cudaStream_t st) {
  if (N <= 0 || nsub <= 0) return GULON_OK;
  GREQUIRE(nsub <= 65535, "too many sub-quantizers in one launch (%d)", nsub);
  const long long impl = g_assign_impl.load();
  const bool tc_ok = tc && tc->blobs && tc->hsubs && tc->hfrom && (tc->prep || tc->blobs->p) &&
                     tc_eligible(dX, N, ld, K, dim);
  GREQUIRE(impl != GULON_ASSIGN_TENSOR || tc_ok,
           "assign_impl=tensor needs K <= %d, window width <= %d, a 16-byte aligned matrix and a row "
           "stride that is a multiple of 4 floats (K=%d, width=%d, ld=%lld)", tca::TN,
           (tca::KP - 3) / 3, K, dim, (long long)ld);
  if constexpr (sizeof(OutT) == 4) {
    if (tc && tc->fuse) {
      GREQUIRE(tc_ok, "internal: fused Lloyd pass on a shape the tensor path does not take");
      switch (dim) {
#define GULON_CASE(DD)                                                                                 \
  case DD:                                                                                             \
    return launch_assign_tc_dim<DD, OutT, true>(dX, N, ld, cb, off, K, dmax, dsubs, nsub, dfrom, ddim, \
                                                *tc, out, out_stride, st);
        GULON_CASE(1) GULON_CASE(2) GULON_CASE(3) GULON_CASE(4) GULON_CASE(5) GULON_CASE(6)
        GULON_CASE(7) GULON_CASE(8) GULON_CASE(9) GULON_CASE(10) GULON_CASE(11) GULON_CASE(12)
        GULON_CASE(13) GULON_CASE(14) GULON_CASE(15)
#undef GULON_CASE
        default:
          return fail(GULON_EINVAL, "internal: fused Lloyd pass with window width %d", dim);
      }
    }
  }
  if constexpr (sizeof(OutT) != 2)  // 16-bit codes mean K > 256: never tensor-eligible
  if (tc_ok && (impl == GULON_ASSIGN_TENSOR ||
                (impl == GULON_ASSIGN_AUTO && N >= g_assign_tc_min_rows.load()))) {
    switch (dim) {
#define GULON_CASE(DD)                                                                           \
  case DD:                                                                                       \
    return launch_assign_tc_dim<DD, OutT>(dX, N, ld, cb, off, K, dmax, dsubs, nsub, dfrom, ddim, \
                                          *tc, out, out_stride, st);
      GULON_CASE(1) GULON_CASE(2) GULON_CASE(3) GULON_CASE(4) GULON_CASE(5) GULON_CASE(6)
      GULON_CASE(7) GULON_CASE(8) GULON_CASE(9) GULON_CASE(10) GULON_CASE(11) GULON_CASE(12)
      GULON_CASE(13) GULON_CASE(14) GULON_CASE(15)
#undef GULON_CASE
      default:
        break;
    }
  }
  switch (dim) {
#define GULON_CASE(DD)                                                                           \
  case DD:                                                                                       \
    return launch_assign_dim<DD, OutT>(dX, N, ld, cb, off, K, dmax, dsubs, nsub, dfrom, out,     \
                                       out_stride, st);
    GULON_CASE(1) GULON_CASE(2) GULON_CASE(3) GULON_CASE(4) GULON_CASE(5) GULON_CASE(6)
    GULON_CASE(7) GULON_CASE(8) GULON_CASE(9) GULON_CASE(10) GULON_CASE(11) GULON_CASE(12)
    GULON_CASE(13) GULON_CASE(14) GULON_CASE(15) GULON_CASE(16)
#undef GULON_CASE
    default:
      break;
  }
  const size_t smem = (size_t)AG_KB * dim * sizeof(float);
  GREQUIRE(smem <= 200 * 1024, "window width %d too large for the generic assignment kernel", dim);
  auto kern = assign_generic_kernel<OutT>;
  GOPTIN(kern, 200 * 1024);
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)nsub);
  GLAUNCH(kern, grid, 128, smem, st, dX, N, ld, cb, off, K, dmax, dsubs, dfrom, ddim, out,
          out_stride);
  return GULON_OK;
}

// ---- a set of k-means problems over column windows of one matrix ---------------------------
// Shared by gulon_kmeans_* (one window) and gulon_pq_train (M windows, G/ProductQuantizer.scala:121-148).
struct Problems {
  int n = 0, K = 0, dmax = 0;
  std::vector<int32_t> from, dim;
  DevBuf cb, off, dfrom, ddim, dsubs, diff, sums, counts, part_sum, part_cnt, tile_hist, base,
      order, rows, tc, tc_groups, sums64, bad, absmax, scale, col2win;
  bool fixed = false;              // GULON_UPDATE_SUM runs as the exact fixed-point sum (kupdate.cuh)
  const int32_t *h_cur = nullptr;  // host copy of the window list for_each_width is handing out
  ~Problems() {
    tc.release();
    tc_groups.release();
    sums64.release(); bad.release(); absmax.release(); scale.release(); col2win.release();
    cb.release(); off.release(); dfrom.release(); ddim.release(); dsubs.release(); diff.release();
    sums.release(); counts.release(); part_sum.release(); part_cnt.release();
    tile_hist.release(); base.release(); order.release(); rows.release();
  }
  int setup(int n_, int K_, const int32_t *from_, const int32_t *dim_, cudaStream_t st) {
    n = n_;
    K = K_;
    from.assign(from_, from_ + n);
    dim.assign(dim_, dim_ + n);
    dmax = 1;
    for (int d : dim) dmax = std::max(dmax, d);
    GCHECK(upload(dfrom, from, st));
    GCHECK(upload(ddim, dim, st));
    GCHECK(cb.ensure((size_t)n * K * dmax * sizeof(float)));
    GCHECK(off.ensure((size_t)n * K * sizeof(float)));
    GCHECK(diff.ensure((size_t)n * sizeof(int32_t)));
    GCHECK(counts.ensure((size_t)n * K * sizeof(int32_t)));
    GCHECK(dsubs.ensure((size_t)n * sizeof(int32_t)));
    GCU(cudaMemsetAsync(cb.p, 0, (size_t)n * K * dmax * sizeof(float), st));
    return GULON_OK;
  }
  int offsets(cudaStream_t st) {
    const i64 t = (i64)n * K;
    GLAUNCH(offsets_kernel, (unsigned)ceil_div(t, 256), 256, 0, st, cb.as<float>(),
            ddim.as<int32_t>(), n, K, dmax, off.as<float>());
    return GULON_OK;
  }
  // subs: host list of problem indices to process; grouped by width, one launch per width.
  template <typename F>
  int for_each_width(const std::vector<int32_t> &subs, cudaStream_t st, F f) {
    std::map<int, std::vector<int32_t>> groups;
    for (int32_t s : subs) groups[dim[s]].push_back(s);
    size_t at = 0;
    std::vector<int32_t> flat;
    for (auto &kv : groups) flat.insert(flat.end(), kv.second.begin(), kv.second.end());
    if (flat.empty()) return GULON_OK;
    GCU(cudaMemcpyAsync(dsubs.p, flat.data(), flat.size() * sizeof(int32_t),
                        cudaMemcpyHostToDevice, st));
    GCU(cudaStreamSynchronize(st));
    for (auto &kv : groups) {
      h_cur = kv.second.data();
      GCHECK(f(kv.first, dsubs.as<int32_t>() + at, (int)kv.second.size()));
      at += kv.second.size();
    }
    h_cur = nullptr;
    return GULON_OK;
  }
  int assign(const float *dX, i64 N, i64 ld, const std::vector<int32_t> &subs, int32_t *out,
             i64 out_stride, cudaStream_t st) {
    return for_each_width(subs, st, [&](int w, const int32_t *ds, int ns) {
      TcCtx ctx;
      ctx.blobs = &tc;
      ctx.n_windows = n;
      ctx.prep = true;
      ctx.hsubs = h_cur;
      ctx.hfrom = from.data();
      ctx.groups = &tc_groups;
      return launch_assign<int32_t>(dX, N, ld, cb.as<float>(), off.as<float>(), K, dmax, ds, ns,
                                    dfrom.as<int32_t>(), ddim.as<int32_t>(), w, &ctx, out,
                                    out_stride, st);
    });
  }
  // One fused Lloyd pass over the windows `subs`: assignments (in place: `inout` holds the previous ones),
  // fixed-point sums / counts of the NEW assignments, rows changed per window (pr.diff).  Needs `fixed`.
  bool can_fuse(const float *dX, i64 N, i64 ld) const {
    if (!fixed || g_assign_impl.load() == GULON_ASSIGN_EXACT || !g_train_fused.load()) return false;
    for (int s2 = 0; s2 < n; s2++)
      if (!tc_eligible(dX, N, ld, K, dim[s2])) return false;
    return true;
  }
  int assign_fused(const float *dX, i64 N, i64 ld, const std::vector<int32_t> &subs, int32_t *inout,
                   i64 stride, cudaStream_t st) {
    GCU(cudaMemsetAsync(sums64.p, 0, (size_t)n * K * dmax * sizeof(unsigned long long), st));
    GCU(cudaMemsetAsync(counts.p, 0, (size_t)n * K * sizeof(int32_t), st));
    GCU(cudaMemsetAsync(bad.p, 0, (size_t)n * K * sizeof(int32_t), st));
    GCU(cudaMemsetAsync(diff.p, 0, (size_t)n * sizeof(int32_t), st));
    if (N <= 0) return GULON_OK;
    return for_each_width(subs, st, [&](int w, const int32_t *ds, int ns) {
      TcCtx ctx;
      ctx.blobs = &tc;
      ctx.n_windows = n;
      ctx.prep = true;
      ctx.hsubs = h_cur;
      ctx.hfrom = from.data();
      ctx.groups = &tc_groups;
      ctx.fuse = true;
      ctx.scale = scale.as<float>();
      ctx.sums = sums64.as<unsigned long long>();
      ctx.counts = counts.as<int32_t>();
      ctx.bad = bad.as<int32_t>();
      ctx.diff = diff.as<int32_t>();
      return launch_assign<int32_t>(dX, N, ld, cb.as<float>(), off.as<float>(), K, dmax, ds, ns,
                                    dfrom.as<int32_t>(), ddim.as<int32_t>(), w, &ctx, inout, stride, st);
    });
  }
  // GULON_UPDATE_SUM as an exact fixed-point sum (kupdate.cuh): finds the per-window scale with one
  // pass over the matrix.  Sets `fixed` when the shapes allow it (K <= 256, widths <= 16, TMA-able
  // matrix, and -- sharded -- a host that provides the int64 / max hooks); otherwise the fp32 path
  // below stays in charge.
  int prepare_fixed(const float *dX, i64 N, i64 ld, const gulon_comm_t *comm, cudaStream_t st) {
    fixed = false;
    const bool sharded = comm && comm->world > 1;
    if (sharded && (!comm->allreduce_sum_i64 || !comm->allreduce_max_f32)) return GULON_OK;
    if (K > upd::KMAX || N < 1) return GULON_OK;
    int ncols = 0;
    for (int s = 0; s < n; s++) {
      if (dim[s] > 16) return GULON_OK;
      ncols = std::max(ncols, from[s] + dim[s]);
    }
    if ((ld % 4) != 0 || (reinterpret_cast<uintptr_t>(dX) & 15) != 0 || N >= (1LL << 31) ||
        !tensor_map_encoder() || (size_t)ncols * 4 > 64 * 1024)
      return GULON_OK;
    std::vector<int32_t> c2w((size_t)ncols, -1);
    for (int s = 0; s < n; s++)
      for (int j = 0; j < dim[s]; j++) c2w[from[s] + j] = s;
    GCHECK(upload(col2win, c2w, st));
    GCHECK(absmax.ensure((size_t)n * sizeof(float)));
    GCHECK(scale.ensure((size_t)n * sizeof(float)));
    GCHECK(sums64.ensure((size_t)n * K * dmax * sizeof(unsigned long long)));
    GCHECK(bad.ensure((size_t)n * K * sizeof(int32_t)));
    GCU(cudaMemsetAsync(absmax.p, 0, (size_t)n * sizeof(float), st));
    const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(4LL * sm_count(), ceil_div(N, 8)));
    GLAUNCH(upd::absmax_kernel, grid, 256, (size_t)ncols * 4, st, dX, N, ncols, ld, col2win.as<int32_t>(), n,
            absmax.as<unsigned int>());
    if (sharded)
      GREQUIRE(comm->allreduce_max_f32(comm->user, absmax.as<float>(), (i64)n, st) == 0,
               "allreduce_max_f32 hook failed");
    GLAUNCH(upd::scale_kernel, (unsigned)ceil_div(n, 128), 128, 0, st, absmax.as<unsigned int>(), n,
            scale.as<float>());
    fixed = true;
    return GULON_OK;
  }
  template <int DIM>
  int launch_fixed(const float *dX, i64 N, i64 ld, const int32_t *assign, i64 astride, int ns,
                   cudaStream_t st) {
    std::vector<int32_t> g = build_tc_groups(h_cur, ns, from.data(), DIM);
    const int n_groups = (int)(g.size() / tca::GRP_MAX);
    GCHECK(upload(tc_groups, g, st));
    int ncols = 0;
    for (int i = 0; i < ns; i++) ncols = std::max(ncols, from[h_cur[i]] + DIM);
    CUtensorMap map;
    GCHECK(make_row_map(dX, N, ld, ncols, upd::ROWS, &map));
    auto kern = upd::update_fixed_kernel<DIM>;
    GOPTIN(kern, upd::smem_bytes(DIM));
    upd::Params p;
    p.N = N;
    const int sms = sm_count();
    i64 ur = 65536;
    while (ur > 2048 && ceil_div(N, ur) * n_groups < 8LL * sms) ur >>= 1;
    p.unit_rows = (int)ur;
    p.groups = tc_groups.as<int32_t>();
    p.from = dfrom.as<int32_t>();
    p.n_groups = n_groups;
    p.K = K;
    p.dmax = dmax;
    p.assign = assign;
    p.astride = astride;
    p.scale = scale.as<float>();
    p.sums = sums64.as<unsigned long long>();
    p.counts = counts.as<int32_t>();
    p.bad = bad.as<int32_t>();
    const i64 units = ceil_div(N, ur) * n_groups;
    GLAUNCH(kern, (unsigned)std::min<i64>(units, sms), upd::NT, upd::smem_bytes(DIM), st, map, p);
    return GULON_OK;
  }
  int partial_sums_fixed(const float *dX, i64 N, i64 ld, const int32_t *assign, i64 astride,
                         const std::vector<int32_t> &subs, cudaStream_t st) {
    GCU(cudaMemsetAsync(sums64.p, 0, (size_t)n * K * dmax * sizeof(unsigned long long), st));
    GCU(cudaMemsetAsync(counts.p, 0, (size_t)n * K * sizeof(int32_t), st));
    GCU(cudaMemsetAsync(bad.p, 0, (size_t)n * K * sizeof(int32_t), st));
    if (N <= 0) return GULON_OK;
    return for_each_width(subs, st, [&](int w, const int32_t *, int ns) -> int {
      switch (w) {
#define GULON_CASE(DD) \
  case DD:             \
    return launch_fixed<DD>(dX, N, ld, assign, astride, ns, st);
        GULON_CASE(1) GULON_CASE(2) GULON_CASE(3) GULON_CASE(4) GULON_CASE(5) GULON_CASE(6)
        GULON_CASE(7) GULON_CASE(8) GULON_CASE(9) GULON_CASE(10) GULON_CASE(11) GULON_CASE(12)
        GULON_CASE(13) GULON_CASE(14) GULON_CASE(15) GULON_CASE(16)
#undef GULON_CASE
        default:
          return fail(GULON_EINVAL, "window width %d not supported by the fixed-point update", w);
      }
    });
  }
  // local per-cluster sums / counts (GULON_UPDATE_SUM); the caller all-reduces and finalises
  int partial_sums(const float *dX, i64 N, i64 ld, const int32_t *assign, i64 astride,
                   const std::vector<int32_t> &subs, cudaStream_t st) {
    if (fixed) return partial_sums_fixed(dX, N, ld, assign, astride, subs, st);
    GCHECK(sums.ensure((size_t)n * K * dmax * sizeof(float)));
    return for_each_width(subs, st, [&](int w, const int32_t *ds, int ns) -> int {
      int W = 8;
      auto need = [&](int ww) { return (size_t)ww * K * (w + 1) * 4; };
      while (W > 1 && need(W) > 96 * 1024) W >>= 1;
      GREQUIRE(need(W) <= 200 * 1024,
               "K=%d x width=%d does not fit the shared-memory segmented sum; use "
               "GULON_UPDATE_RUNNING_MEAN", K, w);
      GOPTIN(update_partial_kernel, 200 * 1024);
      int B = (int)std::max<i64>(1, std::min<i64>(ceil_div(296, ns), ceil_div(N, 2048)));
      GCHECK(part_sum.ensure((size_t)ns * B * K * dmax * sizeof(float)));
      GCHECK(part_cnt.ensure((size_t)ns * B * K * sizeof(int32_t)));
      dim3 grid((unsigned)B, (unsigned)ns);
      GLAUNCH(update_partial_kernel, grid, 32 * W, need(W), st, dX, N, ld, assign, astride, ds,
              dfrom.as<int32_t>(), ddim.as<int32_t>(), K, dmax, part_sum.as<float>(),
              part_cnt.as<int32_t>());
      dim3 g2((unsigned)K, (unsigned)ns);
      int bt = 32;
      while (bt < w) bt <<= 1;
      GLAUNCH(update_reduce_kernel, g2, bt, 0, st, part_sum.as<float>(), part_cnt.as<int32_t>(), B,
              ds, ddim.as<int32_t>(), K, dmax, sums.as<float>(), counts.as<int32_t>());
      return GULON_OK;
    });
  }
  int allreduce_sums(const gulon_comm_t *comm, cudaStream_t st) {
    if (fixed) {
      GREQUIRE(comm->allreduce_sum_i64(comm->user, sums64.as<int64_t>(), (i64)n * K * dmax, st) == 0,
               "allreduce_sum_i64 hook failed");
      GREQUIRE(comm->allreduce_sum_i32(comm->user, bad.as<int32_t>(), (i64)n * K, st) == 0,
               "allreduce_sum_i32 hook failed");
    } else {
      GREQUIRE(comm->allreduce_sum_f32(comm->user, sums.as<float>(), (i64)n * K * dmax, st) == 0,
               "allreduce_sum_f32 hook failed");
    }
    GREQUIRE(comm->allreduce_sum_i32(comm->user, counts.as<int32_t>(), (i64)n * K, st) == 0,
             "allreduce_sum_i32 hook failed");
    return GULON_OK;
  }
  int finalize(const int32_t *d_active, cudaStream_t st) {
    const i64 t = (i64)n * K * dmax;
    if (fixed) {
      GLAUNCH(upd::finalize_fixed_kernel, (unsigned)ceil_div(t, 256), 256, 0, st,
              sums64.as<unsigned long long>(), counts.as<int32_t>(), bad.as<int32_t>(), scale.as<float>(),
              ddim.as<int32_t>(), d_active, n, K, dmax, cb.as<float>());
      return GULON_OK;
    }
    GLAUNCH(update_finalize_kernel, (unsigned)ceil_div(t, 256), 256, 0, st, sums.as<float>(),
            counts.as<int32_t>(), d_active, n, K, dmax, cb.as<float>());
    return GULON_OK;
  }
  // literal running mean (GULON_UPDATE_RUNNING_MEAN): stable grouping by cluster, then one
  // thread per (cluster, dimension) replays c <- p + (x - p) / n in row order.
  int running_mean(const float *dX, i64 N, i64 ld, const int32_t *assign, i64 astride,
                   const std::vector<int32_t> &subs, cudaStream_t st) {
    GREQUIRE(N < (1LL << 31), "N too large for the running-mean update");
    const int tiles = (int)std::max<i64>(1, ceil_div(N, RM_TILE));
    GREQUIRE((size_t)4 * K * sizeof(int) <= 200 * 1024, "K=%d too large for the running-mean update", K);
    GOPTIN(rm_scatter_kernel, 200 * 1024);
    GOPTIN(rm_hist_kernel, 200 * 1024);
    return for_each_width(subs, st, [&](int w, const int32_t *ds, int ns) -> int {
      GCHECK(tile_hist.ensure((size_t)ns * tiles * K * sizeof(int32_t)));
      GCHECK(base.ensure((size_t)ns * K * sizeof(int32_t)));
      GCHECK(part_cnt.ensure((size_t)ns * K * sizeof(int32_t)));
      GCHECK(order.ensure((size_t)ns * std::max<i64>(N, 1) * sizeof(int32_t)));
      dim3 g1((unsigned)tiles, (unsigned)ns);
      GLAUNCH(rm_hist_kernel, g1, 256, (size_t)K * sizeof(int), st, assign, astride, N, ds, K,
              tile_hist.as<int32_t>());
      GLAUNCH(rm_scan_kernel, (unsigned)ns, 256, 0, st, tile_hist.as<int32_t>(), tiles, K,
              part_cnt.as<int32_t>(), base.as<int32_t>());
      dim3 g3((unsigned)ceil_div(tiles, 4), (unsigned)ns);
      GLAUNCH(rm_scatter_kernel, g3, 128, (size_t)4 * K * sizeof(int), st, assign, astride, N, ds, K,
              tiles, tile_hist.as<int32_t>(), base.as<int32_t>(), order.as<int32_t>(), N);
      dim3 g4((unsigned)ceil_div((i64)K * w, 128), (unsigned)ns);
      GLAUNCH(rm_chain_kernel, g4, 128, 0, st, dX, ld, ds, dfrom.as<int32_t>(), ddim.as<int32_t>(),
              K, dmax, order.as<int32_t>(), N, part_cnt.as<int32_t>(), base.as<int32_t>(),
              cb.as<float>());
      // counts[sub][k] for callers that want them
      GLAUNCH(scatter_counts_kernel, (unsigned)ceil_div((i64)ns * K, 256), 256, 0, st,
              part_cnt.as<int32_t>(), ds, ns, K, counts.as<int32_t>());
      return GULON_OK;
    });
  }
};

// SummaryStatsBuilder over MathUtils.distance of centroid pairs: KMeans.stepSize,
// G/KMeans.scala:160-168 with G/MathUtils.scala:46-57,85-98 (host side, fp32 as the reference).
void step_size(const float *prev, const float *next, int K, int dim, int ldc, float *mean,
               float *stddev, int32_t *count = nullptr, float *s_out = nullptr) {
  volatile float m = 0.0f, s = 0.0f;
  int n = 0;
  for (int i = 0; i < K; i++) {
    volatile float sum = 0.0f;
    for (int j = 0; j < dim; j++) {
      volatile float dx = next[(i64)i * ldc + j] - prev[(i64)i * ldc + j];
      volatile float sq = dx * dx;
      sum = sum + sq;
    }
    const float x = (float)sqrt((double)sum);
    n += 1;
    const float m0 = m;
    volatile float t = (x - m0) / (float)n;
    m = m0 + t;
    volatile float a = x - m0, b = x - m;
    volatile float ab = a * b;
    s = s + ab;
  }
  *mean = m;
  *stddev = n > 0 ? (float)sqrt((double)(s / (float)n)) : 0.0f;
  if (count) *count = n;
  if (s_out) *s_out = s;
}

// The driver loop of KMeans.computeClusters for `pr.n` windows at once; every window keeps its own
// iteration counter and convergence flag exactly as an independent computeClusters call would.
int train_problems(Problems &pr, gulon_points_t p, const int32_t *seeds, int max_iter,
                   int update_mode, const gulon_comm_t *comm, i64 n_total, i64 row_offset,
                   gulon_progress_fn report, void *user, int quantizer_base, int32_t *n_updates,
                   int32_t *converged, cudaStream_t st) {
  const int n = pr.n, K = pr.K, dmax = pr.dmax;
  const i64 N = p->N;
  const bool sharded = comm && comm->world > 1;
  if (!sharded) {
    n_total = N;
    row_offset = 0;
  }
  GREQUIRE(n_total >= 1 && n_total < (1LL << 31), "k-means needs 1 <= N < 2^31 rows (N=%lld)",
           (long long)n_total);
  GREQUIRE(!sharded || update_mode == GULON_UPDATE_SUM,
           "sharded k-means requires GULON_UPDATE_SUM (the running mean is order-dependent)");
  GREQUIRE(max_iter >= 0, "max_iter must be >= 0");

  // KMeans.init: K global rows sampled with replacement, one Random(seed) per window
  std::vector<i64> rows((size_t)n * K);
  for (int s = 0; s < n; s++) {
    JRandom rng((int64_t)seeds[s]);
    for (int k = 0; k < K; k++) rows[(size_t)s * K + k] = rng.next_int((int32_t)n_total);
  }
  GCHECK(upload(pr.rows, rows, st));
  for (int s = 0; s < n; s++) {
    GLAUNCH(gather_rows_kernel, (unsigned)K, 32, 0, st, p->d, p->ld, pr.from[s], pr.dim[s],
            pr.rows.as<i64>() + (size_t)s * K, row_offset, N,
            pr.cb.as<float>() + (size_t)s * K * dmax, dmax);
  }
  if (sharded) {
    // rows live on exactly one rank; the others contributed zeros
    GREQUIRE(comm->allreduce_sum_f32(comm->user, pr.cb.as<float>(), (i64)n * K * dmax, st) == 0,
             "allreduce_sum_f32 hook failed");
  }
  GCHECK(pr.offsets(st));

  DevBuf a_prev, a_next, d_active;
  struct Guard {
    DevBuf &a, &b, &c;
    ~Guard() { a.release(); b.release(); c.release(); }
  } guard{a_prev, a_next, d_active};
  const i64 astride = std::max<i64>(N, 1);
  GCHECK(d_active.ensure((size_t)n * sizeof(int32_t)));

  std::vector<int32_t> active(n);
  for (int s = 0; s < n; s++) active[s] = s;
  if (update_mode == GULON_UPDATE_SUM && g_update_fixed.load()) GCHECK(pr.prepare_fixed(p->d, N, p->ld, comm, st));
  // Sum mode on tensor-eligible shapes: ONE pass over the matrix per Lloyd iteration.  The pass that
  // assigns with the current centroids also accumulates the fixed-point sums of the NEW assignments (the
  // next update's input) and counts the rows that changed; assignments live in one buffer, updated in
  // place.  Every rank of a sharded run takes the same branch (the shapes are the same everywhere; a
  // rank without rows launches nothing).
  const bool fused = update_mode == GULON_UPDATE_SUM && pr.can_fuse(p->d, std::max<i64>(N, 1), p->ld);
  GCHECK(a_prev.ensure((size_t)n * astride * sizeof(int32_t)));
  if (!fused) GCHECK(a_next.ensure((size_t)n * astride * sizeof(int32_t)));
  if (fused) {
    GCU(cudaMemsetAsync(a_prev.p, 0xFF, (size_t)n * astride * sizeof(int32_t), st));  // "no previous assignment"
    GCHECK(pr.assign_fused(p->d, N, p->ld, active, a_prev.as<int32_t>(), astride, st));
  } else {
    GCHECK(pr.assign(p->d, N, p->ld, active, a_prev.as<int32_t>(), astride, st));
  }

  std::vector<float> h_prev, h_next;
  if (report) {
    h_prev.resize((size_t)n * K * dmax);
    h_next.resize((size_t)n * K * dmax);
    GCU(cudaMemcpyAsync(h_prev.data(), pr.cb.p, h_prev.size() * sizeof(float),
                        cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    for (int s = 0; s < n; s++) {
      gulon_progress_t r = {quantizer_base + s, 0, max_iter, 0.0f, 0.0f, 0, 0, 0.0f};
      report(user, &r);
    }
  }

  std::vector<int32_t> iter(n, 0), conv(n, 0), upd(n, 0), mask(n), h_diff(n);
  while (!active.empty()) {
    for (int s = 0; s < n; s++) mask[s] = 0;
    for (int32_t s : active) mask[s] = 1;
    GCU(cudaMemcpyAsync(d_active.p, mask.data(), (size_t)n * sizeof(int32_t),
                        cudaMemcpyHostToDevice, st));
    // next = fromAssignment(prevAssignments)
    if (update_mode == GULON_UPDATE_RUNNING_MEAN) {
      GCHECK(pr.running_mean(p->d, N, p->ld, a_prev.as<int32_t>(), astride, active, st));
    } else {
      // (fused: the sums of the previous pass's assignments are already there)
      if (!fused) GCHECK(pr.partial_sums(p->d, N, p->ld, a_prev.as<int32_t>(), astride, active, st));
      if (sharded) GCHECK(pr.allreduce_sums(comm, st));
      GCHECK(pr.finalize(d_active.as<int32_t>(), st));
    }
    GCHECK(pr.offsets(st));
    // assignments = next.parAssign(vecs); converged = Arrays.equals(prev, assignments)
    if (fused) {
      GCHECK(pr.assign_fused(p->d, N, p->ld, active, a_prev.as<int32_t>(), astride, st));
    } else {
      GCHECK(pr.assign(p->d, N, p->ld, active, a_next.as<int32_t>(), astride, st));
      GCU(cudaMemsetAsync(pr.diff.p, 0, (size_t)n * sizeof(int32_t), st));
      if (N > 0) {
        GCHECK(pr.for_each_width(active, st, [&](int, const int32_t *ds, int ns) -> int {
          dim3 grid((unsigned)std::min<i64>(ceil_div(N, 1024), 1184), (unsigned)ns);
          GLAUNCH(count_diff_kernel, grid, 256, 0, st, a_prev.as<int32_t>(), a_next.as<int32_t>(), N,
                  astride, ds, pr.diff.as<int32_t>());
          return GULON_OK;
        }));
      }
    }
    if (sharded) {
      GREQUIRE(comm->allreduce_sum_i32(comm->user, pr.diff.as<int32_t>(), (i64)n, st) == 0,
               "allreduce_sum_i32 hook failed");
    }
    GCU(cudaMemcpyAsync(h_diff.data(), pr.diff.p, (size_t)n * sizeof(int32_t),
                        cudaMemcpyDeviceToHost, st));
    if (report)
      GCU(cudaMemcpyAsync(h_next.data(), pr.cb.p, h_next.size() * sizeof(float),
                          cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    // carry the new assignments of the active windows forward (inactive windows are frozen)
    std::vector<int32_t> still;
    for (int32_t s : active) {
      if (!fused)
        GCU(cudaMemcpyAsync(a_prev.as<int32_t>() + (size_t)s * astride,
                            a_next.as<int32_t>() + (size_t)s * astride, (size_t)N * sizeof(int32_t),
                            cudaMemcpyDeviceToDevice, st));
      const int c = h_diff[s] == 0;
      upd[s] += 1;
      conv[s] = c;
      if (report) {
        gulon_progress_t r = {quantizer_base + s, iter[s], max_iter, 0.0f, 0.0f, c, 0, 0.0f};
        step_size(h_prev.data() + (size_t)s * K * dmax, h_next.data() + (size_t)s * K * dmax, K,
                  pr.dim[s], dmax, &r.step_mean, &r.step_stddev, &r.step_count, &r.step_s);
        report(user, &r);
        memcpy(h_prev.data() + (size_t)s * K * dmax, h_next.data() + (size_t)s * K * dmax,
               (size_t)K * dmax * sizeof(float));
      }
      iter[s] = c ? max_iter + 1 : iter[s] + 1;
      if (iter[s] <= max_iter) still.push_back(s);
    }
    active.swap(still);
  }
  GCU(cudaStreamSynchronize(st));
  long long umax = 0, wp = 0;
  for (int s = 0; s < n; s++) {
    if (n_updates) n_updates[s] = upd[s];
    if (converged) converged[s] = conv[s];
    umax = std::max<long long>(umax, upd[s]);
    wp += upd[s] + 1;
  }
  g_train_updates = umax;
  g_train_window_passes = wp;
  return GULON_OK;
}

// Centroid ids must be < K: the reference fails with ArrayIndexOutOfBounds on the first lookup of a
// bad id (G/Index.scala:401-406); here an id >= K would read outside a table.  One max-reduction
// over the planes when an index is created.
template <typename T>
__global__ void max_code_kernel(const T *__restrict__ codes, i64 ps, i64 N, int M,
                                unsigned int *__restrict__ out) {
  unsigned int mx = 0;
  const i64 total = (i64)M * N;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const i64 m = t / N, r = t % N;
    const unsigned int c = codes[m * ps + r];
    mx = c > mx ? c : mx;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned int v = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = v > mx ? v : mx;
  }
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(out, mx);
}
template <typename T>
int validate_codes_dev(const T *dcodes, i64 ps, i64 N, int M, int K) {
  if (N <= 0 || (sizeof(T) == 1 && K >= 256) || (sizeof(T) == 2 && K >= 65536)) return GULON_OK;
  DevBuf mx;
  int rc = [&]() -> int {
    GCHECK(mx.ensure(sizeof(unsigned int)));
    GCU(cudaMemsetAsync(mx.p, 0, sizeof(unsigned int), 0));
    const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(8LL * sm_count(), ceil_div((i64)M * N, 256)));
    GLAUNCH(max_code_kernel<T>, grid, 256, 0, 0, dcodes, ps, N, M, mx.as<unsigned int>());
    unsigned int h = 0;
    GCU(cudaMemcpy(&h, mx.p, sizeof(h), cudaMemcpyDeviceToHost));
    GREQUIRE((int)h < K, "centroid id %u >= K=%d in the code planes", h, K);
    return GULON_OK;
  }();
  mx.release();
  return rc;
}

int make_codebook(int D, int M, int K, gulon_codebook_t *out) {
  GREQUIRE(D >= 1 && M >= 1 && K >= 1, "codebook needs D, M, K >= 1 (D=%d M=%d K=%d)", D, M, K);
  GREQUIRE(M <= D, "more quantizers (%d) than dimensions (%d)", M, D);
  std::unique_ptr<gulon_codebook_s> cb(new gulon_codebook_s);
  cb->D = D;
  cb->M = M;
  cb->K = K;
  cb->from.resize(M);
  cb->dim.resize(M);
  cb->dmax = split_rule(D, M, cb->from.data(), cb->dim.data());
  GCHECK(upload(cb->dfrom, cb->from, 0));
  GCHECK(upload(cb->ddim, cb->dim, 0));
  for (int m = 0; m < M; m++) cb->by_dim[cb->dim[m]].push_back(m);
  for (auto &kv : cb->by_dim) GCHECK(upload(cb->d_by_dim[kv.first], kv.second, 0));
  GCHECK(cb->cb.ensure((size_t)M * K * cb->dmax * sizeof(float)));
  GCHECK(cb->off.ensure((size_t)M * K * sizeof(float)));
  *out = cb.release();
  return GULON_OK;
}

int codebook_offsets(gulon_codebook_t cb, cudaStream_t st) {
  const i64 t = (i64)cb->M * cb->K;
  GLAUNCH(offsets_kernel, (unsigned)ceil_div(t, 256), 256, 0, st, cb->cb.as<float>(),
          cb->ddim.as<int32_t>(), cb->M, cb->K, cb->dmax, cb->off.as<float>());
  // operands of the tensor-core assignment (tcassign.cuh), prepared once per codebook
  bool any = false;
  for (auto &kv : cb->by_dim) any = any || 3 * kv.first + 3 <= tca::KP;
  if (cb->K <= tca::TN && any) {
    GCHECK(cb->tc.ensure(256 + (size_t)cb->M * tca::BLOB_BYTES));
    for (auto &kv : cb->by_dim) {
      if (3 * kv.first + 3 > tca::KP) continue;
      GLAUNCH(tca::tc_prep_kernel, (unsigned)kv.second.size(), 256, 0, st, cb->cb.as<float>(),
              cb->off.as<float>(), cb->d_by_dim[kv.first].as<int32_t>(), cb->ddim.as<int32_t>(),
              cb->K, cb->dmax, cb->tc.as<unsigned char>() + 256);
      std::vector<int32_t> g =
          build_tc_groups(kv.second.data(), (int)kv.second.size(), cb->from.data(), kv.first);
      cb->tc_ngroups[kv.first] = (int)(g.size() / tca::GRP_MAX);
      GCHECK(upload(cb->tc_groups[kv.first], g, st));
    }
  }
  return GULON_OK;
}

template <typename CodeT>
int encode_dev(gulon_codebook_t cb, const float *dX, i64 N, i64 ld, CodeT *dcodes, i64 ps,
               cudaStream_t st) {
  for (auto &kv : cb->by_dim) {
    TcCtx ctx;
    ctx.blobs = &cb->tc;
    ctx.n_windows = cb->M;
    ctx.hsubs = kv.second.data();
    ctx.hfrom = cb->from.data();
    auto it = cb->tc_groups.find(kv.first);
    if (it != cb->tc_groups.end()) {
      ctx.d_groups = it->second.as<int32_t>();
      ctx.n_groups = cb->tc_ngroups[kv.first];
    }
    GCHECK(launch_assign<CodeT>(dX, N, ld, cb->cb.as<float>(), cb->off.as<float>(), cb->K,
                                  cb->dmax, cb->d_by_dim[kv.first].as<int32_t>(),
                                  (int)kv.second.size(), cb->dfrom.as<int32_t>(),
                                  cb->ddim.as<int32_t>(), kv.first,
                                  ctx.d_groups ? &ctx : nullptr, dcodes, ps, st));
  }
  return GULON_OK;
}

int normalize_dev(const float *dX, i64 N, int D, i64 ld, float *out, i64 ldo, cudaStream_t st) {
  if (N <= 0) return GULON_OK;
  GLAUNCH(normalize_rows_kernel, (unsigned)ceil_div(N, 128), 128, 0, st, dX, N, D, ld, out, ldo);
  return GULON_OK;
}

int unpack(const u64 *keys, i64 stride, i64 rows, int k, i64 id_offset, int32_t *ids, float *dists,
           int32_t *sizes, cudaStream_t st) {
  if (rows <= 0) return GULON_OK;
  dim3 block(32, 8);
  GLAUNCH(unpack_keys_kernel, (unsigned)ceil_div(rows, 8), block, 0, st, keys, stride, rows, k,
          id_offset, ids, dists, sizes);
  return GULON_OK;
}

int fill_empty(i64 nq, int k, int32_t *ids, float *dists, int32_t *sizes, cudaStream_t st) {
  if (nq <= 0) return GULON_OK;
  GLAUNCH(fill_empty_kernel, (unsigned)ceil_div(nq * std::max(k, 1), 256), 256, 0, st, nq, k, ids,
          dists, sizes);
  return GULON_OK;
}

// Launches the fused exact scan of [from, until) for G query groups; lists [S][G*4][k].
int fused_lists(gulon_index_t ix, i64 from, i64 until, int G, int k, DevBuf &lists, int *S_out,
                cudaStream_t st, const u64 *floor = nullptr) {
  gulon_codebook_t cb = ix->cb;
  const i64 range = until - from;
  const int Q4 = G * 4;
  GREQUIRE(k <= fscan::KMAX, "fused scan supports k <= %d (k=%d)", fscan::KMAX, k);
  GOPTIN(fscan::fused_scan_kernel, fscan::SMEM_BYTES);
  const int nsm = sm_count();
  int Bs = std::min(G, nsm);
  int S = std::max(1, nsm / Bs);
  // every split should hold at least a few items; boundaries are multiples of 16 rows
  S = (int)std::max<i64>(1, std::min<i64>(S, range / (2 * fscan::R)));
  i64 split_len = round_up(ceil_div(range, S), 16);
  S = (int)ceil_div(range, split_len);
  const size_t nl = (size_t)S * Q4 * k;
  GCHECK(lists.ensure(nl * sizeof(u64)));
  GCU(cudaMemsetAsync(lists.p, 0xFF, nl * sizeof(u64), st));
  fscan::Params prm;
  prm.codes = ix->codes;
  prm.ps = ix->ps;
  prm.from = from;
  prm.until = until;
  prm.split_len = split_len;
  prm.boot = 0;
  prm.lutI = ix->lutI.as<float4>();
  prm.M = cb->M;
  prm.G = G;
  prm.k = k;
  prm.S = S;
  prm.Bs = Bs;
  prm.lists = lists.as<u64>();
  prm.floor = floor;
  cudaEvent_t ev = g_t_scan.begin(st);
  GLAUNCH(fscan::fused_scan_kernel, (unsigned)(S * Bs), fscan::NT, fscan::SMEM_BYTES, st, prm);
  g_t_scan.end(ev, st);
  *S_out = S;
  return GULON_OK;
}

// [S][Q4][k] lists -> one sorted list per query: *keys / *stride (first k entries of each row)
int collapse_lists(DevBuf &lists, int S, int Q4, int k, DevBuf &merged, Selector &sel, u64 **keys,
                   i64 *stride, cudaStream_t st) {
  if (S == 1) {
    *keys = lists.as<u64>();
    *stride = k;
    return GULON_OK;
  }
  const i64 ms = round_up((i64)S * k, SEL_CHUNK);
  GCHECK(merged.ensure((size_t)Q4 * ms * sizeof(u64)));
  dim3 gg((unsigned)ceil_div(ms, 256), (unsigned)Q4);
  GLAUNCH(gather_lists_kernel, gg, 256, 0, st, lists.as<u64>(), S, (i64)Q4, k, merged.as<u64>(), ms);
  return sel.run(merged.as<u64>(), ms, Q4, k, st, keys, stride);
}

// The codes of a row side by side: what the survivor evaluations gather (built once; the code planes of
// an index must not change after its first query).  rowcodes_state: 1 ready, -1 no memory.
int ensure_rowcodes(gulon_index_t ix, cudaStream_t st) {
  if (ix->rowcodes_state != 0 || !ix->codes) return GULON_OK;
  const int M = ix->cb->M;
  ix->rcs = round_up(M, 16);
  if (ix->rowcodes.ensure((size_t)std::max<i64>(ix->N, 1) * (size_t)ix->rcs) == GULON_OK) {
    if (ix->N > 0)
      GLAUNCH(pscan::rowcodes_kernel, (unsigned)ceil_div(ix->N, 256), 256, 0, st, ix->codes, ix->ps, ix->N, M,
              ix->rcs, ix->rowcodes.as<uint8_t>());
    ix->rowcodes_state = 1;
  } else {
    cudaGetLastError();
    ix->rowcodes_state = -1;
  }
  return GULON_OK;
}

// Wide index (256 < K <= 65536): the data the lower-bound scan needs.  Every quantizer's K centroids are
// clustered into 256 groups (a small k-means over the centroid table itself, on the device, seed = the
// quantizer index); the planes of group ids feed the 8-bit bound pass, whose tables hold each group's
// minimum; a row-major copy of the 16-bit ids serves the survivor evaluation.  Built once per index.
int prepare_wide(gulon_index_t ix, cudaStream_t st) {
  if (ix->wide_state != 0) return GULON_OK;
  ix->wide_state = -1;
  gulon_codebook_t cb = ix->cb;
  const int M = cb->M, K = cb->K, dmax = cb->dmax;
  std::vector<uint8_t> h_map((size_t)M * K);
  std::vector<int32_t> h_members((size_t)M * K), h_start((size_t)M * 257), h_a(K);
  DevBuf a;
  struct Guard {
    DevBuf &b;
    ~Guard() { b.release(); }
  } guard{a};
  GCHECK(a.ensure((size_t)K * sizeof(int32_t)));
  for (int m = 0; m < M; m++) {
    gulon_points_s pts;
    pts.d = cb->cb.as<float>() + (size_t)m * K * dmax;
    pts.N = K;
    pts.D = dmax;
    pts.ld = dmax;
    Problems pr;
    const int32_t zero = 0, dm = cb->dim[m], seed = m;
    GCHECK(pr.setup(1, 256, &zero, &dm, st));
    GCHECK(train_problems(pr, &pts, &seed, 4, GULON_UPDATE_RUNNING_MEAN, nullptr, 0, 0, nullptr, nullptr, 0,
                          nullptr, nullptr, st));
    GCHECK(pr.assign(pts.d, K, pts.ld, {0}, a.as<int32_t>(), K, st));
    GCU(cudaMemcpyAsync(h_a.data(), a.p, (size_t)K * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    int cnt[257] = {0};
    for (int c = 0; c < K; c++) {
      const int g = h_a[c] & 255;
      h_map[(size_t)m * K + c] = (uint8_t)g;
      cnt[g + 1]++;
    }
    for (int g = 0; g < 256; g++) cnt[g + 1] += cnt[g];
    for (int g = 0; g <= 256; g++) h_start[(size_t)m * 257 + g] = cnt[g];
    std::vector<int> at(cnt, cnt + 256);
    for (int c = 0; c < K; c++) h_members[(size_t)m * K + at[h_map[(size_t)m * K + c]]++] = c;
  }
  GCHECK(upload(ix->gmap, h_map, st));
  GCHECK(upload(ix->gmembers, h_members, st));
  GCHECK(upload(ix->gstart, h_start, st));
  ix->gps = round_up(std::max<i64>(ix->N, 1), 16);
  ix->rcs16 = round_up(M, 8);
  if (ix->gcodes.ensure((size_t)M * ix->gps) != GULON_OK ||
      ix->rowcodes16.ensure((size_t)std::max<i64>(ix->N, 1) * ix->rcs16 * sizeof(uint16_t)) != GULON_OK) {
    cudaGetLastError();
    return GULON_OK;   // no memory for the copies: the plain-table scan keeps serving this index
  }
  GCU(cudaMemsetAsync(ix->gcodes.p, 0, (size_t)M * ix->gps, st));
  if (ix->N > 0) {
    dim3 gg((unsigned)ceil_div(ix->N, 256), (unsigned)M);
    GLAUNCH(pscan::group_codes_kernel, gg, 256, 0, st, ix->codes16, ix->ps, ix->N, M, K, ix->gmap.as<uint8_t>(),
            ix->gcodes.as<uint8_t>(), ix->gps);
    GLAUNCH(pscan::rowcodes16_kernel, (unsigned)ceil_div(ix->N, 256), 256, 0, st, ix->codes16, ix->ps, ix->N, M,
            ix->rcs16, ix->rowcodes16.as<uint16_t>());
  }
  ix->wide_state = 1;
  return GULON_OK;
}

// k beyond the in-kernel list size (128) runs the fast kernels in passes of <= 128: a pass only admits keys
// greater than the last key of the previous pass (`floor`), so it returns the NEXT 128 of the (distance,
// id) order.  The reference's recall harness defaults reach k = 1000 (G/Tests.scala:53): 8 passes.
constexpr int KCHUNK_MAX = 1024;

// One batch of queries (device, already normalised if the metric asks for it) over [from, until).
// Caller holds ix->mu.
int scan_batch(gulon_index_t ix, const float *dQ, i64 nq, i64 ldq, int k, i64 from, i64 until,
               i64 id_offset, int32_t *d_ids, float *d_dists, int32_t *d_sizes, cudaStream_t st) {
  gulon_codebook_t cb = ix->cb;
  const int M = cb->M, K = cb->K;
  const i64 range = until - from;
  bool wide = false;
  if (ix->codes16) {
    // wide index.  Long ranges and k <= 128: the lower-bound scan over 8-bit group ids (prepare_wide);
    // everything else: plain tables, materialised keys, selection (any k the selector takes)
    long long want = g_scan_impl.load();
    if (want == GULON_SCAN_TENSOR) want = GULON_SCAN_AUTO;
    if ((want == GULON_SCAN_AUTO || want == GULON_SCAN_PRUNED) && k <= pscan::KMAX && M <= 1024 &&
        range >= g_pruned_min_rows.load() && (double)nq * M * K * 4.0 <= 8e9) {
      GCHECK(prepare_wide(ix, st));
      wide = ix->wide_state == 1;
    }
    GREQUIRE(want != GULON_SCAN_PRUNED || wide, "scan_impl = pruned on a wide index needs k <= %d, a range of at least "
             "pruned_min_rows and room for the group planes", pscan::KMAX);
  }
  if (ix->codes16 && !wide) {
    const i64 n_pad = round_up(range, SEL_CHUNK);
    i64 qb = std::min<i64>(g_simple_scratch.load() / (8 * n_pad), (512LL << 20) / ((i64)M * K * 4));
    qb = std::max<i64>(1, std::min<i64>(std::min<i64>(qb, nq), 32768));
    GCHECK(ix->keys.ensure((size_t)qb * n_pad * sizeof(u64)));
    GCHECK(ix->lutW.ensure((size_t)qb * M * K * sizeof(float)));
    for (i64 q0 = 0; q0 < nq; q0 += qb) {
      const i64 nb = std::min<i64>(qb, nq - q0);
      dim3 lg((unsigned)ceil_div(K, 256), (unsigned)M, (unsigned)nb);
      GLAUNCH(lut_wide_kernel, lg, 256, 0, st, dQ + q0 * ldq, ldq, cb->cb.as<float>(),
              cb->dfrom.as<int32_t>(), cb->ddim.as<int32_t>(), M, K, cb->dmax, ix->lutW.as<float>());
      dim3 kg((unsigned)(n_pad / 256), (unsigned)nb);
      GLAUNCH(adc_keys_wide_kernel, kg, 256, 0, st, ix->codes16, ix->ps, from, until,
              ix->lutW.as<float>(), M, K, ix->keys.as<u64>(), n_pad);
      u64 *res;
      i64 rs;
      GCHECK(ix->sel.run(ix->keys.as<u64>(), n_pad, nb, k, st, &res, &rs));
      GCHECK(unpack(res, rs, nb, k, id_offset, d_ids + q0 * k, d_dists + q0 * k,
                    d_sizes ? d_sizes + q0 : nullptr, st));
    }
    return GULON_OK;
  }
  long long impl = wide ? (long long)GULON_SCAN_PRUNED : g_scan_impl.load();
  if (impl == GULON_SCAN_TENSOR) impl = GULON_SCAN_AUTO;   // a batch the tensor scan handed back
  if (impl == GULON_SCAN_AUTO) {
    if (k <= KCHUNK_MAX && M <= 1024 && range >= g_pruned_min_rows.load())
      impl = GULON_SCAN_PRUNED;
    else if (k <= KCHUNK_MAX && range >= g_fused_min_rows.load())
      impl = GULON_SCAN_FUSED;
    else
      impl = GULON_SCAN_SIMPLE;
  }
  g_last_scan = impl;
  // query groups of 4 (float4 tables); the pruned scan works on tiles of QT = 16 queries
  // (8-bit lower-bound fields, used while they keep >= 3 levels per quantizer) or 8 (16-bit fields).
  // The lower bound of the main stage sums ML <= M quantizers (see pscan::qselect_kernel).
  int ML = M;
  if (impl == GULON_SCAN_PRUNED) {
    const long long want = g_pruned_lb.load();
    if (want > 0) {
      ML = (int)std::min<long long>(want, M);
    } else {
      // Feedback by hill climbing on the measured cost: gulon::MlController (mlctl.h).
      gulon::MlController &c = ix->mlc;
      if (c.M != M || c.hint <= 0) c.reset(M);
      // While it is still searching (not parked) the controller waits for the timed launch before it
      // picks the next size: callers enqueue many launches ahead of the GPU, and a search that only
      // advanced when the host happened to fall behind would take hundreds of launches.  The wait
      // ends when the previous main-stage kernel does (the GPU idles for one launch latency); a parked
      // controller never waits.
      if (ix->tm_pending && c.searching()) cudaEventSynchronize(ix->tm_ev1);
      if (ix->tm_pending && cudaEventQuery(ix->tm_ev1) == cudaSuccess) {
        ix->tm_pending = false;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ix->tm_ev0, ix->tm_ev1) == cudaSuccess && ix->tm_pairs > 0) {
          const double cost = (double)ms / (ix->tm_pairs * 1e-9);
          c.on_measurement(ix->tm_ml, cost, ix->tm_shape);
          static const bool dbg = getenv("GULON_DEBUG_ML") != nullptr;
          if (dbg)
            fprintf(stderr, "[gulon ml] measured ml=%d cost=%.4f -> best=%d (%.4f) next=%d dir=%d rev=%d hold=%d\n",
                    ix->tm_ml, cost, c.best, c.best_cost, c.hint, c.dir, c.reversals, c.hold);
        } else {
          cudaGetLastError();
        }
      }
      ML = c.next_launch();
    }
  }
  int FB = (int)g_pruned_bits.load();
  if (FB == 0) FB = (127 / ML >= 3) ? 8 : 16;
  // words per table entry: 4 (most queries per pass) unless the batch is too small to fill a tile
  int W = (int)g_pruned_words.load();
  if (W == 0) W = nq * FB > 64 ? 4 : (nq * FB > 32 ? 2 : 1);
  if (FB == 16 && W == 1) W = 2;  // a tile is made of whole query groups of 4
  if (wide) W = 4;                // the wide kernels are instantiated for full-width entries only
  const int QT = W * 32 / FB;
  const int G = impl == GULON_SCAN_PRUNED ? (QT / 4) * (int)ceil_div(nq, QT) : (int)ceil_div(nq, 4);
  const int Q4 = G * 4;
  GCHECK(ix->lutI.ensure((size_t)G * M * 256 * sizeof(float4)));
  dim3 lg((unsigned)G, (unsigned)M);
  if (wide) {
    // exact tables [nq][M][K], then the groups' minima in the layout of the 8-bit kernels
    GCHECK(ix->lutW.ensure((size_t)nq * M * K * sizeof(float)));
    dim3 wg((unsigned)ceil_div(K, 256), (unsigned)M, (unsigned)nq);
    GLAUNCH(lut_wide_kernel, wg, 256, 0, st, dQ, ldq, cb->cb.as<float>(), cb->dfrom.as<int32_t>(),
            cb->ddim.as<int32_t>(), M, K, cb->dmax, ix->lutW.as<float>());
    GLAUNCH(pscan::lut_group_min_kernel, lg, 256, 0, st, ix->lutW.as<float>(), nq, M, K,
            ix->gmembers.as<int32_t>(), ix->gstart.as<int32_t>(), ix->lutI.as<float4>());
  } else {
    GLAUNCH(lut_build_kernel, lg, 256, 0, st, dQ, ldq, nq, cb->cb.as<float>(),
            cb->dfrom.as<int32_t>(), cb->ddim.as<int32_t>(), M, K, cb->dmax, ix->lutI.as<float4>());
  }
  const int Kb = wide ? 256 : K;   // entries per table of the bound pass

  if (impl == GULON_SCAN_FUSED || impl == GULON_SCAN_PRUNED) {
    const int k_all = k;
    // one pass of a fast kernel: the k best keys greater than `floor` (null: the k best) -> res [Q4][rs]
    auto fast_pass = [&](int k, const u64 *floor, u64 *&res, i64 &rs) -> int {
    if (impl == GULON_SCAN_FUSED) {
      int S = 1;
      GCHECK(fused_lists(ix, from, until, G, k, ix->lists, &S, st, floor));
      GCHECK(collapse_lists(ix->lists, S, Q4, k, ix->merged, ix->sel, &res, &rs, st));
      return GULON_OK;
    }

    if (impl == GULON_SCAN_PRUNED) {
      GREQUIRE(k <= pscan::KMAX, "pruned scan supports k <= %d (k=%d)", pscan::KMAX, k);
      GREQUIRE(M <= pscan::MSEL_MAX, "pruned scan supports M <= %d (M=%d)", pscan::MSEL_MAX, M);
      GREQUIRE(FB == 16 || 127 / ML >= 1, "8-bit pruned scan needs <= 127 quantizers in the bound (%d)", ML);
  #define GULON_PSCAN_VARIANTS(X) X(8, 4) X(8, 2) X(8, 1) X(16, 4) X(16, 2)
  #define GULON_X(FB_, W_)                                                                        \
    if (FB == FB_ && W == W_) {                                                                   \
      auto kern = pscan::pruned_scan_kernel<FB_, W_>;                                             \
      using CfgT = pscan::Cfg<FB_, W_>;                                                           \
      GOPTIN(kern, CfgT::SMEM_BYTES);                                                             \
    }
      GULON_PSCAN_VARIANTS(GULON_X)
  #undef GULON_X
      if (wide) {
        GREQUIRE(W == 4, "internal: wide pruned scan with %d-word entries", W);
        if (FB == 8) {
          auto kern = pscan::pruned_scan_kernel<8, 4, true>;
          GOPTIN(kern, (pscan::Cfg<8, 4>::SMEM_BYTES));
        } else {
          auto kern = pscan::pruned_scan_kernel<16, 4, true>;
          GOPTIN(kern, (pscan::Cfg<16, 4>::SMEM_BYTES));
        }
      }
      const int T = G / (QT / 4);
      const int RI = pscan::NT * (64 / W);  // rows per work item
      g_last_qt = QT;
      // 1. exact scan of the boot rows -> one sorted list per query (its tail is tau0)
      // (measured on the 1M-row shapes c1 / c5: 16384 boot rows beat 65536 by 3-7 %, and the pruned scan
      // beats the exact kernel 2x there; profiles/README.md round 1d)
      i64 boot_want = g_boot_rows.load();
      // whole 8192-row chunks: the exact kernel pays for a chunk's table fills however few rows it holds
      if (boot_want <= 0) boot_want = std::min<i64>(32768, std::max<i64>(8192, (range / 64) & ~8191LL));
      if (wide) boot_want = std::min<i64>(boot_want, 8192);   // materialised keys: 8 bytes per (boot row, query)
      const i64 boot = std::min<i64>(range, std::max<i64>(boot_want, k));
      int Sb = 1;
      u64 *bkeys;
      i64 bstride;
      if (wide) {
        // boot rows of a wide index: exact keys from the plain tables, selection
        const i64 n_pad = round_up(boot, SEL_CHUNK);
        GCHECK(ix->keys.ensure((size_t)Q4 * n_pad * sizeof(u64)));
        GCU(cudaMemsetAsync(ix->keys.p, 0xFF, (size_t)Q4 * n_pad * sizeof(u64), st));   // padding queries: empty
        dim3 kg((unsigned)(n_pad / 256), (unsigned)nq);
        GLAUNCH(adc_keys_wide_kernel, kg, 256, 0, st, ix->codes16, ix->ps, from, from + boot, ix->lutW.as<float>(),
                M, K, ix->keys.as<u64>(), n_pad);
        GCHECK(ix->sel_boot.run(ix->keys.as<u64>(), n_pad, Q4, k, st, &bkeys, &bstride));
      } else {
        GCHECK(fused_lists(ix, from, from + boot, G, k, ix->lists, &Sb, st, floor));
        GCHECK(collapse_lists(ix->lists, Sb, Q4, k, ix->boot_keys, ix->sel_boot, &bkeys, &bstride, st));
      }
      if (boot == range) {
        res = bkeys;
        rs = bstride;
        return GULON_OK;
      }
      GCHECK(ix->mins.ensure((size_t)Q4 * M * sizeof(float)));
      GCHECK(ix->spread.ensure((size_t)Q4 * M * sizeof(float)));
      GCHECK(ix->sufmin.ensure((size_t)Q4 * (M + 1) * sizeof(float)));
      if (!wide && g_pruned_rowcodes.load() != 0) GCHECK(ensure_rowcodes(ix, st));
      GCHECK(ix->qp.ensure((size_t)Q4 * sizeof(pscan::QParam)));
      GCHECK(ix->boot_tail.ensure((size_t)Q4 * sizeof(u64)));
      GCHECK(ix->pstats.ensure(8 * sizeof(unsigned long long)));
      if (!ix->tm_ev0) {
        GCU(cudaEventCreate(&ix->tm_ev0));
        GCU(cudaEventCreate(&ix->tm_ev1));
      }
      // 2. stages over the remaining rows.  A stage quantises the tables against the best lists known
      //    so far (boot, then boot + earlier stages), scans its rows, and merges its lists into them.
      //    The first, short stage uses every quantizer; its result gives the main stage thresholds
      //    tight enough for a lower bound over a SUBSET of the quantizers.
      const i64 pfrom = from + boot, prange = until - pfrom;
      const long long sdiv = g_pruned_stage_div.load();
      i64 stageA = 0;
      if (ML < M && sdiv > 0) {
        // at least 4 items per CTA
        stageA = std::max<i64>(round_up(prange / sdiv, RI), 4 * (i64)RI);
        if (stageA >= prange) stageA = 0;
      }
      const int nsm = sm_count();
      const bool want_stats = g_profile.load() != 0;
      u64 *cur = bkeys;
      i64 cur_stride = bstride;
      for (int stage = stageA > 0 ? 0 : 1; stage < 2; ++stage) {
        const i64 sfrom = stage == 0 ? pfrom : pfrom + stageA;
        const i64 suntil = stage == 0 ? pfrom + stageA : until;
        const i64 srange = suntil - sfrom;
        // the first stage sums every quantizer, or as many as 8-bit fields hold with >= 3 levels each
        const int sML = stage == 1 ? ML : (FB == 8 ? std::min(M, std::max(ML, 42)) : M);
        GCHECK(ix->qlut.ensure((size_t)T * sML * 256 * W * sizeof(uint32_t)));
        GCHECK(ix->msel.ensure((size_t)T * sML * sizeof(int32_t)));
        GLAUNCH(pscan::qparams_kernel, (unsigned)G, 256, 0, st, ix->lutI.as<float4>(), M, Kb, nq, cur,
                cur_stride, k, pscan::t0_units(FB, sML), ix->mins.as<float>(), ix->spread.as<float>(),
                ix->sufmin.as<float>(), ix->qp.as<pscan::QParam>(), ix->boot_tail.as<u64>());
        GLAUNCH(pscan::qselect_kernel, (unsigned)T, 256, 0, st, ix->spread.as<float>(),
                ix->qp.as<pscan::QParam>(), M, sML, QT, ix->msel.as<int32_t>());
        dim3 qg((unsigned)T, (unsigned)sML);
        bool built = false;
  #define GULON_X(FB_, W_)                                                                        \
    if (FB == FB_ && W == W_) {                                                                   \
      auto kern = pscan::qlut_build_kernel<FB_, W_>;                                              \
      GLAUNCH(kern, qg, 256, 0, st, ix->lutI.as<float4>(), ix->mins.as<float>(),                  \
              ix->qp.as<pscan::QParam>(), ix->msel.as<int32_t>(), M, sML, Kb,                     \
              ix->qlut.as<uint32_t>());                                                           \
      built = true;                                                                               \
    }
        GULON_PSCAN_VARIANTS(GULON_X)
  #undef GULON_X
        GREQUIRE(built, "unsupported pruned-scan variant: %d-bit fields, %d words", FB, W);
        int Bs = std::min(T, nsm);
        int S = std::max(1, nsm / Bs);
        S = (int)std::max<i64>(1, std::min<i64>(S, srange / (2 * (i64)RI)));
        const i64 split_len = round_up(ceil_div(srange, S), 16);
        S = (int)ceil_div(srange, split_len);
        const size_t nl = (size_t)S * Q4 * k;
        GCHECK(ix->plists.ensure(nl * sizeof(u64)));
        GCU(cudaMemsetAsync(ix->plists.p, 0xFF, nl * sizeof(u64), st));
        unsigned long long *dstats = ix->pstats.as<unsigned long long>() + 4 * stage;
        GCU(cudaMemsetAsync(dstats, 0, 3 * sizeof(unsigned long long), st));
        pscan::Params prm;
        prm.codes = wide ? ix->gcodes.as<uint8_t>() : ix->codes;
        prm.ps = wide ? ix->gps : ix->ps;
        prm.rowcodes16 = wide ? ix->rowcodes16.as<uint16_t>() : nullptr;
        prm.rcs16 = ix->rcs16;
        prm.lutW = wide ? ix->lutW.as<float>() : nullptr;
        prm.K = K;
        const bool use_rows = !wide && ix->rowcodes_state == 1 && g_pruned_rowcodes.load() != 0;
        prm.rowcodes = use_rows ? ix->rowcodes.as<uint8_t>() : nullptr;
        prm.rcs = ix->rcs;
        prm.sufmin = ix->sufmin.as<float>();
        prm.from = sfrom;
        prm.until = suntil;
        prm.split_len = split_len;
        prm.qlut = ix->qlut.as<uint32_t>();
        prm.msel = ix->msel.as<int32_t>();
        prm.lutI = ix->lutI.as<float4>();
        prm.qp = ix->qp.as<pscan::QParam>();
        prm.boot_tail = ix->boot_tail.as<u64>();
        prm.floor = floor;
        prm.lists = ix->plists.as<u64>();
        prm.stats = dstats;
        prm.nq = nq;
        prm.M = M;
        prm.ML = sML;
        prm.T = T;
        prm.k = k;
        prm.S = S;
        prm.Bs = Bs;
        const bool time_it = stage == 1 && g_pruned_lb.load() == 0 && !ix->tm_pending;
        if (time_it) GCU(cudaEventRecord(ix->tm_ev0, st));
        KernelTimer &ktm = stage == 1 ? g_t_pscan : g_t_pscan_first;
        cudaEvent_t ev = ktm.begin(st);
  #define GULON_X(FB_, W_)                                                                        \
    if (!wide && FB == FB_ && W == W_) {                                                          \
      auto kern = pscan::pruned_scan_kernel<FB_, W_>;                                             \
      using CfgT = pscan::Cfg<FB_, W_>;                                                           \
      GLAUNCH(kern, (unsigned)(S * Bs), pscan::NT, CfgT::SMEM_BYTES, st, prm);                    \
    }
        GULON_PSCAN_VARIANTS(GULON_X)
  #undef GULON_X
        if (wide && FB == 8) {
          auto kern = pscan::pruned_scan_kernel<8, 4, true>;
          GLAUNCH(kern, (unsigned)(S * Bs), pscan::NT, (pscan::Cfg<8, 4>::SMEM_BYTES), st, prm);
        } else if (wide) {
          auto kern = pscan::pruned_scan_kernel<16, 4, true>;
          GLAUNCH(kern, (unsigned)(S * Bs), pscan::NT, (pscan::Cfg<16, 4>::SMEM_BYTES), st, prm);
        }
        ktm.end(ev, st);
        if (time_it) {
          GCU(cudaEventRecord(ix->tm_ev1, st));
          ix->tm_pending = true;
          ix->tm_ml = sML;
          ix->tm_pairs = (double)srange * (double)nq;
          // (the whole pruned range, not this stage's share: the full bound runs without a first stage)
          ix->tm_shape = (long long)T * 1000003LL + (long long)(prange >> 12) + 1;
        }
        if (want_stats) {
          unsigned long long h[3];
          GCU(cudaMemcpyAsync(h, dstats, sizeof(h), cudaMemcpyDeviceToHost, st));
          GCU(cudaStreamSynchronize(st));
          for (int i = 0; i < 3; i++) g_pstats[i] += h[i];
          g_ppairs += (unsigned long long)srange * (unsigned long long)nq;
          if (stage == 1) g_ppairs_main += (unsigned long long)srange * (unsigned long long)nq;
        }
        if (stage == 1) g_last_ml = sML;
        // best lists so far + this stage's split lists -> best lists so far
        DevBuf &mb = stage == 0 ? ix->merged2 : ix->merged;
        Selector &sl = stage == 0 ? ix->sel2 : ix->sel;
        if ((i64)(S + 1) * k <= pscan::MERGE_SMALL_MAX) {
          GCHECK(mb.ensure((size_t)Q4 * k * sizeof(u64)));
          GLAUNCH(pscan::merge_small_kernel, (unsigned)ceil_div(Q4, 4), 128, 0, st, ix->plists.as<u64>(), S,
                  (i64)Q4, k, cur, cur_stride, mb.as<u64>());
          cur = mb.as<u64>();
          cur_stride = k;
        } else {
          const i64 ms = round_up((i64)(S + 1) * k, SEL_CHUNK);
          GCHECK(mb.ensure((size_t)Q4 * ms * sizeof(u64)));
          dim3 gg((unsigned)ceil_div(ms, 256), (unsigned)Q4);
          GLAUNCH(pscan::gather_lists2_kernel, gg, 256, 0, st, ix->plists.as<u64>(), S, (i64)Q4, k, cur,
                  cur_stride, mb.as<u64>(), ms);
          GCHECK(sl.run(mb.as<u64>(), ms, Q4, k, st, &cur, &cur_stride));
        }
      }
  #undef GULON_PSCAN_VARIANTS
      res = cur;
      rs = cur_stride;
      return GULON_OK;
    }

      return fail(GULON_EINVAL, "internal: no fast scan implementation selected");
    };
    u64 *res = nullptr;
    i64 rs = 0;
    if (k_all <= pscan::KMAX) {
      GCHECK(fast_pass(k_all, nullptr, res, rs));
      return unpack(res, rs, nq, k_all, id_offset, d_ids, d_dists, d_sizes, st);
    }
    GREQUIRE(k_all <= KCHUNK_MAX, "the fast scans take k <= %d (k=%d): use scan_impl = simple", KCHUNK_MAX, k_all);
    GCHECK(ix->kacc.ensure((size_t)Q4 * k_all * sizeof(u64)));
    GCHECK(ix->kfloor.ensure((size_t)Q4 * sizeof(u64)));
    GCU(cudaMemsetAsync(ix->kacc.p, 0xFF, (size_t)Q4 * k_all * sizeof(u64), st));
    for (int done = 0; done < k_all; done += pscan::KMAX) {
      const int kk = std::min(pscan::KMAX, k_all - done);
      GCHECK(fast_pass(kk, done ? ix->kfloor.as<u64>() : nullptr, res, rs));
      dim3 block(32, 8);
      GLAUNCH(append_pass_kernel, (unsigned)ceil_div(Q4, 8), block, 0, st, res, rs, (i64)Q4, k_all, done, kk,
              ix->kacc.as<u64>(), ix->kfloor.as<u64>());
    }
    return unpack(ix->kacc.as<u64>(), k_all, nq, k_all, id_offset, d_ids, d_dists, d_sizes, st);
  }

  // simple path: materialise keys for a few query groups at a time, select
  const i64 n_pad = round_up(range, SEL_CHUNK);
  i64 qb = g_simple_scratch.load() / (8 * n_pad);
  qb = std::max<i64>(4, qb & ~3LL);
  qb = std::min<i64>(qb, Q4);
  GCHECK(ix->keys.ensure((size_t)qb * n_pad * sizeof(u64)));
  for (i64 q0 = 0; q0 < nq; q0 += qb) {
    const i64 nb = std::min<i64>(qb, Q4 - q0);       // padded queries in this pass
    const i64 nreal = std::min<i64>(qb, nq - q0);
    for (i64 y0 = 0; y0 < nb; y0 += 32768) {
      dim3 grid((unsigned)(n_pad / 256), (unsigned)std::min<i64>(32768, nb - y0), 1);
      GLAUNCH(adc_keys_kernel, grid, 256, 0, st, ix->codes, ix->ps, from, until, range, range,
              ix->lutI.as<float4>() + (size_t)((q0 + y0) / 4) * M * 256, M, (int)nb,
              ix->keys.as<u64>() + (size_t)y0 * n_pad, n_pad);
    }
    u64 *res;
    i64 rs;
    GCHECK(ix->sel.run(ix->keys.as<u64>(), n_pad, nb, k, st, &res, &rs));
    GCHECK(unpack(res, rs, nreal, k, id_offset, d_ids + q0 * k, d_dists + q0 * k,
                  d_sizes ? d_sizes + q0 : nullptr, st));
  }
  return GULON_OK;
}

// ---- tensor scan (tscan.cuh) --------------------------------------------------------------------
int make_bf16_map(const void *base, i64 rows, int KP, int box_rows, CUtensorMap *map) {
  tensor_map_encode_fn enc = tensor_map_encoder();
  GREQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)KP, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)KP * 2};
  cuuint32_t box[2] = {(cuuint32_t)tscan::KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(GULON_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a bf16 operand of %lld rows x %d", (int)r,
                (long long)rows, KP);
  return GULON_OK;
}

// Can this index ever take the tensor scan?  (shape only; prepare_tensor decides about memory and data)
bool tensor_shape_ok(gulon_index_t ix) {
  gulon_codebook_t cb = ix->cb;
  return ix->codes && !ix->codes16 && cb->K <= 256 && cb->M <= 1024 && tscan::padded_k(cb->D) <= tscan::KP_STREAM_MAX &&
         ix->N < (1LL << 31) && ix->N > 0 && tensor_map_encoder() != nullptr;
}

// Builds the A operand of the filter: every row decoded to bf16 plus its norm terms.  Once per index.
int prepare_tensor(gulon_index_t ix, cudaStream_t st) {
  if (ix->tensor_state != 0) return GULON_OK;
  ix->tensor_state = -1;
  if (!tensor_shape_ok(ix)) return GULON_OK;
  gulon_codebook_t cb = ix->cb;
  const int D = cb->D, M = cb->M, KP = tscan::padded_k(D);
  if ((long long)ix->N * KP * 2 > g_tensor_max_bytes.load()) return GULON_OK;
  GCHECK(ensure_rowcodes(ix, st));
  if (ix->rowcodes_state != 1) return GULON_OK;
  std::vector<int16_t> col(2 * (size_t)D, 0);
  for (int m = 0; m < M; m++)
    for (int t = 0; t < cb->dim[m]; t++) {
      const int j = cb->from[m] + t;
      if (j < 0 || j >= D) return GULON_OK;
      col[j] = (int16_t)m;
      col[D + j] = (int16_t)t;
    }
  GCHECK(upload(ix->tcol, col, st));
  if (ix->xb.ensure((size_t)ix->N * KP * 2) != GULON_OK) {
    cudaGetLastError();
    return GULON_OK;   // no memory for the decoded copy: the pruned scan keeps serving this index
  }
  GCHECK(ix->tflag.ensure(4 * sizeof(int)));
  GCU(cudaMemsetAsync(ix->tflag.p, 0, 4 * sizeof(int), st));
  GLAUNCH(tscan::decode_rows_kernel, (unsigned)ceil_div(ix->N, 8), 256, 0, st, ix->rowcodes.as<uint8_t>(), ix->rcs,
          ix->N, cb->cb.as<float>(), cb->K, cb->dmax, ix->tcol.as<int16_t>(), ix->tcol.as<int16_t>() + D, D, KP,
          ix->xb.as<uint16_t>(), ix->tflag.as<int>());
  int bad = 0;
  GCU(cudaMemcpyAsync(&bad, ix->tflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GCU(cudaStreamSynchronize(st));
  if (bad) {
    ix->xb.release();
    return GULON_OK;   // non-finite or huge rows: not for a bf16 bound
  }
  ix->tensor_kp = KP;
  ix->tensor_state = 1;
  return GULON_OK;
}

// One stage of the filter over rows [sfrom, suntil) for NB blocks of 256 query slots (operand rows in ix->tqb).
int launch_filter(gulon_index_t ix, i64 sfrom, i64 suntil, int NB, unsigned capb, float *dump, cudaStream_t st) {
  const int KP = ix->tensor_kp, sms = sm_count();
  const bool sb = KP > tscan::KP_MAX;   // the query block does not fit shared memory: streamed with the rows (pair kernel only)
  const bool pair = (g_tensor_pair.load() != 0 || sb) && sms >= 2;
  GREQUIRE(pair || !sb, "the tensor scan of an index with D > %d needs CTA pairs", tscan::KP_MAX - tscan::NEXTRA);
  CUtensorMap mapA, mapB;
  GCHECK(make_bf16_map(ix->xb.p, ix->N, KP, tscan::TM, &mapA));
  GCHECK(make_bf16_map(ix->tqb.p, (i64)NB * tscan::TN, KP, pair ? tscan::TN / 2 : tscan::TN, &mapB));
  const i64 len = suntil - sfrom;
  const i64 units = pair ? sms / 2 : sms;               // CTAs, or CTA pairs, that take items
  const i64 tile_rows = pair ? 2 * tscan::TM : tscan::TM;
  // Items are ordered split-major, so the CTAs of the grid work on the same row split (or two adjacent
  // ones) at any time, each for another query block: the operand rows of a split are read from HBM once
  // and served to the other CTAs from L2.  Short stages: ~8 items per unit, at least 64 tiles per item
  // (the item's B operand is reloaded for every item).
  i64 chunk_rows = std::max<i64>(64 * tile_rows, (g_tensor_chunk_bytes.load() / (KP * 2)) / tile_rows * tile_rows);
  i64 S = std::max<i64>(1, (8 * units) / NB);
  S = std::max<i64>(S, ceil_div(len, chunk_rows));
  S = std::max<i64>(1, std::min<i64>(S, len / (64 * tile_rows)));
  const i64 split_len = round_up(ceil_div(len, S), tile_rows);
  S = ceil_div(len, split_len);
  tscan::FParams fp;
  fp.sfrom = sfrom;
  fp.suntil = suntil;
  fp.split_len = split_len;
  fp.S = (int)S;
  fp.NB = NB;
  fp.nkc = (int)ceil_div(KP, tscan::KC);
  fp.ksteps = KP / 16;
  fp.surv = ix->tsurv.as<u64>();
  fp.scount = ix->tscount.as<unsigned>();
  fp.capb = capb;
  fp.flag = ix->tflag.as<int>();
  fp.dump = dump;
  fp.stats = g_profile.load() ? ix->tstats.as<unsigned long long>() : nullptr;
  fp.epi_wait = (int)g_tensor_epi_wait.load();
  const i64 n_items = (i64)NB * S;
  cudaEvent_t ev = g_t_tscan.begin(st);
  if (pair) {
    const unsigned grid = 2u * (unsigned)std::min<i64>(n_items, units);
    if (sb) {
      GOPTIN(tscan::filter2_kernel<true>, tscan::SMEM2_BYTES);
      tscan::filter2_kernel<true><<<grid, tscan::NT, tscan::SMEM2_BYTES, st>>>(mapA, mapB, fp);
    } else {
      GOPTIN(tscan::filter2_kernel<false>, tscan::SMEM2_BYTES);
      tscan::filter2_kernel<false><<<grid, tscan::NT, tscan::SMEM2_BYTES, st>>>(mapA, mapB, fp);
    }
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess && sb) {
      g_t_tscan.end(ev, st);
      return fail(GULON_ECUDA, "tscan::filter2_kernel<streamed B>: %s", cudaGetErrorString(le));
    }
    if (le != cudaSuccess) {
      // a context that cannot co-schedule 2-CTA clusters with 227 KB each (MPS partitions, odd SM masks):
      // the single-CTA form serves from now on
      g_tensor_pair = 0;
      g_t_tscan.end(ev, st);
      return launch_filter(ix, sfrom, suntil, NB, capb, dump, st);
    }
  } else {
    GOPTIN(tscan::filter_kernel, tscan::SMEM_BYTES);
    const unsigned grid = (unsigned)std::min<i64>(n_items, units);
    GLAUNCH(tscan::filter_kernel, grid, tscan::NT, tscan::SMEM_BYTES, st, mapA, mapB, fp);
  }
  g_t_tscan.end(ev, st);
  return GULON_OK;
}

constexpr unsigned TENSOR_CAPQ = 1024;                 // exact candidates per query and stage
constexpr unsigned TENSOR_CAPB = 256 * TENSOR_CAPQ;    // survivors per query block and stage

// One batch of queries over [from, until) by the tensor scan.  *redo is set when the batch has to be
// answered by the pruned scan instead (a query the bound cannot serve, or an overflowing list: adversarial
// data such as millions of identical rows); the outputs are then unspecified.  Caller holds ix->mu.
int tensor_scan_batch(gulon_index_t ix, const float *dQ, i64 nq, i64 ldq, int k, i64 from, i64 until,
                      i64 id_offset, int32_t *d_ids, float *d_dists, int32_t *d_sizes, cudaStream_t st,
                      bool *redo) {
  gulon_codebook_t cb = ix->cb;
  const int M = cb->M, K = cb->K, D = cb->D, KP = ix->tensor_kp;
  const i64 range = until - from;
  *redo = false;
  const int G = (int)ceil_div(nq, 4), Q4 = G * 4;
  GCHECK(ix->lutI.ensure((size_t)G * M * 256 * sizeof(float4)));
  dim3 lg((unsigned)G, (unsigned)M);
  GLAUNCH(lut_build_kernel, lg, 256, 0, st, dQ, ldq, nq, cb->cb.as<float>(), cb->dfrom.as<int32_t>(),
          cb->ddim.as<int32_t>(), M, K, cb->dmax, ix->lutI.as<float4>());
  // 1. the first rows exactly: every query gets k keys and a threshold.  By default the first 256 rows through
  //    tscan::boot_kernel; "tensor_boot_rows" > 0 asks for that many rows through the exact scan kernel instead
  //    (measured, profiles/README.md: -12 ms of a 465 ms step at c2, -12 of 329 at a c4 shard; with k = 100 the
  //    stages grow by 2x only and the five extra ones cost more than the exact kernel's 8192 rows)
  i64 boot = g_tensor_boot.load();
  i64 ratio = g_tensor_ratio.load();
  if (ratio <= 0) ratio = std::max<i64>(2, std::min<i64>(4, (i64)TENSOR_CAPQ / (6 * (i64)k)));
  if (boot <= 0 && ratio < 4) boot = fscan::R;
  u64 *cur;
  i64 cur_stride;
  if (boot <= 0) {
    boot = std::min<i64>(range, tscan::BOOT_ROWS);
    GCHECK(ix->boot_keys.ensure((size_t)Q4 * k * sizeof(u64)));
    GLAUNCH(tscan::boot_kernel, (unsigned)G, tscan::BOOT_ROWS, 0, st, ix->codes, ix->ps, from, (int)boot,
            ix->lutI.as<float4>(), M, k, ix->boot_keys.as<u64>());
    cur = ix->boot_keys.as<u64>();
    cur_stride = k;
  } else {
    boot = std::min<i64>(range, std::max<i64>(boot, k));
    int Sb = 1;
    GCHECK(fused_lists(ix, from, from + boot, G, k, ix->lists, &Sb, st));
    GCHECK(collapse_lists(ix->lists, Sb, Q4, k, ix->boot_keys, ix->sel_boot, &cur, &cur_stride, st));
  }
  if (boot < range) {
    const int NB = (int)ceil_div(nq, tscan::TN);
    const i64 nslots = (i64)NB * tscan::TN;
    const unsigned capq = TENSOR_CAPQ, capb = TENSOR_CAPB;
    static_assert(fscan::KMAX + TENSOR_CAPQ <= tscan::MERGE_SORTN, "merge area too small");
    GCHECK(ix->tqb.ensure((size_t)nslots * KP * 2));
    GCHECK(ix->tsurv.ensure((size_t)NB * capb * sizeof(u64)));
    GCHECK(ix->tscount.ensure((size_t)NB * sizeof(unsigned)));
    GCHECK(ix->tcand.ensure((size_t)nq * capq * sizeof(u64)));
    GCHECK(ix->tccount.ensure((size_t)nq * sizeof(unsigned)));
    GCHECK(ix->tcur0.ensure((size_t)nq * k * sizeof(u64)));
    GCHECK(ix->tcur1.ensure((size_t)nq * k * sizeof(u64)));
    GCHECK(ix->tflag.ensure(4 * sizeof(int)));
    GCHECK(ix->tstats.ensure(8 * sizeof(unsigned long long)));
    GCU(cudaMemsetAsync(ix->tccount.p, 0, (size_t)nq * sizeof(unsigned), st));
    GCU(cudaMemsetAsync(ix->tflag.p, 0, 4 * sizeof(int), st));
    GCU(cudaMemsetAsync(ix->tstats.p, 0, 8 * sizeof(unsigned long long), st));
    GCHECK(ix->tbad.ensure((size_t)nq));
    GCU(cudaMemsetAsync(ix->tbad.p, 0, (size_t)nq, st));
    // a stage scans `ratio` x the rows seen so far: ~ratio * k rows per query beat the threshold it
    // starts with (plus the bound's slack), which the candidate lists must hold several times over
    const auto merge = tscan::merge_kernel;
    GOPTIN(merge, tscan::MERGE_WARPS * tscan::MERGE_SORTN * sizeof(u64));
    const bool want_stats = g_profile.load() != 0;
    i64 done = boot;
    int pp = 0, stages = 0;
    while (done < range) {
      i64 len = std::min<i64>(range - done, round_up(done * ratio, tscan::TM));
      if (range - done - len < len / 4) len = range - done;   // no short last stage
      const i64 sfrom = from + done, suntil = sfrom + len;
      GLAUNCH(tscan::qprep_kernel, (unsigned)ceil_div(nslots, 8), 256, 0, st, dQ, ldq, nq, nslots, D, KP, cur,
              cur_stride, k, ix->tqb.as<uint16_t>(), ix->tflag.as<int>(), ix->tbad.as<uint8_t>());
      GCU(cudaMemsetAsync(ix->tscount.p, 0, (size_t)NB * sizeof(unsigned), st));
      GCHECK(launch_filter(ix, sfrom, suntil, NB, capb, nullptr, st));
      tscan::EParams ep;
      ep.surv = ix->tsurv.as<u64>();
      ep.scount = ix->tscount.as<unsigned>();
      ep.capb = capb;
      ep.rowcodes = ix->rowcodes.as<uint8_t>();
      ep.rcs = ix->rcs;
      ep.lutI = ix->lutI.as<float4>();
      ep.M = M;
      ep.nq = nq;
      ep.cur = cur;
      ep.stride = cur_stride;
      ep.k = k;
      ep.cand = ix->tcand.as<u64>();
      ep.ccount = ix->tccount.as<unsigned>();
      ep.capq = capq;
      ep.flag = ix->tflag.as<int>();
      ep.stats = want_stats ? ix->tstats.as<unsigned long long>() : nullptr;
      // CTAs per query block: with few, many query blocks' tables (7.7 MB each at M = 30) are in flight at a
      // time; with many, they stay in L2 but most threads find no survivor
      dim3 eg((unsigned)std::max<long long>(1, std::min<long long>(1024, g_tensor_eval_blocks.load())), (unsigned)NB);
      GLAUNCH(tscan::eval_kernel, eg, 256, 0, st, ep);
      u64 *nxt = (pp ? ix->tcur1 : ix->tcur0).as<u64>();
      GLAUNCH(merge, (unsigned)ceil_div(nq, tscan::MERGE_WARPS), 32 * tscan::MERGE_WARPS,
              tscan::MERGE_WARPS * tscan::MERGE_SORTN * sizeof(u64), st, cur, cur_stride, nq, k, ix->tcand.as<u64>(),
              ix->tccount.as<unsigned>(), capq, nxt);
      cur = nxt;
      cur_stride = k;
      pp ^= 1;
      done += len;
      stages++;
    }
    int flag = 0;
    unsigned long long hs[4] = {0, 0, 0, 0};
    GCU(cudaMemcpyAsync(&flag, ix->tflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (want_stats) GCU(cudaMemcpyAsync(hs, ix->tstats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    if (want_stats) {
      for (int i = 0; i < 4; i++) g_tstats[i] += hs[i];
      g_tstats[4] += (unsigned long long)(range - boot) * (unsigned long long)nq;
      g_tstats[6] += 1;
      g_tstats[7] += (unsigned long long)stages;
    }
    if (flag & 6) {   // a survivor or candidate list overflowed: the whole batch goes to the pruned scan
      g_tstats[5] += 1;
      *redo = true;
      return GULON_OK;
    }
    GCHECK(unpack(cur, cur_stride, nq, k, id_offset, d_ids, d_dists, d_sizes, st));
    if (flag & 1) {
      // queries the bound cannot serve (non-finite coordinates, norms beyond 1e30): only THEY are answered by
      // the pruned scan; their rows of the output are overwritten
      std::vector<uint8_t> hb((size_t)nq);
      GCU(cudaMemcpyAsync(hb.data(), ix->tbad.p, (size_t)nq, cudaMemcpyDeviceToHost, st));
      GCU(cudaStreamSynchronize(st));
      std::vector<int32_t> idx;
      for (i64 q = 0; q < nq; q++)
        if (hb[q]) idx.push_back((int32_t)q);
      const i64 nb = (i64)idx.size();
      if (nb > 0) {
        g_tscan_bad_queries += nb;
        GCHECK(upload(ix->tbx, idx, st));
        GCHECK(ix->tbq.ensure((size_t)nb * D * sizeof(float)));
        GCHECK(ix->tbi.ensure((size_t)nb * k * sizeof(int32_t)));
        GCHECK(ix->tbd.ensure((size_t)nb * k * sizeof(float)));
        GCHECK(ix->tbs.ensure((size_t)nb * sizeof(int32_t)));
        GLAUNCH(tscan::gather_queries_kernel, (unsigned)nb, 128, 0, st, dQ, ldq, ix->tbx.as<int32_t>(), D,
                ix->tbq.as<float>());
        i64 qb = g_query_batch.load();
        if (qb <= 0) qb = (i64)sm_count() * 16;
        qb = std::min<i64>(round_up(qb, 16), 32768);
        for (i64 q0 = 0; q0 < nb; q0 += qb) {
          const i64 n1 = std::min<i64>(qb, nb - q0);
          GCHECK(scan_batch(ix, ix->tbq.as<float>() + q0 * D, n1, D, k, from, until, id_offset,
                            ix->tbi.as<int32_t>() + q0 * k, ix->tbd.as<float>() + q0 * k, ix->tbs.as<int32_t>() + q0, st));
        }
        GLAUNCH(tscan::scatter_results_kernel, (unsigned)nb, 128, 0, st, ix->tbx.as<int32_t>(), k, ix->tbi.as<int32_t>(),
                ix->tbd.as<float>(), ix->tbs.as<int32_t>(), d_ids, d_dists, d_sizes);
      }
    }
    return GULON_OK;
  }
  return unpack(cur, cur_stride, nq, k, id_offset, d_ids, d_dists, d_sizes, st);
}

int query_dev(gulon_index_t ix, const float *dQ, i64 nq, i64 ldq, int k, i64 from, i64 until,
              int normalize, i64 id_offset, int32_t *d_ids, float *d_dists, int32_t *d_sizes,
              cudaStream_t st) {
  GREQUIRE(from <= until, "expected: from <= until (from=%lld until=%lld)", (long long)from,
           (long long)until);
  GREQUIRE(from >= 0 && until <= ix->N, "expected: from >= 0 && until <= length (until=%lld N=%lld)",
           (long long)until, (long long)ix->N);
  GREQUIRE(k >= 0, "k must be >= 0");
  GREQUIRE(ldq >= ix->cb->D, "query leading dimension %lld < dimension %d", (long long)ldq,
           ix->cb->D);
  if (nq <= 0) return GULON_OK;
  if (k == 0 || until == from) return fill_empty(nq, k, d_ids, d_dists, d_sizes, st);
  std::lock_guard<std::mutex> lock(ix->mu);
  GCHECK(ix->chain.enter(st));
  const int D = ix->cb->D;
  i64 qb = g_query_batch.load();
  if (qb <= 0) qb = (i64)sm_count() * 16;
  qb = std::min<i64>(round_up(qb, 16), 32768);
  // the tensor scan takes large batches on long ranges (tscan.cuh); everything else, and the batches it
  // hands back, go through scan_batch in passes of `qb`
  const long long want = g_scan_impl.load();
  bool tensor = false;
  if (want == GULON_SCAN_TENSOR ||
      (want == GULON_SCAN_AUTO && nq >= g_tensor_min_queries.load() && until - from >= g_tensor_min_rows.load() &&
       (double)nq * (double)(until - from) >= (double)g_tensor_min_pairs.load())) {
    if (k <= fscan::KMAX && tensor_shape_ok(ix)) {
      GCHECK(prepare_tensor(ix, st));
      tensor = ix->tensor_state == 1;
    }
    if (want == GULON_SCAN_TENSOR && !tensor)
      return fail(GULON_EUNSUPPORTED, "scan_impl = tensor needs an 8-bit index with D <= %d, k <= %d, finite rows and "
                  "room for the decoded copy (D=%d K=%d k=%d)", tscan::KP_STREAM_MAX - tscan::NEXTRA, fscan::KMAX, D,
                  ix->cb->K, k);
  }
  i64 tb = nq;
  if (tensor) {
    // queries per pass: all of them while the tables (M KB per query), the candidate lists (8 KB) and the
    // survivor lists (8 KB) fit a few GB; passes of equal size, whole blocks of 256
    tb = g_tensor_query_batch.load();
    if (tb <= 0) tb = std::max<i64>(4096, (12LL << 30) / ((i64)ix->cb->M * 1024 + 16384 + 2 * tscan::KP_STREAM_MAX));
    const i64 passes = ceil_div(nq, tb);
    tb = round_up(ceil_div(nq, passes), tscan::TN);
  }
  const int rc = [&]() -> int {
    for (i64 t0 = 0; t0 < nq; t0 += tb) {
      const i64 tn = std::min<i64>(tb, nq - t0);
      bool redo = !tensor;
      if (tensor) {
        const float *q = dQ + t0 * ldq;
        i64 ql = ldq;
        if (normalize) {
          GCHECK(ix->qbuf.ensure((size_t)tn * D * sizeof(float)));
          GCHECK(normalize_dev(q, tn, D, ldq, ix->qbuf.as<float>(), D, st));
          q = ix->qbuf.as<float>();
          ql = D;
        }
        const int trc = tensor_scan_batch(ix, q, tn, ql, k, from, until, id_offset, d_ids + t0 * k, d_dists + t0 * k,
                                          d_sizes ? d_sizes + t0 : nullptr, st, &redo);
        if (trc == GULON_ENOMEM && want != GULON_SCAN_TENSOR) {
          // no room for this batch's tables / lists: the pruned scan works in passes of a few thousand queries
          cudaGetLastError();
          g_tstats[5] += 1;
          redo = true;
        } else if (trc != GULON_OK) {
          return trc;
        }
      }
      if (!redo) {
        g_last_scan = GULON_SCAN_TENSOR;
        continue;
      }
      for (i64 q0 = t0; q0 < t0 + tn; q0 += qb) {
        const i64 nb = std::min<i64>(qb, t0 + tn - q0);
        const float *q = dQ + q0 * ldq;
        i64 ql = ldq;
        if (normalize) {
          GCHECK(ix->qbuf.ensure((size_t)nb * D * sizeof(float)));
          GCHECK(normalize_dev(q, nb, D, ldq, ix->qbuf.as<float>(), D, st));
          q = ix->qbuf.as<float>();
          ql = D;
        }
        GCHECK(scan_batch(ix, q, nb, ql, k, from, until, id_offset, d_ids + q0 * k, d_dists + q0 * k,
                          d_sizes ? d_sizes + q0 : nullptr, st));
      }
    }
    return GULON_OK;
  }();
  GCHECK(ix->chain.leave(st));  // also after a failure: kernels already enqueued still use the scratch
  return rc;
}


// ---- shard merge (TopKHeap#merge, G/TopKHeap.scala:44-53) -------------------------------------
// Scratch per DEVICE (a thread bound to another GPU must not touch this one's memory), handed from
// stream to stream by a StreamChain.
struct MergeScratch {
  std::mutex mu;
  StreamChain chain;
  DevBuf keys;
  Selector sel;
};
MergeScratch &merge_scratch() {
  static std::mutex mu;
  static std::map<int, std::unique_ptr<MergeScratch>> per_dev;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto &slot = per_dev[dev];
  if (!slot) slot.reset(new MergeScratch);
  return *slot;
}

// S result sets, shard s at element offset s * shard_stride of d_ids / d_dists, each [nq][k].
int merge_shards(const int32_t *d_ids, const float *d_dists, int S, i64 shard_stride, i64 nq, int k,
                 int32_t *d_out_ids, float *d_out_dists, int32_t *d_out_sizes, cudaStream_t st) {
  if (nq == 0) return GULON_OK;
  if ((i64)S * k <= MERGE_WARP_MAX) {   // short lists: one warp per query, no scratch
    int P = 2;
    while (P < S * k) P <<= 1;
    GLAUNCH(merge_results_small_kernel, (unsigned)ceil_div(nq, 4), 128, (size_t)4 * P * sizeof(u64), st, d_ids,
            d_dists, S, shard_stride, k, (i64)nq, P, d_out_ids, d_out_dists, d_out_sizes);
    return GULON_OK;
  }
  MergeScratch &ms = merge_scratch();
  std::lock_guard<std::mutex> lock(ms.mu);
  GCHECK(ms.chain.enter(st));
  const int rc = [&]() -> int {
    const i64 stride = round_up((i64)S * k, SEL_CHUNK);
    const i64 qb = std::max<i64>(1, (256LL << 20) / (stride * 8));
    GCHECK(ms.keys.ensure((size_t)std::min<i64>(qb, nq) * stride * sizeof(u64)));
    for (i64 q0 = 0; q0 < nq; q0 += qb) {
      const i64 nb = std::min<i64>(qb, nq - q0);
      for (i64 y0 = 0; y0 < nb; y0 += 32768) {
        const i64 ny = std::min<i64>(32768, nb - y0);
        dim3 gg((unsigned)ceil_div(stride, 256), (unsigned)ny);
        GLAUNCH(pack_results_kernel, gg, 256, 0, st, d_ids, d_dists, S, shard_stride, k, q0 + y0,
                ms.keys.as<u64>() + (size_t)y0 * stride, stride);
      }
      u64 *res;
      i64 rs;
      GCHECK(ms.sel.run(ms.keys.as<u64>(), stride, nb, k, st, &res, &rs));
      GCHECK(unpack(res, rs, nb, k, 0, d_out_ids + q0 * k, d_out_dists + q0 * k,
                    d_out_sizes ? d_out_sizes + q0 : nullptr, st));
    }
    return GULON_OK;
  }();
  GCHECK(ms.chain.leave(st));
  return rc;
}

// ---- sharded PQIndex#batchQuery ---------------------------------------------------------------
// slices [C][per][2k+1] words (ids | dists | size per query) -> ids / dists / sizes [nq]...
__global__ void scatter_slices_kernel(const int32_t *__restrict__ all, int C, i64 per, int k, i64 nq,
                                      int32_t *__restrict__ ids, float *__restrict__ dists,
                                      int32_t *__restrict__ sizes) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * (k + 1)) return;
  const i64 q = t / (k + 1);
  const int i = (int)(t % (k + 1));
  const i64 c = q / per, ql = q % per;
  const int32_t *blk = all + c * per * (2 * k + 1);
  if (i < k) {
    ids[q * k + i] = blk[ql * k + i];
    dists[q * k + i] = __int_as_float(blk[per * k + ql * k + i]);
  } else if (sizes) {
    sizes[q] = blk[2 * per * k + ql];
  }
}

// The queries of this rank's query group are dq[0, n_mine) (already on the device); `per` is the
// padded slice length every group uses; the whole batch has nq queries.
int sharded_query_body(gulon_index_t ix, const gulon_comm_t *row_comm, const gulon_comm_t *query_comm,
                       const float *dq, i64 n_mine, i64 per, i64 nq, i64 ldq, int k, int normalize,
                       i64 row_offset, int32_t *d_ids, float *d_dists, int32_t *d_sizes, cudaStream_t st) {
  const int R = row_comm ? row_comm->world : 1;
  const int C = query_comm ? query_comm->world : 1;
  {
    // 1. local scan of this rank's row shard: [ids | dists] of the padded slice (global row ids)
    const size_t blk = (size_t)per * k;  // elements per array
    GCHECK(ix->sh_send.ensure(2 * blk * 4));
    int32_t *l_ids = ix->sh_send.as<int32_t>();
    float *l_dists = reinterpret_cast<float *>(l_ids + blk);
    GCHECK(ix->sh_slice.ensure((2 * blk + (size_t)per) * 4));
    int32_t *s_ids = ix->sh_slice.as<int32_t>();
    float *s_dists = reinterpret_cast<float *>(s_ids + blk);
    int32_t *s_sizes = s_ids + 2 * blk;
    const bool direct = C == 1;  // the merged slice IS the result
    int32_t *m_ids = direct ? d_ids : s_ids;
    float *m_dists = direct ? d_dists : s_dists;
    int32_t *m_sizes = direct ? d_sizes : s_sizes;
    int32_t *q_ids = R > 1 ? l_ids : m_ids;
    float *q_dists = R > 1 ? l_dists : m_dists;
    int32_t *q_sizes = R > 1 ? nullptr : m_sizes;
    if (n_mine < per)
      GCHECK(fill_empty(per - n_mine, k, q_ids + n_mine * k, q_dists + n_mine * k,
                        q_sizes ? q_sizes + n_mine : nullptr, st));
    if (n_mine > 0)
      GCHECK(query_dev(ix, dq, n_mine, ldq, k, 0, ix->N, normalize, row_offset, q_ids, q_dists,
                       q_sizes, st));
    // 2. the R shards of the group exchange their candidates; every rank merges (TopKHeap#merge)
    if (R > 1) {
      GREQUIRE(row_comm->allgather, "row_comm has no allgather hook");
      GCHECK(ix->sh_recv.ensure((size_t)R * 2 * blk * 4));
      if (row_comm->allgather(row_comm->user, l_ids, ix->sh_recv.p, (i64)(2 * blk * 4), st) != 0)
        return fail(GULON_ECOMM, "allgather hook failed (row shards)");
      const int32_t *a_ids = ix->sh_recv.as<int32_t>();
      const float *a_dists = reinterpret_cast<const float *>(a_ids + blk);
      GCHECK(merge_shards(a_ids, a_dists, R, (i64)(2 * blk), per, k, m_ids, m_dists, m_sizes, st));
    }
    // 3. the C query groups exchange their slices
    if (C > 1) {
      GREQUIRE(query_comm->allgather, "query_comm has no allgather hook");
      const size_t words = 2 * blk + (size_t)per;
      GCHECK(ix->sh_all.ensure((size_t)C * words * 4));
      if (query_comm->allgather(query_comm->user, s_ids, ix->sh_all.p, (i64)(words * 4), st) != 0)
        return fail(GULON_ECOMM, "allgather hook failed (query groups)");
      const i64 t = nq * (k + 1);
      GLAUNCH(scatter_slices_kernel, (unsigned)ceil_div(t, 256), 256, 0, st, ix->sh_all.as<int32_t>(),
              C, per, k, nq, d_ids, d_dists, d_sizes);
    }
    return GULON_OK;
  }
}

int sharded_query(gulon_index_t ix, const gulon_comm_t *row_comm, const gulon_comm_t *query_comm,
                  const float *dq, i64 n_mine, i64 per, i64 nq, i64 ldq, int k, int normalize,
                  i64 row_offset, int32_t *d_ids, float *d_dists, int32_t *d_sizes, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(ix->mu_sh);
  GCHECK(ix->chain_sh.enter(st));
  const int rc = sharded_query_body(ix, row_comm, query_comm, dq, n_mine, per, nq, ldq, k, normalize,
                                    row_offset, d_ids, d_dists, d_sizes, st);
  GCHECK(ix->chain_sh.leave(st));
  return rc;
}

// exact re-rank of candidate lists against the rows this device owns: [nq][R] global ids -> [nq][k]
int rerank_dev(gulon_points_t p, const float *dq, i64 nq, i64 ldq, const int32_t *d_cand, int R, int k,
               i64 id_lo, int32_t *d_ids, float *d_dists, int32_t *d_sizes, DevBuf &keys, Selector &sel,
               cudaStream_t st) {
  if (nq <= 0) return GULON_OK;
  if (k == 0 || R == 0) return fill_empty(nq, k, d_ids, d_dists, d_sizes, st);
  const int D = p->D;
  GREQUIRE((size_t)D * sizeof(float) <= 48 * 1024, "dimension %d too large for rerank", D);
  if (R <= RERANK_RMAX) {
    int P = 2;
    while (P < R) P <<= 1;
    const size_t smem = (size_t)P * sizeof(u64) + ((size_t)((D + 3) & ~3) + 4 * RERANK_TILE) * sizeof(float);
    GOPTIN(rerank_topk_kernel, 96 * 1024);
    for (i64 q0 = 0; q0 < nq; q0 += 1 << 30) {
      const i64 nb = std::min<i64>(1 << 30, nq - q0);
      GLAUNCH(rerank_topk_kernel, (unsigned)nb, 128, smem, st, p->d, p->ld, D, p->N, id_lo, dq + q0 * ldq,
              ldq, d_cand + q0 * R, R, P, k, d_ids + q0 * k, d_dists + q0 * k, d_sizes ? d_sizes + q0 : nullptr);
    }
    return GULON_OK;
  }
  GREQUIRE(id_lo == 0, "re-rank of more than %d candidates per query is single-shard only", RERANK_RMAX);
  const i64 n_pad = round_up(R, SEL_CHUNK);
  i64 qb = std::max<i64>(1, g_simple_scratch.load() / (8 * n_pad));
  qb = std::min<i64>(std::min<i64>(qb, nq), 32768);
  GCHECK(keys.ensure((size_t)qb * n_pad * sizeof(u64)));
  for (i64 q0 = 0; q0 < nq; q0 += qb) {
    const i64 nb = std::min<i64>(qb, nq - q0);
    dim3 grid((unsigned)(n_pad / 128), (unsigned)nb);
    GLAUNCH(rerank_keys_kernel, grid, 128, (size_t)D * sizeof(float), st, p->d, p->ld, D, p->N, dq + q0 * ldq,
            ldq, d_cand + q0 * R, R, keys.as<u64>(), n_pad);
    u64 *res;
    i64 rs;
    GCHECK(sel.run(keys.as<u64>(), n_pad, nb, k, st, &res, &rs));
    GCHECK(unpack(res, rs, nb, k, 0, d_ids + q0 * k, d_dists + q0 * k, d_sizes ? d_sizes + q0 : nullptr, st));
  }
  return GULON_OK;
}

// PQ top-R over the row-sharded index -> exact re-rank on the owner of each row -> top-k merge.
// dq: the whole batch on the device.
int rerank_query(gulon_index_t ix, gulon_points_t pts, const gulon_comm_t *row_comm, const float *dq, i64 nq,
                 i64 ldq, int k, int R, int normalize, i64 row_offset, int32_t *d_ids, float *d_dists,
                 int32_t *d_sizes, cudaStream_t st) {
  const int Rw = row_comm ? row_comm->world : 1;
  std::lock_guard<std::mutex> lock(ix->mu_sh);
  GCHECK(ix->chain_sh.enter(st));
  const int rc = [&]() -> int {
    const size_t cblk = (size_t)nq * R, kblk = (size_t)nq * k;
    GCHECK(ix->sh_cand.ensure(2 * cblk * 4));
    int32_t *c_ids = ix->sh_cand.as<int32_t>();
    float *c_dists = reinterpret_cast<float *>(c_ids + cblk);
    // 1. the GLOBAL PQ top-R of every query, on every rank (one exchange)
    GCHECK(sharded_query_body(ix, row_comm, nullptr, dq, nq, nq, nq, ldq, R, normalize, row_offset, c_ids,
                              c_dists, nullptr, st));
    const float *q = dq;
    i64 ql = ldq;
    if (normalize) {   // exact distances are taken against the normalised query, as SortedIndex.prepare does
      GCHECK(ix->sh_q.ensure((size_t)nq * pts->D * sizeof(float)));
      GCHECK(normalize_dev(dq, nq, pts->D, ldq, ix->sh_q.as<float>(), pts->D, st));
      q = ix->sh_q.as<float>();
      ql = pts->D;
    }
    // 2. every rank re-ranks the candidates whose raw vectors it holds
    if (Rw == 1)
      return rerank_dev(pts, q, nq, ql, c_ids, R, k, row_offset, d_ids, d_dists, d_sizes, ix->keys, ix->sel2, st);
    GCHECK(ix->sh_send.ensure(2 * kblk * 4));
    int32_t *l_ids = ix->sh_send.as<int32_t>();
    float *l_dists = reinterpret_cast<float *>(l_ids + kblk);
    GCHECK(rerank_dev(pts, q, nq, ql, c_ids, R, k, row_offset, l_ids, l_dists, nullptr, ix->keys, ix->sel2, st));
    // 3. second exchange: the k exact candidates of every rank, merged by (distance, id)
    GCHECK(ix->sh_recv.ensure((size_t)Rw * 2 * kblk * 4));
    if (row_comm->allgather(row_comm->user, l_ids, ix->sh_recv.p, (i64)(2 * kblk * 4), st) != 0)
      return fail(GULON_ECOMM, "allgather hook failed (re-ranked candidates)");
    const int32_t *a_ids = ix->sh_recv.as<int32_t>();
    return merge_shards(a_ids, reinterpret_cast<const float *>(a_ids + kblk), Rw, (i64)(2 * kblk), nq, k, d_ids,
                        d_dists, d_sizes, st);
  }();
  GCHECK(ix->chain_sh.leave(st));
  return rc;
}

int check_sharded_args(gulon_index_t ix, const gulon_comm_t *row_comm, const gulon_comm_t *query_comm,
                       i64 nq, i64 ldq, int k, i64 row_offset) {
  GREQUIRE(ix, "null handle");
  GREQUIRE(!ix->codes16, "the sharded query serves 8-bit indexes");
  GREQUIRE(nq >= 0 && k >= 1, "nq must be >= 0 and k >= 1 (nq=%lld k=%d)", (long long)nq, k);
  GREQUIRE(ldq >= ix->cb->D, "query leading dimension %lld < dimension %d", (long long)ldq, ix->cb->D);
  GREQUIRE(!row_comm || (row_comm->world >= 1 && row_comm->rank >= 0 && row_comm->rank < row_comm->world),
           "bad row_comm rank/world");
  GREQUIRE(!query_comm ||
               (query_comm->world >= 1 && query_comm->rank >= 0 && query_comm->rank < query_comm->world),
           "bad query_comm rank/world");
  GREQUIRE(row_offset >= 0 && row_offset + ix->N < (1LL << 31),
           "global row ids are Int in the reference: row_offset + N must be < 2^31");
  return GULON_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int gulon_version(void) { return GULON_VERSION; }
const char *gulon_last_error(void) { return err_slot().c_str(); }

int gulon_device_count(int32_t *n) {
  GREQUIRE(n, "null argument");
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *n = c;
  return GULON_OK;
}
int gulon_set_device(int32_t device) {
  GCHECK(need_device());
  GCU(cudaSetDevice(device));
  return GULON_OK;
}
int gulon_get_device(int32_t *device) {
  GREQUIRE(device, "null argument");
  GCHECK(need_device());
  int d = 0;
  GCU(cudaGetDevice(&d));
  *device = d;
  return GULON_OK;
}
int gulon_device_sync(void) {
  GCHECK(need_device());
  GCU(cudaDeviceSynchronize());
  return GULON_OK;
}

int gulon_init(const int32_t *devices, int32_t n) {
  GREQUIRE(n >= 0 && (devices || n == 0), "bad device list");
  GCHECK(need_device());
  int count = 0, cur = 0;
  GCU(cudaGetDeviceCount(&count));
  GCU(cudaGetDevice(&cur));
  for (int i = 0; i < n; i++) {
    GREQUIRE(devices[i] >= 0 && devices[i] < count, "device %d does not exist (%d visible)", devices[i], count);
    int major = 0;
    GCU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, devices[i]));
    if (major != 10)
      return fail(GULON_EUNSUPPORTED, "device %d has compute capability %d.x: libgulon_b200 is built for "
                  "sm_100a only", devices[i], major);
    GCU(cudaSetDevice(devices[i]));
    GCU(cudaFree(nullptr));  // create the context now, not inside the first timed call
  }
  GCU(cudaSetDevice(cur));
  return GULON_OK;
}
int gulon_shutdown(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return GULON_OK;
  }
  GCU(cudaDeviceSynchronize());
  return GULON_OK;
}

int gulon_set_option(const char *name, int64_t value) {
  GREQUIRE(name, "null option name");
  std::string s(name);
  if (s == "scan_impl") {
    GREQUIRE(value >= GULON_SCAN_AUTO && value <= GULON_SCAN_TENSOR, "scan_impl must be 0..4");
    g_scan_impl = value;
  } else if (s == "query_batch") {
    GREQUIRE(value >= 0, "query_batch must be >= 0");
    g_query_batch = value;
  } else if (s == "simple_scratch_bytes") {
    GREQUIRE(value >= (1 << 20), "simple_scratch_bytes must be >= 1 MiB");
    g_simple_scratch = value;
  } else if (s == "encode_chunk_rows") {
    GREQUIRE(value >= 1, "encode_chunk_rows must be >= 1");
    g_encode_chunk = value;
  } else if (s == "assign_impl") {
    GREQUIRE(value >= GULON_ASSIGN_AUTO && value <= GULON_ASSIGN_TENSOR, "assign_impl must be 0..2");
    g_assign_impl = value;
  } else if (s == "update_fixed") {
    g_update_fixed = value ? 1 : 0;
  } else if (s == "train_fused") {
    g_train_fused = value ? 1 : 0;
  } else if (s == "assign_tc_min_rows") {
    GREQUIRE(value >= 0, "assign_tc_min_rows must be >= 0");
    g_assign_tc_min_rows = value;
  } else if (s == "profile") {
    g_profile = value ? 1 : 0;
    for (int i = 0; i < 3; i++) g_tc_stats[i] = 0;
    g_t_scan.reset();
    g_t_assign.reset();
    g_t_pscan.reset();
    g_t_pscan_first.reset();
    g_t_tscan.reset();
    for (int i = 0; i < 8; i++) g_tstats[i] = 0;
    g_tscan_bad_queries = 0;
    for (int i = 0; i < 3; i++) g_pstats[i] = 0;
    g_ppairs = 0;
    g_ppairs_main = 0;
  } else if (s == "boot_rows") {
    GREQUIRE(value >= 0, "boot_rows must be >= 0 (0 = auto)");
    g_boot_rows = value;
  } else if (s == "pruned_bits") {
    GREQUIRE(value == 0 || value == 8 || value == 16, "pruned_bits must be 0 (auto), 8 or 16");
    g_pruned_bits = value;
  } else if (s == "pruned_words") {
    GREQUIRE(value == 0 || value == 1 || value == 2 || value == 4,
             "pruned_words must be 0 (auto), 1, 2 or 4");
    g_pruned_words = value;
  } else if (s == "pruned_lb_quantizers") {
    GREQUIRE(value >= 0, "pruned_lb_quantizers must be >= 0 (0 = auto)");
    g_pruned_lb = value;
  } else if (s == "pruned_rowcodes") {
    g_pruned_rowcodes = value != 0;
  } else if (s == "pruned_stage_div") {
    GREQUIRE(value >= 0, "pruned_stage_div must be >= 0 (0 = one stage)");
    g_pruned_stage_div = value;
  } else if (s == "pruned_min_rows") {
    GREQUIRE(value >= 0, "pruned_min_rows must be >= 0");
    g_pruned_min_rows = value;
  } else if (s == "tensor_min_rows" || s == "tensor_min_queries" || s == "tensor_min_pairs" || s == "tensor_query_batch" ||
             s == "tensor_stage_ratio" || s == "tensor_boot_rows" || s == "tensor_max_bytes" ||
             s == "tensor_chunk_bytes" || s == "tensor_pair" || s == "tensor_epi_wait" || s == "tensor_eval_blocks") {
    GREQUIRE(value >= 0, "%s must be >= 0", name);
    GREQUIRE(s != "tensor_epi_wait" || value <= 7, "tensor_epi_wait must be 0..7");
    (s == "tensor_min_rows" ? g_tensor_min_rows : s == "tensor_min_queries" ? g_tensor_min_queries
     : s == "tensor_min_pairs" ? g_tensor_min_pairs
     : s == "tensor_query_batch" ? g_tensor_query_batch : s == "tensor_stage_ratio" ? g_tensor_ratio
     : s == "tensor_boot_rows" ? g_tensor_boot : s == "tensor_chunk_bytes" ? g_tensor_chunk_bytes
     : s == "tensor_pair" ? g_tensor_pair : s == "tensor_epi_wait" ? g_tensor_epi_wait
     : s == "tensor_eval_blocks" ? g_tensor_eval_blocks
     : g_tensor_max_bytes) = value;
  } else if (s == "fused_min_rows") {
    GREQUIRE(value >= 0, "fused_min_rows must be >= 0");
    g_fused_min_rows = value;
  } else {
    return fail(GULON_EINVAL, "unknown option '%s'", name);
  }
  return GULON_OK;
}

int gulon_get_counter(const char *name, int64_t *value) {
  GREQUIRE(name && value, "null argument");
  const std::string s(name);
  if (s == "kernel_launches") {
    *value = launch_counter().load();
    return GULON_OK;
  }
  if (s == "scan_kernel_ns" || s == "scan_kernel_launches") {
    g_t_scan.drain();
    *value = s == "scan_kernel_ns" ? (int64_t)g_t_scan.ns : g_t_scan.launches;
    return GULON_OK;
  }
  if (s == "pscan_kernel_ns" || s == "pscan_kernel_launches") {
    g_t_pscan.drain();
    *value = s == "pscan_kernel_ns" ? (int64_t)g_t_pscan.ns : g_t_pscan.launches;
    return GULON_OK;
  }
  if (s == "pscan_first_kernel_ns" || s == "pscan_first_kernel_launches") {
    g_t_pscan_first.drain();
    *value = s == "pscan_first_kernel_ns" ? (int64_t)g_t_pscan_first.ns : g_t_pscan_first.launches;
    return GULON_OK;
  }
  if (s == "tscan_kernel_ns" || s == "tscan_kernel_launches") {
    g_t_tscan.drain();
    *value = s == "tscan_kernel_ns" ? (int64_t)g_t_tscan.ns : g_t_tscan.launches;
    return GULON_OK;
  }
  {
    static const char *tn[8] = {"tscan_tiles", "tscan_slow_paths", "tscan_survivors", "tscan_candidates",
                                "tscan_pairs", "tscan_fallbacks", "tscan_batches", "tscan_stages"};
    for (int i = 0; i < 8; i++)
      if (s == tn[i]) {
        *value = (int64_t)g_tstats[i].load();
        return GULON_OK;
      }
  }
  if (s == "tscan_handed_back_queries") {
    *value = g_tscan_bad_queries.load();
    return GULON_OK;
  }
  if (s == "scan_last_impl") {
    *value = g_last_scan.load();
    return GULON_OK;
  }
  if (s == "pscan_main_pairs") {
    *value = (int64_t)g_ppairs_main.load();
    return GULON_OK;
  }
  if (s == "train_updates" || s == "train_window_passes") {
    *value = s == "train_updates" ? g_train_updates.load() : g_train_window_passes.load();
    return GULON_OK;
  }
  if (s == "pscan_survivors" || s == "pscan_candidates" || s == "pscan_slow_items") {
    *value = (int64_t)g_pstats[s == "pscan_survivors" ? 0 : s == "pscan_candidates" ? 1 : 2].load();
    return GULON_OK;
  }
  if (s == "pscan_lb_quantizers") {
    *value = g_last_ml.load();
    return GULON_OK;
  }
  if (s == "pscan_qt") {
    *value = (int64_t)g_last_qt.load();
    return GULON_OK;
  }
  if (s == "pscan_pairs") {
    *value = (int64_t)g_ppairs.load();
    return GULON_OK;
  }
  if (s == "assign_tc_pairs" || s == "assign_tc_rows") {
    *value = (int64_t)g_tc_stats[s == "assign_tc_pairs" ? 0 : 2].load();
    return GULON_OK;
  }
  if (s == "assign_kernel_ns" || s == "assign_kernel_launches") {
    g_t_assign.drain();
    *value = s == "assign_kernel_ns" ? (int64_t)g_t_assign.ns : g_t_assign.launches;
    return GULON_OK;
  }
  return fail(GULON_EINVAL, "unknown counter '%s'", name);
}

int gulon_subvectors(int32_t D, int32_t M, int32_t *from, int32_t *dim) {
  GREQUIRE(from && dim, "null argument");
  GREQUIRE(M >= 1 && D >= 0, "subvectors needs M >= 1 and D >= 0 (D=%d M=%d)", D, M);
  return split_rule(D, M, from, dim);
}

// ---- Matrix ---------------------------------------------------------------------------------
int gulon_points_create(const float *X, int64_t N, int32_t D, int64_t ld, gulon_points_t *out) {
  GREQUIRE(out, "null argument");
  GREQUIRE(N >= 0 && D >= 1 && ld >= D, "bad matrix shape N=%lld D=%d ld=%lld", (long long)N, D,
           (long long)ld);
  GREQUIRE(X || N == 0, "null matrix");
  GCHECK(need_device());
  std::unique_ptr<gulon_points_s> p(new gulon_points_s);
  p->N = N;
  p->D = D;
  p->ld = round_up(D, 4);  // 16-byte row pitch: the TMA path of the tensor-core assignment needs it
  p->owned = true;
  if (N > 0) {
    const size_t bytes = (size_t)N * p->ld * sizeof(float);
    cudaError_t e = cudaMalloc(&p->d, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(GULON_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    if (p->ld != D) cudaMemset(p->d, 0, bytes);
    e = copy2d(p->d, (size_t)p->ld * 4, X, (size_t)ld * 4, (size_t)D * 4, (size_t)N,
                     cudaMemcpyHostToDevice, 0, false);
    if (e != cudaSuccess) {
      cudaFree(p->d);
      return fail(GULON_ECUDA, "cudaMemcpy2D failed: %s", cudaGetErrorString(e));
    }
  }
  *out = p.release();
  return GULON_OK;
}

int gulon_points_wrap_dev(const float *dX, int64_t N, int32_t D, int64_t ld, gulon_points_t *out) {
  GREQUIRE(out, "null argument");
  GREQUIRE(N >= 0 && D >= 1 && ld >= D, "bad matrix shape N=%lld D=%d ld=%lld", (long long)N, D,
           (long long)ld);
  GREQUIRE(dX || N == 0, "null matrix");
  GCHECK(need_device());
  gulon_points_s *p = new gulon_points_s;
  p->d = const_cast<float *>(dX);
  p->N = N;
  p->D = D;
  p->ld = ld;
  p->owned = false;
  *out = p;
  return GULON_OK;
}

int gulon_points_info(gulon_points_t p, int64_t *N, int32_t *D, int64_t *ld, const float **dptr) {
  GREQUIRE(p, "null handle");
  if (N) *N = p->N;
  if (D) *D = p->D;
  if (ld) *ld = p->ld;
  if (dptr) *dptr = p->d;
  return GULON_OK;
}

int gulon_points_destroy(gulon_points_t p) {
  if (!p) return GULON_OK;
  if (p->owned && p->d) cudaFree(p->d);
  delete p;
  return GULON_OK;
}

int gulon_points_normalize(gulon_points_t p) {
  GREQUIRE(p, "null handle");
  GCHECK(need_device());
  GCHECK(normalize_dev(p->d, p->N, p->D, p->ld, p->d, p->ld, 0));
  GCU(cudaStreamSynchronize(0));
  return GULON_OK;
}

int gulon_normalize(const float *X, int64_t N, int32_t D, int64_t ld, float *out, int64_t ldo) {
  GREQUIRE(out || N == 0, "null output");
  GREQUIRE(ldo >= D, "output leading dimension %lld < D=%d", (long long)ldo, D);
  gulon_points_t p = nullptr;
  GCHECK(gulon_points_create(X, N, D, ld, &p));
  int rc = gulon_points_normalize(p);
  if (rc == GULON_OK && N > 0) {
    cudaError_t e = copy2d(out, (size_t)ldo * 4, p->d, (size_t)p->ld * 4, (size_t)D * 4,
                                 (size_t)N, cudaMemcpyDeviceToHost, 0, false);
    if (e != cudaSuccess) rc = fail(GULON_ECUDA, "cudaMemcpy2D failed: %s", cudaGetErrorString(e));
  }
  gulon_points_destroy(p);
  return rc;
}

// ---- KMeans ---------------------------------------------------------------------------------
static int check_window(gulon_points_t p, int32_t from, int32_t dim, int32_t K) {
  GREQUIRE(p, "null handle");
  GREQUIRE(from >= 0 && dim >= 1 && (i64)from + dim <= p->D,
           "column window [%d, %d) outside the matrix (D=%d)", from, from + dim, p->D);
  GREQUIRE(K >= 1, "K must be >= 1 (K=%d)", K);
  return GULON_OK;
}

int gulon_kmeans_assign(gulon_points_t p, int32_t from, int32_t dim, const float *C, int32_t K,
                        int64_t batch, int32_t tie_mode, int32_t *out) {
  GCHECK(check_window(p, from, dim, K));
  GREQUIRE(C && (out || p->N == 0), "null argument");
  GREQUIRE(tie_mode == GULON_TIE_LOWEST, "unsupported tie_mode %d (only GULON_TIE_LOWEST)", tie_mode);
  GREQUIRE(batch >= 0, "batch must be >= 0");
  GCHECK(need_device());
  Problems pr;
  GCHECK(pr.setup(1, K, &from, &dim, 0));
  GCU(cudaMemcpy(pr.cb.p, C, (size_t)K * dim * sizeof(float), cudaMemcpyHostToDevice));
  GCHECK(pr.offsets(0));
  if (p->N == 0) return GULON_OK;
  DevBuf a;
  GCHECK(a.ensure((size_t)p->N * sizeof(int32_t)));
  int rc = pr.assign(p->d, p->N, p->ld, {0}, a.as<int32_t>(), p->N, 0);
  if (rc == GULON_OK) {
    cudaError_t e = cudaMemcpy(out, a.p, (size_t)p->N * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(GULON_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  a.release();
  return rc;
}

int gulon_kmeans_update(gulon_points_t p, int32_t from, int32_t dim, const int32_t *assign,
                        int32_t K, int32_t update_mode, float *out_C, int32_t *out_counts) {
  GCHECK(check_window(p, from, dim, K));
  GREQUIRE((assign || p->N == 0) && out_C, "null argument");
  GREQUIRE(update_mode == GULON_UPDATE_RUNNING_MEAN || update_mode == GULON_UPDATE_SUM,
           "unknown update_mode %d", update_mode);
  for (i64 i = 0; i < p->N; i++)
    GREQUIRE(assign[i] >= 0 && assign[i] < K, "assignment %d at row %lld outside [0, %d)",
             assign[i], (long long)i, K);
  GCHECK(need_device());
  Problems pr;
  GCHECK(pr.setup(1, K, &from, &dim, 0));
  DevBuf a;
  GCHECK(a.ensure((size_t)std::max<i64>(p->N, 1) * sizeof(int32_t)));
  int rc = GULON_OK;
  if (p->N > 0) {
    cudaError_t e = cudaMemcpy(a.p, assign, (size_t)p->N * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail(GULON_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  if (rc == GULON_OK) {
    if (update_mode == GULON_UPDATE_RUNNING_MEAN) {
      rc = pr.running_mean(p->d, p->N, p->ld, a.as<int32_t>(), std::max<i64>(p->N, 1), {0}, 0);
    } else {
      if (g_update_fixed.load()) rc = pr.prepare_fixed(p->d, p->N, p->ld, nullptr, 0);
      if (rc == GULON_OK)
        rc = pr.partial_sums(p->d, p->N, p->ld, a.as<int32_t>(), std::max<i64>(p->N, 1), {0}, 0);
      if (rc == GULON_OK) rc = pr.finalize(nullptr, 0);
    }
  }
  if (rc == GULON_OK) {
    cudaError_t e = cudaMemcpy(out_C, pr.cb.p, (size_t)K * dim * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_counts)
      e = cudaMemcpy(out_counts, pr.counts.p, (size_t)K * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(GULON_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  a.release();
  return rc;
}

int gulon_kmeans_init(gulon_points_t p, int32_t from, int32_t dim, int32_t K, int32_t seed,
                      float *out_C, int32_t *out_rows) {
  GCHECK(check_window(p, from, dim, K));
  GREQUIRE(out_C, "null argument");
  GREQUIRE(p->N >= 1 && p->N < (1LL << 31), "KMeans.init needs 1 <= N < 2^31 rows");
  GCHECK(need_device());
  std::vector<i64> rows(K);
  JRandom rng((int64_t)seed);
  for (int k = 0; k < K; k++) {
    rows[k] = rng.next_int((int32_t)p->N);
    if (out_rows) out_rows[k] = (int32_t)rows[k];
  }
  DevBuf dr, dc;
  int rc = upload(dr, rows, 0);
  if (rc == GULON_OK) rc = dc.ensure((size_t)K * dim * sizeof(float));
  if (rc == GULON_OK) {
    rc = [&]() -> int {
      GLAUNCH(gather_rows_kernel, (unsigned)K, 32, 0, 0, p->d, p->ld, from, dim, dr.as<i64>(),
              (i64)0, p->N, dc.as<float>(), dim);
      GCU(cudaMemcpy(out_C, dc.p, (size_t)K * dim * sizeof(float), cudaMemcpyDeviceToHost));
      return GULON_OK;
    }();
  }
  dr.release();
  dc.release();
  return rc;
}

int gulon_kmeans_train(gulon_points_t p, int32_t from, int32_t dim, int32_t K, int32_t max_iter,
                       int32_t seed, int32_t tie_mode, int32_t update_mode,
                       const gulon_comm_t *comm, int64_t n_total, int64_t row_offset,
                       gulon_progress_fn report, void *user, float *out_C, int32_t *n_updates,
                       int32_t *converged) {
  GCHECK(check_window(p, from, dim, K));
  GREQUIRE(out_C, "null argument");
  GREQUIRE(tie_mode == GULON_TIE_LOWEST, "unsupported tie_mode %d (only GULON_TIE_LOWEST)", tie_mode);
  GREQUIRE(update_mode == GULON_UPDATE_RUNNING_MEAN || update_mode == GULON_UPDATE_SUM,
           "unknown update_mode %d", update_mode);
  GCHECK(need_device());
  Problems pr;
  GCHECK(pr.setup(1, K, &from, &dim, 0));
  GCHECK(train_problems(pr, p, &seed, max_iter, update_mode, comm, n_total, row_offset, report,
                        user, 0, n_updates, converged, 0));
  GCU(cudaMemcpy(out_C, pr.cb.p, (size_t)K * dim * sizeof(float), cudaMemcpyDeviceToHost));
  return GULON_OK;
}

// ---- ProductQuantizer -----------------------------------------------------------------------
int gulon_pq_train(gulon_points_t p, int32_t M, int32_t K, int32_t max_iter, int32_t tie_mode,
                   int32_t update_mode, const gulon_comm_t *comm, int64_t n_total,
                   int64_t row_offset, gulon_progress_fn report, void *user,
                   gulon_codebook_t *out) {
  GREQUIRE(p && out, "null argument");
  GREQUIRE(tie_mode == GULON_TIE_LOWEST, "unsupported tie_mode %d (only GULON_TIE_LOWEST)", tie_mode);
  GREQUIRE(update_mode == GULON_UPDATE_RUNNING_MEAN || update_mode == GULON_UPDATE_SUM,
           "unknown update_mode %d", update_mode);
  GCHECK(need_device());
  gulon_codebook_t cb = nullptr;
  GCHECK(make_codebook(p->D, M, K, &cb));
  std::unique_ptr<gulon_codebook_s> hold(cb);
  Problems pr;
  GCHECK(pr.setup(M, K, cb->from.data(), cb->dim.data(), 0));
  std::vector<int32_t> seeds(M);
  for (int m = 0; m < M; m++) seeds[m] = m;  // G/ProductQuantizer.scala:139
  GCHECK(train_problems(pr, p, seeds.data(), max_iter, update_mode, comm, n_total, row_offset,
                        report, user, 0, nullptr, nullptr, 0));
  GCU(cudaMemcpy(cb->cb.p, pr.cb.p, (size_t)M * K * cb->dmax * sizeof(float),
                 cudaMemcpyDeviceToDevice));
  GCHECK(codebook_offsets(cb, 0));
  GCU(cudaStreamSynchronize(0));
  *out = hold.release();
  return GULON_OK;
}

int gulon_codebook_create(int32_t D, int32_t M, int32_t K, const float *centroids,
                          gulon_codebook_t *out) {
  GREQUIRE(centroids && out, "null argument");
  GCHECK(need_device());
  gulon_codebook_t cb = nullptr;
  GCHECK(make_codebook(D, M, K, &cb));
  std::unique_ptr<gulon_codebook_s> hold(cb);
  GCU(cudaMemcpy(cb->cb.p, centroids, (size_t)M * K * cb->dmax * sizeof(float),
                 cudaMemcpyHostToDevice));
  GCHECK(codebook_offsets(cb, 0));
  GCU(cudaStreamSynchronize(0));
  *out = hold.release();
  return GULON_OK;
}

int gulon_codebook_info(gulon_codebook_t cb, int32_t *D, int32_t *M, int32_t *K, int32_t *dmax) {
  GREQUIRE(cb, "null handle");
  if (D) *D = cb->D;
  if (M) *M = cb->M;
  if (K) *K = cb->K;
  if (dmax) *dmax = cb->dmax;
  return GULON_OK;
}

int gulon_codebook_export(gulon_codebook_t cb, float *centroids) {
  GREQUIRE(cb && centroids, "null argument");
  GCU(cudaMemcpy(centroids, cb->cb.p, (size_t)cb->M * cb->K * cb->dmax * sizeof(float),
                 cudaMemcpyDeviceToHost));
  return GULON_OK;
}

int gulon_codebook_destroy(gulon_codebook_t cb) {
  delete cb;
  return GULON_OK;
}

int gulon_pq_encode_dev(gulon_codebook_t cb, const float *dX, int64_t N, int64_t ld,
                        int32_t tie_mode, uint8_t *dcodes, int64_t plane_stride, void *stream) {
  GREQUIRE(cb, "null handle");
  GREQUIRE(tie_mode == GULON_TIE_LOWEST, "unsupported tie_mode %d (only GULON_TIE_LOWEST)", tie_mode);
  GREQUIRE(cb->K <= 256, "Coder8 needs K <= 256 (K=%d)", cb->K);
  GREQUIRE(N >= 0 && ld >= cb->D && plane_stride >= N, "bad shapes N=%lld ld=%lld stride=%lld",
           (long long)N, (long long)ld, (long long)plane_stride);
  GREQUIRE((dX && dcodes) || N == 0, "null argument");
  GCHECK(need_device());
  return encode_dev<uint8_t>(cb, dX, N, ld, dcodes, plane_stride, (cudaStream_t)stream);
}

}  // extern "C"
template <typename CodeT>
static int pq_encode_host(gulon_codebook_t cb, const float *X, int64_t N, int64_t ld, int32_t tie_mode,
                          CodeT *codes) {
  GREQUIRE(cb, "null handle");
  GREQUIRE(tie_mode == GULON_TIE_LOWEST, "unsupported tie_mode %d (only GULON_TIE_LOWEST)", tie_mode);
  GREQUIRE(sizeof(CodeT) == 2 ? cb->K <= 65536 : cb->K <= 256,
           "K=%d does not fit %d-byte codes", cb->K, (int)sizeof(CodeT));
  GREQUIRE(N >= 0 && ld >= cb->D, "bad shapes N=%lld ld=%lld", (long long)N, (long long)ld);
  GREQUIRE((X && codes) || N == 0, "null argument");
  GCHECK(need_device());
  if (N == 0) return GULON_OK;
  // Double-buffered row chunks: copy chunk c+1 to HBM while chunk c is encoded.
  const int D = cb->D, M = cb->M;
  const int Dp = (int)round_up(D, 4);  // 16-byte row pitch of the staged chunks (TMA)
  const i64 chunk = std::min<i64>(N, g_encode_chunk.load());
  const i64 cps = round_up(chunk, 16);
  std::lock_guard<std::mutex> lock(cb->mu);   // the staging buffers live in the codebook handle
  float *dx[2];
  CodeT *dc[2];
  for (int i = 0; i < 2; i++) {
    if (!cb->enc_st[i]) GCU(cudaStreamCreateWithFlags(&cb->enc_st[i], cudaStreamNonBlocking));
    const bool grew = cb->enc_x[i].cap < (size_t)chunk * Dp * sizeof(float);
    GCHECK(cb->enc_x[i].ensure((size_t)chunk * Dp * sizeof(float)));
    GCHECK(cb->enc_c[i].ensure((size_t)M * cps * sizeof(CodeT)));
    dx[i] = cb->enc_x[i].template as<float>();
    dc[i] = cb->enc_c[i].template as<CodeT>();
    if (Dp != D && grew) GCU(cudaMemset(dx[i], 0, cb->enc_x[i].cap));   // the pad columns stay zero
  }
  int b = 0;
  for (i64 r0 = 0; r0 < N; r0 += chunk, b ^= 1) {
    const i64 n = std::min<i64>(chunk, N - r0);
    GCU(copy2d(dx[b], (size_t)Dp * 4, X + r0 * ld, (size_t)ld * 4, (size_t)D * 4, (size_t)n,
               cudaMemcpyHostToDevice, cb->enc_st[b], true));
    GCHECK(encode_dev(cb, dx[b], n, Dp, dc[b], cps, cb->enc_st[b]));
    GCU(copy2d(codes + r0, (size_t)N * sizeof(CodeT), dc[b], (size_t)cps * sizeof(CodeT),
               (size_t)n * sizeof(CodeT), (size_t)M, cudaMemcpyDeviceToHost, cb->enc_st[b], true));
    // buffer pair b is reused two chunks later on the same stream, so reuse is stream-ordered
  }
  GCU(cudaStreamSynchronize(cb->enc_st[0]));
  GCU(cudaStreamSynchronize(cb->enc_st[1]));
  return GULON_OK;
}

extern "C" {
int gulon_pq_encode(gulon_codebook_t cb, const float *X, int64_t N, int64_t ld, int32_t tie_mode,
                    uint8_t *codes) {
  return pq_encode_host<uint8_t>(cb, X, N, ld, tie_mode, codes);
}

int gulon_pq_encode16(gulon_codebook_t cb, const float *X, int64_t N, int64_t ld, int32_t tie_mode,
                      uint16_t *codes) {
  return pq_encode_host<uint16_t>(cb, X, N, ld, tie_mode, codes);
}

int gulon_pq_encode16_dev(gulon_codebook_t cb, const float *dX, int64_t N, int64_t ld,
                          int32_t tie_mode, uint16_t *dcodes, int64_t plane_stride, void *stream) {
  GREQUIRE(cb, "null handle");
  GREQUIRE(tie_mode == GULON_TIE_LOWEST, "unsupported tie_mode %d (only GULON_TIE_LOWEST)", tie_mode);
  GREQUIRE(cb->K <= 65536, "too many clusters: %d", cb->K);
  GREQUIRE(N >= 0 && ld >= cb->D && plane_stride >= N, "bad shapes N=%lld ld=%lld stride=%lld",
           (long long)N, (long long)ld, (long long)plane_stride);
  GREQUIRE((dX && dcodes) || N == 0, "null argument");
  GCHECK(need_device());
  return encode_dev<uint16_t>(cb, dX, N, ld, dcodes, plane_stride, (cudaStream_t)stream);
}

}  // extern "C"
template <typename CodeT>
static int pq_decode_host(gulon_codebook_t cb, const CodeT *codes, int64_t N, int64_t plane_stride,
                          float *out, int64_t ldo) {
  GREQUIRE(cb, "null handle");
  GREQUIRE(N >= 0 && plane_stride >= N && ldo >= cb->D, "bad shapes");
  GREQUIRE((codes && out) || N == 0, "null argument");
  GREQUIRE(sizeof(CodeT) == 2 ? cb->K <= 65536 : cb->K <= 256,
           "K=%d does not fit %d-byte codes", cb->K, (int)sizeof(CodeT));
  GCHECK(need_device());
  if (N == 0) return GULON_OK;
  for (int m = 0; m < cb->M; m++)
    for (i64 i = 0; i < N; i++)
      GREQUIRE(codes[(i64)m * plane_stride + i] < cb->K, "code %d >= K=%d (plane %d row %lld)",
               (int)codes[(i64)m * plane_stride + i], cb->K, m, (long long)i);
  DevBuf dc, dout;
  int rc = [&]() -> int {
    GCHECK(dc.ensure((size_t)cb->M * N * sizeof(CodeT)));
    GCHECK(dout.ensure((size_t)N * cb->D * sizeof(float)));
    GCU(copy2d(dc.p, (size_t)N * sizeof(CodeT), codes, (size_t)plane_stride * sizeof(CodeT),
                     (size_t)N * sizeof(CodeT), (size_t)cb->M, cudaMemcpyHostToDevice, 0, false));
    dim3 block(32, 8);
    GLAUNCH(decode_kernel<CodeT>, (unsigned)ceil_div(N, 8), block, 0, 0, dc.as<CodeT>(), N, N,
            cb->cb.as<float>(), cb->dfrom.as<int32_t>(), cb->ddim.as<int32_t>(), cb->M, cb->K,
            cb->dmax, dout.as<float>(), (i64)cb->D);
    GCU(copy2d(out, (size_t)ldo * 4, dout.p, (size_t)cb->D * 4, (size_t)cb->D * 4, (size_t)N,
                     cudaMemcpyDeviceToHost, 0, false));
    return GULON_OK;
  }();
  dc.release();
  dout.release();
  return rc;
}

extern "C" {
int gulon_pq_decode(gulon_codebook_t cb, const uint8_t *codes, int64_t N, int64_t plane_stride,
                    float *out, int64_t ldo) {
  return pq_decode_host<uint8_t>(cb, codes, N, plane_stride, out, ldo);
}

int gulon_pq_decode16(gulon_codebook_t cb, const uint16_t *codes, int64_t N, int64_t plane_stride,
                      float *out, int64_t ldo) {
  return pq_decode_host<uint16_t>(cb, codes, N, plane_stride, out, ldo);
}

// ---- Index.PQIndex ----------------------------------------------------------------------------
int gulon_index_create(gulon_codebook_t cb, const uint8_t *codes, int64_t N, int64_t plane_stride,
                       gulon_index_t *out) {
  GREQUIRE(cb && out, "null argument");
  GREQUIRE(N >= 0 && plane_stride >= N, "bad shapes N=%lld stride=%lld", (long long)N,
           (long long)plane_stride);
  GREQUIRE(codes || N == 0, "null codes");
  GREQUIRE(cb->K <= 256, "Coder8 needs K <= 256 (K=%d)", cb->K);
  GREQUIRE(N < (1LL << 31), "row ids are Int in the reference: N must be < 2^31");
  GCHECK(need_device());
  std::unique_ptr<gulon_index_s> ix(new gulon_index_s);
  ix->cb = cb;
  ix->N = N;
  ix->ps = round_up(std::max<i64>(N, 1), 16);
  ix->owned = true;
  uint8_t *d = nullptr;
  cudaError_t e = cudaMalloc(&d, (size_t)cb->M * ix->ps);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GULON_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", (size_t)cb->M * ix->ps,
                cudaGetErrorString(e));
  }
  ix->codes = d;
  GCU(cudaMemset(d, 0, (size_t)cb->M * ix->ps));
  if (N > 0)
    GCU(copy2d(d, (size_t)ix->ps, codes, (size_t)plane_stride, (size_t)N, (size_t)cb->M,
                     cudaMemcpyHostToDevice, 0, false));
  GCHECK(validate_codes_dev<uint8_t>(ix->codes, ix->ps, N, cb->M, cb->K));
  *out = ix.release();
  return GULON_OK;
}

int gulon_index_create_dev(gulon_codebook_t cb, const uint8_t *dcodes, int64_t N,
                           int64_t plane_stride, gulon_index_t *out) {
  GREQUIRE(cb && out, "null argument");
  GREQUIRE(N >= 0 && plane_stride >= round_up(N, 16) && plane_stride % 16 == 0,
           "plane_stride (%lld) must be a multiple of 16 and >= N rounded up to 16 (N=%lld)",
           (long long)plane_stride, (long long)N);
  GREQUIRE(dcodes || N == 0, "null codes");
  GREQUIRE(((uintptr_t)dcodes & 15) == 0, "device codes must be 16-byte aligned");
  GREQUIRE(cb->K <= 256, "Coder8 needs K <= 256 (K=%d)", cb->K);
  GREQUIRE(N < (1LL << 31), "row ids are Int in the reference: N must be < 2^31");
  GCHECK(need_device());
  GCHECK(validate_codes_dev<uint8_t>(dcodes, plane_stride, N, cb->M, cb->K));
  gulon_index_s *ix = new gulon_index_s;
  ix->cb = cb;
  ix->codes = dcodes;
  ix->N = N;
  ix->ps = plane_stride;
  ix->owned = false;
  *out = ix;
  return GULON_OK;
}

/* Wide index: 16-bit centroid ids (256 < K <= 65536).  plane_stride in elements. */
int gulon_index_create16(gulon_codebook_t cb, const uint16_t *codes, int64_t N, int64_t plane_stride,
                         gulon_index_t *out) {
  GREQUIRE(cb && out, "null argument");
  GREQUIRE(N >= 0 && plane_stride >= N, "bad shapes N=%lld stride=%lld", (long long)N,
           (long long)plane_stride);
  GREQUIRE(codes || N == 0, "null codes");
  GREQUIRE(cb->K <= 65536, "too many clusters: %d", cb->K);
  GREQUIRE(N < (1LL << 31), "row ids are Int in the reference: N must be < 2^31");
  GCHECK(need_device());
  for (int m = 0; m < cb->M; m++)
    for (i64 i = 0; i < N; i++)
      GREQUIRE(codes[(i64)m * plane_stride + i] < cb->K, "code %d >= K=%d (plane %d row %lld)",
               (int)codes[(i64)m * plane_stride + i], cb->K, m, (long long)i);
  std::unique_ptr<gulon_index_s> ix(new gulon_index_s);
  ix->cb = cb;
  ix->N = N;
  ix->ps = round_up(std::max<i64>(N, 1), 16);
  ix->owned = true;
  uint16_t *d = nullptr;
  const size_t bytes = (size_t)cb->M * ix->ps * sizeof(uint16_t);
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GULON_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  ix->codes16 = d;
  GCU(cudaMemset(d, 0, bytes));
  if (N > 0)
    GCU(copy2d(d, (size_t)ix->ps * 2, codes, (size_t)plane_stride * 2, (size_t)N * 2,
                     (size_t)cb->M, cudaMemcpyHostToDevice, 0, false));
  *out = ix.release();
  return GULON_OK;
}

/* Adopts device codes (borrowed): uint16 [M][plane_stride], ids < K. */
int gulon_index_create16_dev(gulon_codebook_t cb, const uint16_t *dcodes, int64_t N,
                             int64_t plane_stride, gulon_index_t *out) {
  GREQUIRE(cb && out, "null argument");
  GREQUIRE(N >= 0 && plane_stride >= N, "bad shapes N=%lld stride=%lld", (long long)N,
           (long long)plane_stride);
  GREQUIRE(dcodes || N == 0, "null codes");
  GREQUIRE(cb->K <= 65536, "too many clusters: %d", cb->K);
  GREQUIRE(N < (1LL << 31), "row ids are Int in the reference: N must be < 2^31");
  GCHECK(need_device());
  GCHECK(validate_codes_dev<uint16_t>(dcodes, plane_stride, N, cb->M, cb->K));
  gulon_index_s *ix = new gulon_index_s;
  ix->cb = cb;
  ix->codes16 = dcodes;
  ix->N = N;
  ix->ps = plane_stride;
  ix->owned = false;
  *out = ix;
  return GULON_OK;
}

int gulon_index_info(gulon_index_t ix, int64_t *N, int32_t *M, int32_t *K, int32_t *D) {
  GREQUIRE(ix, "null handle");
  if (N) *N = ix->N;
  if (M) *M = ix->cb->M;
  if (K) *K = ix->cb->K;
  if (D) *D = ix->cb->D;
  return GULON_OK;
}

int gulon_index_destroy(gulon_index_t ix) {
  delete ix;
  return GULON_OK;
}

int gulon_prepare_query(gulon_codebook_t cb, const float *queries, int64_t nq, int64_t ldq,
                        float *lut) {
  GREQUIRE(cb, "null handle");
  GREQUIRE(nq >= 0 && ldq >= cb->D, "bad query shape nq=%lld ldq=%lld", (long long)nq,
           (long long)ldq);
  GREQUIRE((queries && lut) || nq == 0, "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  std::lock_guard<std::mutex> lock(cb->mu);
  const int M = cb->M, K = cb->K, D = cb->D;
  if (K > 256) {
    const i64 qw = std::max<i64>(1, std::min<i64>(4096, (256LL << 20) / ((i64)M * K * 4)));
    GCHECK(cb->scratch_q.ensure((size_t)qw * D * sizeof(float) + (size_t)qw * M * K * sizeof(float)));
    float *dq = cb->scratch_q.as<float>();
    float *dl = dq + (size_t)qw * D;
    for (i64 q0 = 0; q0 < nq; q0 += qw) {
      const i64 nb = std::min<i64>(qw, nq - q0);
      GCU(copy2d(dq, (size_t)D * 4, queries + q0 * ldq, (size_t)ldq * 4, (size_t)D * 4,
                            (size_t)nb, cudaMemcpyHostToDevice, 0, true));
      dim3 lg((unsigned)ceil_div(K, 256), (unsigned)M, (unsigned)nb);
      GLAUNCH(lut_wide_kernel, lg, 256, 0, 0, dq, (i64)D, cb->cb.as<float>(), cb->dfrom.as<int32_t>(),
              cb->ddim.as<int32_t>(), M, K, cb->dmax, dl);
      GCU(cudaMemcpyAsync(lut + q0 * M * K, dl, (size_t)nb * M * K * sizeof(float),
                          cudaMemcpyDeviceToHost, 0));
      GCU(cudaStreamSynchronize(0));
    }
    return GULON_OK;
  }
  const i64 qb = 4096;
  GCHECK(cb->scratch_q.ensure((size_t)qb * D * sizeof(float) + (size_t)qb * M * K * sizeof(float)));
  GCHECK(cb->scratch_lut.ensure((size_t)(qb / 4) * M * 256 * sizeof(float4)));
  float *dq = cb->scratch_q.as<float>();
  float *dl = dq + (size_t)qb * D;
  for (i64 q0 = 0; q0 < nq; q0 += qb) {
    const i64 nb = std::min<i64>(qb, nq - q0);
    GCU(copy2d(dq, (size_t)D * 4, queries + q0 * ldq, (size_t)ldq * 4, (size_t)D * 4,
                          (size_t)nb, cudaMemcpyHostToDevice, 0, true));
    dim3 lg((unsigned)ceil_div(nb, 4), (unsigned)M);
    GLAUNCH(lut_build_kernel, lg, 256, 0, 0, dq, (i64)D, nb, cb->cb.as<float>(),
            cb->dfrom.as<int32_t>(), cb->ddim.as<int32_t>(), M, K, cb->dmax,
            cb->scratch_lut.as<float4>());
    const i64 total = nb * M * K;
    GLAUNCH(lut_export_kernel, (unsigned)ceil_div(total, 256), 256, 0, 0,
            cb->scratch_lut.as<float4>(), nb, M, K, dl);
    GCU(cudaMemcpyAsync(lut + q0 * M * K, dl, (size_t)total * sizeof(float),
                        cudaMemcpyDeviceToHost, 0));
    GCU(cudaStreamSynchronize(0));
  }
  return GULON_OK;
}

int gulon_debug_tscan(gulon_index_t ix, const float *dqueries, int64_t nq, int64_t ldq, const float *taus,
                      int64_t from, int64_t until, uint16_t *xb, uint16_t *qb, float *acc, int32_t *kp) {
  GCHECK(need_device());
  GREQUIRE(ix && dqueries && taus, "null argument");
  GREQUIRE(nq >= 1 && nq <= tscan::TN, "gulon_debug_tscan takes 1..%d queries", tscan::TN);
  GREQUIRE(from >= 0 && from < until && until <= ix->N && until - from <= (1 << 20), "bad row range");
  GREQUIRE(ldq >= ix->cb->D, "query leading dimension %lld < dimension %d", (long long)ldq, ix->cb->D);
  std::lock_guard<std::mutex> lock(ix->mu);
  cudaStream_t st = nullptr;
  GCHECK(prepare_tensor(ix, st));
  if (ix->tensor_state != 1) return fail(GULON_EUNSUPPORTED, "this index cannot take the tensor scan");
  const int KP = ix->tensor_kp, D = ix->cb->D;
  const i64 rows = until - from;
  if (kp) *kp = KP;
  std::vector<u64> keys((size_t)nq);
  for (i64 q = 0; q < nq; q++) keys[q] = make_key(taus[q], 0);
  DevBuf dk, dacc;
  struct Guard {
    DevBuf &a, &b;
    ~Guard() {
      a.release();
      b.release();
    }
  } guard{dk, dacc};
  GCHECK(upload(dk, keys, st));
  GCHECK(dacc.ensure((size_t)rows * tscan::TN * sizeof(float)));
  GCU(cudaMemsetAsync(dacc.p, 0, (size_t)rows * tscan::TN * sizeof(float), st));
  GCHECK(ix->tqb.ensure((size_t)tscan::TN * KP * 2));
  GCHECK(ix->tsurv.ensure((size_t)TENSOR_CAPB * sizeof(u64)));
  GCHECK(ix->tscount.ensure(sizeof(unsigned)));
  GCHECK(ix->tflag.ensure(4 * sizeof(int)));
  GCHECK(ix->tstats.ensure(8 * sizeof(unsigned long long)));
  GCU(cudaMemsetAsync(ix->tscount.p, 0, sizeof(unsigned), st));
  GCU(cudaMemsetAsync(ix->tflag.p, 0, 4 * sizeof(int), st));
  GLAUNCH(tscan::qprep_kernel, (unsigned)ceil_div(tscan::TN, 8), 256, 0, st, dqueries, (i64)ldq, (i64)nq,
          (i64)tscan::TN, D, KP, dk.as<u64>(), (i64)1, 1, ix->tqb.as<uint16_t>(), ix->tflag.as<int>(), (uint8_t *)nullptr);
  GCHECK(launch_filter(ix, from, until, 1, TENSOR_CAPB, dacc.as<float>(), st));
  if (xb) GCU(cudaMemcpyAsync(xb, ix->xb.as<uint16_t>() + (size_t)from * KP, (size_t)rows * KP * 2, cudaMemcpyDeviceToHost, st));
  if (qb) GCU(cudaMemcpyAsync(qb, ix->tqb.p, (size_t)tscan::TN * KP * 2, cudaMemcpyDeviceToHost, st));
  if (acc) GCU(cudaMemcpyAsync(acc, dacc.p, (size_t)rows * tscan::TN * sizeof(float), cudaMemcpyDeviceToHost, st));
  GCU(cudaStreamSynchronize(st));
  return GULON_OK;
}

int gulon_pq_query_dev(gulon_index_t ix, const float *dqueries, int64_t nq, int64_t ldq, int32_t k,
                       int64_t from, int64_t until, int32_t normalize, int64_t id_offset,
                       int32_t *d_ids, float *d_dists, int32_t *d_sizes, void *stream) {
  GREQUIRE(ix, "null handle");
  GREQUIRE(nq >= 0, "nq must be >= 0");
  GREQUIRE((dqueries && d_ids && d_dists) || nq == 0 || k == 0, "null argument");
  GCHECK(need_device());
  return query_dev(ix, dqueries, nq, ldq, k, from, until, normalize, id_offset, d_ids, d_dists,
                   d_sizes, (cudaStream_t)stream);
}

int gulon_pq_query(gulon_index_t ix, const float *queries, int64_t nq, int64_t ldq, int32_t k,
                   int64_t from, int64_t until, int32_t normalize, int64_t id_offset,
                   int32_t *out_ids, float *out_dists, int32_t *out_sizes) {
  GREQUIRE(ix, "null handle");
  GREQUIRE(nq >= 0 && k >= 0, "nq and k must be >= 0");
  GREQUIRE(ldq >= ix->cb->D, "query leading dimension %lld < dimension %d", (long long)ldq,
           ix->cb->D);
  GREQUIRE((queries && out_ids && out_dists) || nq == 0 || k == 0, "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  const int D = ix->cb->D;
  const size_t nk = (size_t)nq * std::max(k, 1);
  // staging buffers live in the handle (grow-only): no cudaMalloc / cudaFree per call
  std::unique_lock<std::mutex> lock(ix->mu_host);
  GCHECK(ix->h_q.ensure((size_t)nq * D * sizeof(float)));
  GCHECK(ix->h_ids.ensure(nk * sizeof(int32_t)));
  GCHECK(ix->h_dists.ensure(nk * sizeof(float)));
  GCHECK(ix->h_sizes.ensure((size_t)nq * sizeof(int32_t)));
  float *dq = ix->h_q.as<float>();
  int32_t *di = ix->h_ids.as<int32_t>(), *dz = ix->h_sizes.as<int32_t>();
  float *dd = ix->h_dists.as<float>();
  GCU(copy2d(dq, (size_t)D * 4, queries, (size_t)ldq * 4, (size_t)D * 4, (size_t)nq,
                        cudaMemcpyHostToDevice, 0, true));
  GCHECK(query_dev(ix, dq, nq, D, k, from, until, normalize, id_offset, di, dd, dz, 0));
  if (k > 0) {
    GCU(cudaMemcpyAsync(out_ids, di, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
    GCU(cudaMemcpyAsync(out_dists, dd, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, 0));
  }
  if (out_sizes)
    GCU(cudaMemcpyAsync(out_sizes, dz, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
  GCU(cudaStreamSynchronize(0));
  return GULON_OK;
}

int gulon_subtract_rows_dev(const float *dX, int64_t ldx, const int64_t *d_src, const float *dC,
                            int64_t ldc, const int32_t *d_group, int64_t n, int32_t D, float *d_out,
                            int64_t ldo, void *stream) {
  GREQUIRE(n >= 0 && D >= 1 && ldx >= D && ldo >= D, "bad shapes n=%lld D=%d", (long long)n, D);
  GREQUIRE((dX && d_out) || n == 0, "null argument");
  GREQUIRE(!d_group || (dC && ldc >= D), "centroids missing");
  GCHECK(need_device());
  if (n == 0) return GULON_OK;
  GLAUNCH(subtract_rows_kernel, (unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream, dX, (i64)ldx,
          (const i64 *)d_src, dC, (i64)ldc, d_group, (i64)n, D, d_out, (i64)ldo);
  return GULON_OK;
}

int gulon_topk_merge_dev(const int32_t *d_ids, const float *d_dists, int32_t S, int64_t nq,
                         int32_t k, int32_t *d_out_ids, float *d_out_dists, int32_t *d_out_sizes,
                         void *stream) {
  GREQUIRE(S >= 1 && nq >= 0 && k >= 1, "bad merge shape S=%d nq=%lld k=%d", S, (long long)nq, k);
  GREQUIRE((d_ids && d_dists && d_out_ids && d_out_dists) || nq == 0, "null argument");
  GCHECK(need_device());
  return merge_shards(d_ids, d_dists, S, (i64)nq * k, nq, k, d_out_ids, d_out_dists, d_out_sizes,
                      (cudaStream_t)stream);
}

int gulon_pq_query_sharded_dev(gulon_index_t ix, const gulon_comm_t *row_comm,
                               const gulon_comm_t *query_comm, const float *dqueries, int64_t nq,
                               int64_t ldq, int32_t k, int32_t normalize, int64_t row_offset,
                               int32_t *d_ids, float *d_dists, int32_t *d_sizes, void *stream) {
  GCHECK(check_sharded_args(ix, row_comm, query_comm, nq, ldq, k, row_offset));
  GREQUIRE((dqueries && d_ids && d_dists) || nq == 0, "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  const int C = query_comm ? query_comm->world : 1, g = query_comm ? query_comm->rank : 0;
  const i64 per = ceil_div(nq, C), lo = std::min<i64>((i64)g * per, nq), hi = std::min<i64>(lo + per, nq);
  return sharded_query(ix, row_comm, query_comm, dqueries + lo * ldq, hi - lo, per, nq, ldq, k,
                       normalize, row_offset, d_ids, d_dists, d_sizes, (cudaStream_t)stream);
}

int gulon_pq_query_sharded(gulon_index_t ix, const gulon_comm_t *row_comm,
                           const gulon_comm_t *query_comm, const float *queries, int64_t nq,
                           int64_t ldq, int32_t k, int32_t normalize, int64_t row_offset,
                           int32_t *out_ids, float *out_dists, int32_t *out_sizes) {
  GCHECK(check_sharded_args(ix, row_comm, query_comm, nq, ldq, k, row_offset));
  GREQUIRE((queries && out_ids && out_dists) || nq == 0, "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  const int D = ix->cb->D;
  const int C = query_comm ? query_comm->world : 1, g = query_comm ? query_comm->rank : 0;
  const i64 per = ceil_div(nq, C), lo = std::min<i64>((i64)g * per, nq), hi = std::min<i64>(lo + per, nq);
  std::unique_lock<std::mutex> lock(ix->mu_host);
  // only this group's slice of the batch crosses PCIe; the assembled answer comes back whole
  GCHECK(ix->h_q.ensure((size_t)std::max<i64>(hi - lo, 1) * D * sizeof(float)));
  GCHECK(ix->h_ids.ensure((size_t)nq * k * sizeof(int32_t)));
  GCHECK(ix->h_dists.ensure((size_t)nq * k * sizeof(float)));
  GCHECK(ix->h_sizes.ensure((size_t)nq * sizeof(int32_t)));
  if (hi > lo)
    GCU(copy2d(ix->h_q.p, (size_t)D * 4, queries + lo * ldq, (size_t)ldq * 4, (size_t)D * 4,
                          (size_t)(hi - lo), cudaMemcpyHostToDevice, 0, true));
  GCHECK(sharded_query(ix, row_comm, query_comm, ix->h_q.as<float>(), hi - lo, per, nq, D, k, normalize,
                       row_offset, ix->h_ids.as<int32_t>(), ix->h_dists.as<float>(),
                       ix->h_sizes.as<int32_t>(), 0));
  GCU(cudaMemcpyAsync(out_ids, ix->h_ids.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
  GCU(cudaMemcpyAsync(out_dists, ix->h_dists.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, 0));
  if (out_sizes)
    GCU(cudaMemcpyAsync(out_sizes, ix->h_sizes.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
  GCU(cudaStreamSynchronize(0));
  return GULON_OK;
}

/* Synthetic data of SURVEY.md 8d (benchmark / test tooling; CPU twin: oracle/synth.c). */
int gulon_synth_tables_dev(const gulon_synth_params_t *prm, float *d_centres, float *d_map, void *stream) {
  GREQUIRE(prm && prm->D >= 1 && prm->centres >= 0 && prm->latent >= 0 && prm->latent <= synth::LMAX,
           "bad synth parameters");
  GREQUIRE(prm->centres == 0 || d_centres, "null centre table");
  GREQUIRE(prm->latent == 0 || d_map, "null map table");
  GCHECK(need_device());
  if (prm->centres == 0) return GULON_OK;
  gs_params p;
  static_assert(sizeof(gs_params) == sizeof(gulon_synth_params_t), "synth parameter layouts differ");
  memcpy(&p, prm, sizeof(p));
  const i64 W = p.latent > 0 ? p.latent : p.D;
  const i64 total = (i64)p.centres * W + (i64)p.latent * p.D;
  GLAUNCH(synth::tables_kernel, (unsigned)std::min<i64>(ceil_div(total, 256), 4096), 256, 0,
          (cudaStream_t)stream, p, d_centres, d_map);
  return GULON_OK;
}

int gulon_synth_rows_dev(const gulon_synth_params_t *prm, int64_t stream_id, int64_t lo, int64_t n,
                         const float *d_centres, const float *d_map, float *d_out, int64_t ld,
                         void *stream) {
  GREQUIRE(prm && prm->D >= 1 && prm->centres >= 0 && prm->latent >= 0 && prm->latent <= synth::LMAX,
           "bad synth parameters");
  GREQUIRE(n >= 0 && lo >= 0 && ld >= prm->D && stream_id >= 0, "bad synth row range");
  GREQUIRE(n == 0 || d_out, "null output");
  GREQUIRE(prm->centres == 0 || d_centres, "null centre table");
  GREQUIRE(prm->latent == 0 || d_map, "null map table");
  GCHECK(need_device());
  if (n == 0) return GULON_OK;
  gs_params p;
  memcpy(&p, prm, sizeof(p));
  const unsigned grid = (unsigned)std::min<i64>(ceil_div(n, synth::ROWS), 16LL * sm_count());
  GLAUNCH(synth::rows_kernel, grid, 256, 0, (cudaStream_t)stream, p, (i64)stream_id, (i64)lo, (i64)n,
          d_centres, d_map, d_out, (i64)ld);
  return GULON_OK;
}

int gulon_grouped_query_dev(gulon_index_t ix, const float *dqueries, int64_t nq, int64_t ldq,
                            const float *d_centroids, int32_t n_partitions, const int32_t *d_bounds,
                            const int32_t *d_pair_query, const int32_t *d_pair_partition,
                            const int32_t *d_pair_slot, int64_t n_pairs, int32_t slots, int32_t k,
                            int32_t *d_ids, float *d_dists, int32_t *d_sizes, void *stream) {
  GREQUIRE(ix, "null handle");
  GREQUIRE(!ix->codes16, "the grouped scan serves 8-bit indexes");
  GREQUIRE(nq >= 0 && n_pairs >= 0 && slots >= 0 && n_partitions >= 0, "bad grouped-query shape");
  GREQUIRE(k >= 1 && k <= gscan::GS_CHUNK / 2, "the grouped scan takes 1 <= k <= %d (k=%d)", gscan::GS_CHUNK / 2, k);
  GREQUIRE(ldq >= ix->cb->D, "query leading dimension %lld < dimension %d", (long long)ldq, ix->cb->D);
  GREQUIRE((dqueries && d_ids && d_dists) || nq == 0, "null argument");
  GREQUIRE((d_centroids && d_bounds && d_pair_query && d_pair_partition && d_pair_slot) || n_pairs == 0,
           "null work list");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  cudaStream_t st = (cudaStream_t)stream;
  gulon_codebook_t cb = ix->cb;
  const int D = cb->D, M = cb->M;
  if (n_pairs == 0 || slots == 0) return fill_empty(nq, k, d_ids, d_dists, d_sizes, st);
  const size_t smem = (size_t)gscan::GS_CHUNK * 8 + ((size_t)((D + 3) & ~3) + (size_t)M * 256) * sizeof(float);
  GREQUIRE(smem <= 227 * 1024, "D=%d, M=%d do not fit the shared memory of the grouped scan", D, M);
  GOPTIN(gscan::grouped_pairs_kernel, 227 * 1024);
  std::lock_guard<std::mutex> lock(ix->mu);
  GCHECK(ix->chain.enter(st));
  const int rc = [&]() -> int {
    // keys [nq][stride]: slot s of query q at [q][s * k ..); unused slots stay empty
    const i64 want = (i64)slots * k;
    const bool small = want <= MERGE_WARP_MAX;
    i64 stride = SEL_CHUNK;
    if (small) {
      stride = 2;
      while (stride < want) stride <<= 1;
    } else {
      stride = round_up(want, SEL_CHUNK);
    }
    GCHECK(ix->keys.ensure((size_t)nq * stride * sizeof(u64)));
    GCU(cudaMemsetAsync(ix->keys.p, 0xFF, (size_t)nq * stride * sizeof(u64), st));
    for (i64 p0 = 0; p0 < n_pairs; p0 += 1 << 30) {
      const i64 np = std::min<i64>(1 << 30, n_pairs - p0);
      GLAUNCH(gscan::grouped_pairs_kernel, (unsigned)np, gscan::NT, smem, st, ix->codes, ix->ps, dqueries,
              (i64)ldq, d_centroids, D, d_bounds, d_pair_query + p0, d_pair_partition + p0, d_pair_slot + p0,
              cb->cb.as<float>(), cb->dfrom.as<int32_t>(), cb->ddim.as<int32_t>(), M, cb->K, cb->dmax, k,
              ix->keys.as<u64>(), stride);
    }
    // TopKHeap#merge of a query's partition heaps
    if (small) {
      GLAUNCH(sort_rows_small_kernel, (unsigned)ceil_div(nq, 4), 128, (size_t)4 * stride * sizeof(u64), st,
              ix->keys.as<u64>(), (int)stride, (i64)nq);
      return unpack(ix->keys.as<u64>(), stride, nq, k, 0, d_ids, d_dists, d_sizes, st);
    }
    u64 *res;
    i64 rs;
    GCHECK(ix->sel.run(ix->keys.as<u64>(), stride, nq, k, st, &res, &rs));
    return unpack(res, rs, nq, k, 0, d_ids, d_dists, d_sizes, st);
  }();
  GCHECK(ix->chain.leave(st));
  return rc;
}

int gulon_exact_topk(gulon_points_t p, const float *queries, int64_t nq, int64_t ldq, int32_t k,
                     int64_t from, int64_t until, int32_t *out_ids, float *out_dists,
                     int32_t *out_sizes) {
  GREQUIRE(p, "null handle");
  GREQUIRE(from <= until, "invalid range: expected from=%lld <= until=%lld", (long long)from,
           (long long)until);
  GREQUIRE(from >= 0 && until <= p->N, "invalid range: expected until=%lld <= vectors.length=%lld",
           (long long)until, (long long)p->N);
  GREQUIRE(nq >= 0 && k >= 0 && ldq >= p->D, "bad query shape");
  GREQUIRE((queries && out_ids && out_dists) || nq == 0 || k == 0, "null argument");
  GREQUIRE(p->N < (1LL << 31), "row ids are Int in the reference: N must be < 2^31");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  const int D = p->D;
  const i64 range = until - from;
  DevBuf dq, keys, di, dd, dz;
  Selector sel;
  const size_t nk = (size_t)nq * std::max(k, 1);
  int rc = [&]() -> int {
    GCHECK(di.ensure(nk * sizeof(int32_t)));
    GCHECK(dd.ensure(nk * sizeof(float)));
    GCHECK(dz.ensure((size_t)nq * sizeof(int32_t)));
    if (k == 0 || range == 0) {
      GCHECK(fill_empty(nq, k, di.as<int32_t>(), dd.as<float>(), dz.as<int32_t>(), 0));
    } else {
      GREQUIRE((size_t)D * sizeof(float) <= 48 * 1024, "dimension %d too large for exact_topk", D);
      GCHECK(dq.ensure((size_t)nq * D * sizeof(float)));
      GCU(copy2d(dq.p, (size_t)D * 4, queries, (size_t)ldq * 4, (size_t)D * 4,
                            (size_t)nq, cudaMemcpyHostToDevice, 0, true));
      const i64 n_pad = round_up(range, SEL_CHUNK);
      i64 qb = std::max<i64>(1, g_simple_scratch.load() / (8 * n_pad));
      qb = std::min<i64>(std::min<i64>(qb, nq), 32768);
      GCHECK(keys.ensure((size_t)qb * n_pad * sizeof(u64)));
      for (i64 q0 = 0; q0 < nq; q0 += qb) {
        const i64 nb = std::min<i64>(qb, nq - q0);
        dim3 grid((unsigned)(n_pad / 128), (unsigned)nb);
        GLAUNCH(exact_keys_kernel, grid, 128, (size_t)D * sizeof(float), 0, p->d, p->ld, D, from,
                until, dq.as<float>() + q0 * D, (i64)D, keys.as<u64>(), n_pad);
        u64 *res;
        i64 rs;
        GCHECK(sel.run(keys.as<u64>(), n_pad, nb, k, 0, &res, &rs));
        GCHECK(unpack(res, rs, nb, k, 0, di.as<int32_t>() + q0 * k, dd.as<float>() + q0 * k,
                      dz.as<int32_t>() + q0, 0));
      }
    }
    if (k > 0) {
      GCU(cudaMemcpyAsync(out_ids, di.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
      GCU(cudaMemcpyAsync(out_dists, dd.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, 0));
    }
    if (out_sizes)
      GCU(cudaMemcpyAsync(out_sizes, dz.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
    GCU(cudaStreamSynchronize(0));
    return GULON_OK;
  }();
  dq.release(); keys.release(); di.release(); dd.release(); dz.release();
  sel.release();
  return rc;
}

int gulon_rerank(gulon_points_t p, const float *queries, int64_t nq, int64_t ldq,
                 const int32_t *cand_ids, int32_t R, int32_t k, int32_t *out_ids,
                 float *out_dists, int32_t *out_sizes) {
  GREQUIRE(p, "null handle");
  GREQUIRE(nq >= 0 && R >= 0 && k >= 0 && ldq >= p->D, "bad re-rank shape");
  GREQUIRE((queries && cand_ids && out_ids && out_dists) || nq == 0 || k == 0 || R == 0,
           "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  const int D = p->D;
  DevBuf dq, dc, keys, di, dd, dz;
  Selector sel;
  const size_t nk = (size_t)nq * std::max(k, 1);
  int rc = [&]() -> int {
    GCHECK(di.ensure(nk * sizeof(int32_t)));
    GCHECK(dd.ensure(nk * sizeof(float)));
    GCHECK(dz.ensure((size_t)nq * sizeof(int32_t)));
    GCHECK(dq.ensure((size_t)nq * D * sizeof(float)));
    GCHECK(dc.ensure((size_t)std::max<i64>(1, nq * R) * sizeof(int32_t)));
    if (k > 0 && R > 0) {
      GCU(copy2d(dq.p, (size_t)D * 4, queries, (size_t)ldq * 4, (size_t)D * 4,
                            (size_t)nq, cudaMemcpyHostToDevice, 0, true));
      GCU(cudaMemcpyAsync(dc.p, cand_ids, (size_t)nq * R * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
    }
    GCHECK(rerank_dev(p, dq.as<float>(), nq, D, dc.as<int32_t>(), R, k, 0, di.as<int32_t>(), dd.as<float>(),
                      dz.as<int32_t>(), keys, sel, 0));
    if (k > 0) {
      GCU(cudaMemcpyAsync(out_ids, di.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
      GCU(cudaMemcpyAsync(out_dists, dd.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, 0));
    }
    if (out_sizes)
      GCU(cudaMemcpyAsync(out_sizes, dz.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
    GCU(cudaStreamSynchronize(0));
    return GULON_OK;
  }();
  dq.release(); dc.release(); keys.release(); di.release(); dd.release(); dz.release();
  sel.release();
  return rc;
}

int gulon_rerank_dev(gulon_points_t p, const float *dqueries, int64_t nq, int64_t ldq,
                     const int32_t *d_cand_ids, int32_t R, int32_t k, int64_t id_lo, int32_t *d_ids,
                     float *d_dists, int32_t *d_sizes, void *stream) {
  GREQUIRE(p, "null handle");
  GREQUIRE(nq >= 0 && R >= 0 && R <= RERANK_RMAX && k >= 0 && ldq >= p->D && id_lo >= 0,
           "bad re-rank shape (R <= %d on the device path)", RERANK_RMAX);
  GREQUIRE((dqueries && d_cand_ids && d_ids && d_dists) || nq == 0 || k == 0 || R == 0, "null argument");
  GCHECK(need_device());
  DevBuf none;
  Selector nosel;
  return rerank_dev(p, dqueries, nq, ldq, d_cand_ids, R, k, id_lo, d_ids, d_dists, d_sizes, none, nosel,
                    (cudaStream_t)stream);
}

static int check_rerank_query(gulon_index_t ix, gulon_points_t pts, const gulon_comm_t *row_comm, i64 nq,
                              i64 ldq, int k, int R, i64 row_offset) {
  GCHECK(check_sharded_args(ix, row_comm, nullptr, nq, ldq, std::max(k, 1), row_offset));
  GREQUIRE(pts, "null points handle");
  GREQUIRE(pts->D == ix->cb->D && pts->N == ix->N, "the raw vectors must be the rows the index encodes "
           "(points %lld x %d, index %lld x %d)", (long long)pts->N, pts->D, (long long)ix->N, ix->cb->D);
  GREQUIRE(k >= 1 && R >= k && R <= RERANK_RMAX, "need 1 <= k <= R <= %d (k=%d R=%d)", RERANK_RMAX, k, R);
  return GULON_OK;
}

int gulon_pq_rerank_query_dev(gulon_index_t ix, gulon_points_t points, const gulon_comm_t *row_comm,
                              const float *dqueries, int64_t nq, int64_t ldq, int32_t k, int32_t R,
                              int32_t normalize, int64_t row_offset, int32_t *d_ids, float *d_dists,
                              int32_t *d_sizes, void *stream) {
  GCHECK(check_rerank_query(ix, points, row_comm, nq, ldq, k, R, row_offset));
  GREQUIRE((dqueries && d_ids && d_dists) || nq == 0, "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  return rerank_query(ix, points, row_comm, dqueries, nq, ldq, k, R, normalize, row_offset, d_ids, d_dists,
                      d_sizes, (cudaStream_t)stream);
}

int gulon_pq_rerank_query(gulon_index_t ix, gulon_points_t points, const gulon_comm_t *row_comm,
                          const float *queries, int64_t nq, int64_t ldq, int32_t k, int32_t R,
                          int32_t normalize, int64_t row_offset, int32_t *out_ids, float *out_dists,
                          int32_t *out_sizes) {
  GCHECK(check_rerank_query(ix, points, row_comm, nq, ldq, k, R, row_offset));
  GREQUIRE((queries && out_ids && out_dists) || nq == 0, "null argument");
  GCHECK(need_device());
  if (nq == 0) return GULON_OK;
  const int D = ix->cb->D;
  std::unique_lock<std::mutex> lock(ix->mu_host);
  GCHECK(ix->h_q.ensure((size_t)nq * D * sizeof(float)));
  GCHECK(ix->h_ids.ensure((size_t)nq * k * sizeof(int32_t)));
  GCHECK(ix->h_dists.ensure((size_t)nq * k * sizeof(float)));
  GCHECK(ix->h_sizes.ensure((size_t)nq * sizeof(int32_t)));
  GCU(copy2d(ix->h_q.p, (size_t)D * 4, queries, (size_t)ldq * 4, (size_t)D * 4, (size_t)nq,
                        cudaMemcpyHostToDevice, 0, true));
  GCHECK(rerank_query(ix, points, row_comm, ix->h_q.as<float>(), nq, D, k, R, normalize, row_offset,
                      ix->h_ids.as<int32_t>(), ix->h_dists.as<float>(), ix->h_sizes.as<int32_t>(), 0));
  GCU(cudaMemcpyAsync(out_ids, ix->h_ids.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
  GCU(cudaMemcpyAsync(out_dists, ix->h_dists.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, 0));
  if (out_sizes)
    GCU(cudaMemcpyAsync(out_sizes, ix->h_sizes.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
  GCU(cudaStreamSynchronize(0));
  return GULON_OK;
}

}  // extern "C"
