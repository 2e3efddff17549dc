/*
 * gulon_b200.h -- C ABI of libgulon_b200.so: the B200 (sm_100a) implementation of tixxit/gulon's
 * product-quantization hot path (k-means codebook training, PQ encoding, ADC scan + top-k).
 *
 * The reference has no FFI seam: the hot path is plain Scala method calls inside one JVM.  Each
 * entry point below names the Scala method it replaces (G/ = core/src/main/scala/net/tixxit/gulon/).
 * INTEGRATION.md shows the JNI / Panama binding a gulon maintainer would add.
 *
 * Conventions
 *   - every function returns a status (GULON_OK == 0, negative on error); gulon_last_error()
 *     returns the calling thread's message for the last failure;
 *   - matrices are flat row-major float32 with an explicit leading dimension `ld` (in floats);
 *     the Scala facade flattens Matrix.data (jagged Array[Array[Float]], G/Matrix.scala:3) once;
 *   - PQ codes are plane-major uint8 [M][plane_stride] (one Coder8 byte array per quantizer,
 *     G/EncodedMatrix.scala:11-23, G/Coder.scala:129-140);
 *   - codebooks are float32 [M][K][dmax], dmax = ceil(D/M); quantizer m uses the first dim[m]
 *     floats of each centroid, with (from[m], dim[m]) given by the Vectors.subvectors split rule
 *     (G/Vectors.scala:84-104) -- see gulon_subvectors();
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     GULON_ENODEVICE;
 *   - functions suffixed _dev take DEVICE pointers (and a cudaStream_t passed as void*), are
 *     asynchronous on that stream, and exist so that a host can keep data resident in HBM and
 *     compose the path with its own collectives (torch.distributed / NCCL).
 *
 * Semantics fixed by this library where the reference is order- or RNG-dependent
 *   - argmin ties: lowest centroid index (GULON_TIE_LOWEST).  The reference breaks exact ties with
 *     a stateful Random(0).nextBoolean() stream (G/KMeans.scala:47,90);
 *   - top-k ties: (distance ascending, id ascending) -- the stable order T/TopKHeapSpec.scala:16-31
 *     asserts; the reference heap's order inside an equal-distance group is structure-dependent.
 */
#ifndef GULON_B200_H
#define GULON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GULON_VERSION 100 /* 0.1.0 */

/* status codes */
#define GULON_OK            0
#define GULON_EINVAL       -1  /* IllegalArgumentException on the Scala side (require(...)) */
#define GULON_ECUDA        -2
#define GULON_ENOMEM       -3
#define GULON_ENODEVICE    -4
#define GULON_ECOMM        -5
#define GULON_EUNSUPPORTED -6
#define GULON_ESTATE       -7  /* IllegalStateException */

/* argmin tie rule (G/KMeans.scala:47,90) */
#define GULON_TIE_LOWEST  1

/* centroid update rule (G/KMeans.scala:198-226) */
#define GULON_UPDATE_RUNNING_MEAN 0 /* literal sequential running mean: bit-exact with the reference */
#define GULON_UPDATE_SUM          1 /* per-cluster sum / count; shardable (all-reduce of sums+counts).
                                      * On K <= 256, widths <= 16 the sums are exact FIXED-POINT integers:
                                      * every coordinate is first rounded to 2^-28 of the window's largest
                                      * |x| (so the centroid error is <= 2^-28 max|x| per coordinate, i.e.
                                      * relative to the WINDOW maximum, not to the coordinate: one outlier of
                                      * 1e3 among values of 1e-3 leaves ~2e-6 absolute = 2e-3 relative on the
                                      * small ones); the result is bit-identical for any number of ranks. */

/* scan implementation selector for gulon_set_option("scan_impl", ...) */
#define GULON_SCAN_AUTO   0
#define GULON_SCAN_SIMPLE 1 /* distance materialisation + selection (any k, slow, cross-check path) */
#define GULON_SCAN_FUSED  2 /* replicated-LUT gather kernel with in-kernel top-k (k <= 128)        */
#define GULON_SCAN_PRUNED 3 /* 16-bit lower-bound pass + exact fp32 re-evaluation of survivors      */
#define GULON_SCAN_TENSOR 4 /* tcgen05 lower bound over the decoded rows (bf16 contraction with a
                             * rigorous error term) + exact fp32 re-evaluation of survivors; large
                             * query batches on long ranges, D <= 1020, K <= 256, k <= 128.  Forced on an
                             * index / batch it cannot serve it fails with GULON_EUNSUPPORTED.          */

/* assignment / encode implementation selector for gulon_set_option("assign_impl", ...).
 * Both produce the same bits: the tensor path only discards centroids that provably cannot win. */
#define GULON_ASSIGN_AUTO   0
#define GULON_ASSIGN_EXACT  1 /* CUDA-core kernel: all K scores in the reference's arithmetic          */
#define GULON_ASSIGN_TENSOR 2 /* tcgen05 approximate scores + exact recheck of near-minimum centroids  */

typedef struct gulon_points_s   *gulon_points_t;   /* device-resident float32 matrix (Matrix)      */
typedef struct gulon_codebook_s *gulon_codebook_t; /* device-resident ProductQuantizer codebooks   */
typedef struct gulon_index_s    *gulon_index_t;    /* device-resident Index.PQIndex                */

/*
 * Exchange hooks for the sharded paths (one process per GPU).  Buffers are DEVICE pointers;
 * each hook must enqueue on `stream` (a cudaStream_t) or synchronise before returning.
 * The Python host wires these to torch.distributed (NCCL); a JVM host would wire them to NCCL.
 * The reference has no counterpart (single JVM).
 */
typedef struct gulon_comm {
  int32_t rank, world;
  int (*allreduce_sum_f32)(void *user, float *dev_buf, int64_t n, void *stream);
  int (*allreduce_sum_i32)(void *user, int32_t *dev_buf, int64_t n, void *stream);
  int (*allgather)(void *user, const void *dev_send, void *dev_recv, int64_t bytes_per_rank,
                   void *stream);
  void *user;
  /* Optional (may be NULL).  With both present, GULON_UPDATE_SUM accumulates in fixed point and
   * all-reduces the sums as int64: the trained centroids are then bit-identical for ANY number of
   * ranks.  Without them the sums are all-reduced in fp32 (rank-count dependent in the last bits). */
  int (*allreduce_sum_i64)(void *user, int64_t *dev_buf, int64_t n, void *stream);
  int (*allreduce_max_f32)(void *user, float *dev_buf, int64_t n, void *stream);
} gulon_comm_t;

/* KMeans.ProgressReport (G/KMeans.scala:119-127), delivered synchronously on the calling thread.
 * quantizer = index of the sub-quantizer (0 for a plain k-means). */
typedef struct gulon_progress {
  int32_t quantizer;
  int32_t num_iterations;
  int32_t max_iterations;
  float step_mean, step_stddev; /* SummaryStats of centroid displacement, G/KMeans.scala:160-168 */
  int32_t converged;
  /* the SummaryStats(count, mean, s) triple itself (G/MathUtils.scala:5), so that a host can rebuild
   * the reference's value exactly: count = number of centroids, s = sum of squared deviations */
  int32_t step_count;
  float step_s;
} gulon_progress_t;
typedef void (*gulon_progress_fn)(void *user, const gulon_progress_t *report);

/* ---- library / device ------------------------------------------------------------------- */
int gulon_version(void);
const char *gulon_last_error(void);
int gulon_device_count(int32_t *n);
int gulon_set_device(int32_t device); /* device used by the calling thread's later calls */
int gulon_get_device(int32_t *device);
int gulon_device_sync(void);
/* Optional: validates the devices a host will use (they must exist and be sm_100) and creates their
 * contexts up front; the calling thread's current device is unchanged.  gulon_shutdown waits for the
 * current device's outstanding work.  Handles are destroyed explicitly, not by gulon_shutdown. */
int gulon_init(const int32_t *devices, int32_t n);
int gulon_shutdown(void);
/* Tuning options (results never depend on them).  Scan: "scan_impl" (GULON_SCAN_*), "query_batch",
 * "pruned_min_rows" (ranges shorter than this use the exact kernel; default 262144), "boot_rows"
 * (0 = range / 64 in whole 8192-row chunks, clamped to [8192, 32768]), "pruned_bits" (0 auto | 8 | 16), "pruned_words"
 * (0 auto | 1 | 2 | 4), "pruned_lb_quantizers" (0 = measured-cost feedback per index, else the
 * number of quantizers the lower bound sums), "pruned_stage_div" (first stage = range / div rows,
 * 0 = one stage), "pruned_rowcodes" (row-major copy of the codes for the survivor evaluation).
 * Tensor scan: "tensor_min_rows" (default 2^16), "tensor_min_queries" (default 256) and "tensor_min_pairs"
 * (rows x queries, default 2^27): smaller ranges / batches keep the pruned or exact scan under
 * GULON_SCAN_AUTO; "tensor_query_batch" (queries per pass, 0 = auto), "tensor_stage_ratio" (a stage scans
 * ratio x the rows seen so far, 0 = auto from k), "tensor_boot_rows" (rows scanned exactly first by the
 * exact scan kernel; 0 = auto: the first 256 rows through a dedicated kernel when the stages grow 4x, 8192
 * through the exact kernel when k is large and they grow 2x), "tensor_max_bytes" (the decoded bf16 copy of
 * the index is built only below this size; default 64 GiB), "tensor_chunk_bytes" (operand rows per row
 * split = the L2 working set shared by the CTAs; default 16 MiB), "tensor_pair" (1: the filter runs on CTA
 * pairs with tcgen05 cta_group::2, the default; 0: on single CTAs), "tensor_eval_blocks" (CTAs per query
 * block of the survivor evaluation; default 16), "tensor_epi_wait" (how the filter's warps wait: bits 0-1
 * the epilogue warps -- 0 spin, 1 try_wait, 2 try_wait with a suspend hint (default), 3 a named barrier
 * released by a sentinel warp; bit 2: the TMA thread suspends instead of spinning).
 * Assignment: "assign_impl", "assign_tc_min_rows", "update_fixed".  "profile" = 1 turns the kernel
 * timers and counters of gulon_get_counter on. */
int gulon_set_option(const char *name, int64_t value);
int gulon_get_counter(const char *name, int64_t *value); /* e.g. "kernel_launches" */
/* Diagnostics of the tensor scan (tests): runs the operand builders and the filter kernel of
 * GULON_SCAN_TENSOR for nq <= 256 device queries over rows [from, until) with the given per-query
 * thresholds (host array), and returns the operands and the raw accumulators: xb [until - from][KP]
 * and qb [256][KP] bf16 bit patterns, acc [until - from][256] fp32 (host buffers, any may be NULL);
 * *kp receives KP.  A survivor is an accumulator <= 0. */
int gulon_debug_tscan(gulon_index_t ix, const float *dqueries, int64_t nq, int64_t ldq, const float *taus,
                      int64_t from, int64_t until, uint16_t *xb, uint16_t *qb, float *acc, int32_t *kp);

/* Vectors.subvectors split rule, G/Vectors.scala:84-104.  Returns dmax (>0) or a status (<0). */
int gulon_subvectors(int32_t D, int32_t M, int32_t *from, int32_t *dim);

/* ---- Matrix ----------------------------------------------------------------------------- */
/* Copies a host matrix into HBM (G/Matrix.scala:3). */
int gulon_points_create(const float *X, int64_t N, int32_t D, int64_t ld, gulon_points_t *out);
/* Wraps an existing device matrix without copying (borrowed; caller keeps it alive). */
int gulon_points_wrap_dev(const float *dX, int64_t N, int32_t D, int64_t ld, gulon_points_t *out);
int gulon_points_info(gulon_points_t p, int64_t *N, int32_t *D, int64_t *ld, const float **dptr);
int gulon_points_destroy(gulon_points_t p);

/* MathUtils.normalize per row, G/MathUtils.scala:100-120 (cosine metric prep). In place on device. */
int gulon_points_normalize(gulon_points_t p);
/* Host convenience: out[i] = normalize(X[i]). */
int gulon_normalize(const float *X, int64_t N, int32_t D, int64_t ld, float *out, int64_t ldo);

/*
 * MathUtils.subtract per row on device buffers: out[i] = X[src ? src[i] : i] - C[group[i]] (fp32).
 * This is WordVectors.Grouped#residuals (G/WordVectors.scala:118-138: rows minus the centroid of
 * their group) and the residual query of GroupedIndex#query (G/Index.scala:277).  d_src (int64 row
 * ids) may be NULL (identity); d_group NULL gathers rows without subtracting (the grouped matrix of
 * WordVectors#grouped, G/WordVectors.scala:24-58).
 */
int gulon_subtract_rows_dev(const float *dX, int64_t ldx, const int64_t *d_src, const float *dC,
                            int64_t ldc, const int32_t *d_group, int64_t n, int32_t D, float *d_out,
                            int64_t ldo, void *stream);

/* ---- KMeans (G/KMeans.scala) -------------------------------------------------------------- */
/*
 * KMeans#assign (batch = 0, G/KMeans.scala:18-22,70-98) and #parAssign (batch = 25000, :57-68)
 * over the column window [from, from+dim).  C is host float32 [K][dim].  With GULON_TIE_LOWEST
 * the batch size does not change the result (it only scopes the reference's RNG).
 * out: host int32 [N].
 */
int gulon_kmeans_assign(gulon_points_t p, int32_t from, int32_t dim, const float *C, int32_t K,
                        int64_t batch, int32_t tie_mode, int32_t *out);
/* KMeans.fromAssignment, G/KMeans.scala:198-226.  assign: host int32 [N]; out_C host [K][dim];
 * out_counts (optional) host int32 [K]. Empty clusters stay all-zero. */
int gulon_kmeans_update(gulon_points_t p, int32_t from, int32_t dim, const int32_t *assign,
                        int32_t K, int32_t update_mode, float *out_C, int32_t *out_counts);
/* KMeans.init, G/KMeans.scala:188-196: K rows sampled with replacement by java.util.Random(seed). */
int gulon_kmeans_init(gulon_points_t p, int32_t from, int32_t dim, int32_t K, int32_t seed,
                      float *out_C, int32_t *out_rows);
/*
 * KMeans.computeClusters, G/KMeans.scala:134-157: init -> parAssign -> loop i = 0..max_iter
 * {fromAssignment -> parAssign -> converged = (assignments unchanged)}.  comm may be NULL (one GPU);
 * with comm the points are this rank's row shard, update_mode must be GULON_UPDATE_SUM, and
 * `n_total`/`row_offset` give the global row count and this shard's first global row (init samples
 * global rows; the owner broadcasts them through allreduce).  out_C host [K][dim].
 */
int gulon_kmeans_train(gulon_points_t p, int32_t from, int32_t dim, int32_t K, int32_t max_iter,
                       int32_t seed, int32_t tie_mode, int32_t update_mode,
                       const gulon_comm_t *comm, int64_t n_total, int64_t row_offset,
                       gulon_progress_fn report, void *user, float *out_C, int32_t *n_updates,
                       int32_t *converged);

/* ---- ProductQuantizer (G/ProductQuantizer.scala) ------------------------------------------ */
/* ProductQuantizer.apply, :121-153: M independent k-means, seed = quantizer index (:139). */
int gulon_pq_train(gulon_points_t p, int32_t M, int32_t K, int32_t max_iter, int32_t tie_mode,
                   int32_t update_mode, const gulon_comm_t *comm, int64_t n_total,
                   int64_t row_offset, gulon_progress_fn report, void *user,
                   gulon_codebook_t *out);
/* ProductQuantizer(numClusters, quantizers) from explicit centroids: host float32 [M][K][dmax]. */
int gulon_codebook_create(int32_t D, int32_t M, int32_t K, const float *centroids,
                          gulon_codebook_t *out);
int gulon_codebook_info(gulon_codebook_t cb, int32_t *D, int32_t *M, int32_t *K, int32_t *dmax);
int gulon_codebook_export(gulon_codebook_t cb, float *centroids); /* host [M][K][dmax] */
int gulon_codebook_destroy(gulon_codebook_t cb);

/* ProductQuantizer#encode, :25-35 (+ Coder8, G/Coder.scala:129-140). Requires K <= 256.
 * X host [N][ld]; codes host uint8 [M][N].  Streams X through HBM in row chunks. */
int gulon_pq_encode(gulon_codebook_t cb, const float *X, int64_t N, int64_t ld, int32_t tie_mode,
                    uint8_t *codes);
/* Device-resident form: dX device [N][ld], dcodes device uint8 [M][plane_stride]. */
int gulon_pq_encode_dev(gulon_codebook_t cb, const float *dX, int64_t N, int64_t ld,
                        int32_t tie_mode, uint8_t *dcodes, int64_t plane_stride, void *stream);
/* The same with 16-bit centroid ids (K <= 65536). */
int gulon_pq_encode16(gulon_codebook_t cb, const float *X, int64_t N, int64_t ld, int32_t tie_mode,
                      uint16_t *codes);
int gulon_pq_encode16_dev(gulon_codebook_t cb, const float *dX, int64_t N, int64_t ld,
                          int32_t tie_mode, uint16_t *dcodes, int64_t plane_stride, void *stream);
int gulon_pq_decode16(gulon_codebook_t cb, const uint16_t *codes, int64_t N, int64_t plane_stride,
                      float *out, int64_t ldo);
/* ProductQuantizer#decode(EncodedMatrix), :58-78. codes host [M][N] -> out host [N][ldo]. */
int gulon_pq_decode(gulon_codebook_t cb, const uint8_t *codes, int64_t N, int64_t plane_stride,
                    float *out, int64_t ldo);

/* ---- Index.PQIndex (G/Index.scala:385-441) ------------------------------------------------ */
/* PQIndex(productQuantizer, data): copies host codes [M][plane_stride] into HBM. */
int gulon_index_create(gulon_codebook_t cb, const uint8_t *codes, int64_t N, int64_t plane_stride,
                       gulon_index_t *out);
/* Adopts device codes (borrowed). plane_stride must be a multiple of 16 and dcodes 16-byte aligned. */
int gulon_index_create_dev(gulon_codebook_t cb, const uint8_t *dcodes, int64_t N,
                           int64_t plane_stride, gulon_index_t *out);
/* Wide indexes: 256 < K <= 65536, the reference's BytePlus coders (G/Coder.scala:142-168).  The ids
 * cross the boundary unpacked, one uint16 per (quantizer, row); plane_stride counts elements.  Queries
 * go through gulon_pq_query / gulon_pq_query_dev like any index.  Ranges of >= pruned_min_rows rows with
 * k <= 128 take the lower-bound scan: on first use the index clusters every quantizer's K centroids into
 * 256 groups and keeps 8-bit group planes (N x M bytes) plus a row-major copy of the ids (N x 2M bytes);
 * the bound pass runs over group ids with per-group minimum tables, survivors are evaluated with the
 * real ids -- same bits as the plain-table scan + selection that serves every other case. */
int gulon_index_create16(gulon_codebook_t cb, const uint16_t *codes, int64_t N, int64_t plane_stride,
                         gulon_index_t *out);
int gulon_index_create16_dev(gulon_codebook_t cb, const uint16_t *dcodes, int64_t N,
                             int64_t plane_stride, gulon_index_t *out);
int gulon_index_info(gulon_index_t ix, int64_t *N, int32_t *M, int32_t *K, int32_t *D);
int gulon_index_destroy(gulon_index_t ix);

/* Index.prepareQuery, G/Index.scala:352-383: lut host float32 [nq][M][K]. */
int gulon_prepare_query(gulon_codebook_t cb, const float *queries, int64_t nq, int64_t ldq,
                        float *lut);

/*
 * PQIndex#batchQuery(k, vectors, from, until), G/Index.scala:414-440 (+ SortedIndex.prepare,
 * :324-331, when normalize != 0; + Result.fromHeap ordering, :83-94).
 * queries host [nq][ldq]; out_ids host int32 [nq][k] (row ids + id_offset), out_dists host
 * float32 [nq][k] ascending, out_sizes host int32 [nq] = min(k, until-from).  Unused slots hold
 * id -1 / +inf.
 */
int gulon_pq_query(gulon_index_t ix, const float *queries, int64_t nq, int64_t ldq, int32_t k,
                   int64_t from, int64_t until, int32_t normalize, int64_t id_offset,
                   int32_t *out_ids, float *out_dists, int32_t *out_sizes);
/* Device-resident form; results are packed keys-free arrays on device. */
int gulon_pq_query_dev(gulon_index_t ix, const float *dqueries, int64_t nq, int64_t ldq, int32_t k,
                       int64_t from, int64_t until, int32_t normalize, int64_t id_offset,
                       int32_t *d_ids, float *d_dists, int32_t *d_sizes, void *stream);

/*
 * TopKHeap#merge across shards (G/TopKHeap.scala:44-53; T/TopKHeapSpec.scala:33-52): merges S
 * result sets laid out [S][nq][k] (as produced by an all-gather of per-rank gulon_pq_query_dev
 * outputs) into [nq][k] by (distance, id).  Device pointers.
 *
 * Threading: handles may be used from several threads and streams.  The scratch memory inside a
 * handle (and the per-device scratch of this merge) serves one call at a time: calls are serialised
 * by a mutex while they enqueue, and a call that arrives on a DIFFERENT stream than the previous one
 * first waits on the device (cudaStreamWaitEvent) for the previous call's last kernel.  Concurrent
 * callers therefore get correct results but no overlap on one handle; use one index handle per
 * stream for overlap.  A handle belongs to the device that was current when it was created.
 */
int gulon_topk_merge_dev(const int32_t *d_ids, const float *d_dists, int32_t S, int64_t nq,
                         int32_t k, int32_t *d_out_ids, float *d_out_dists, int32_t *d_out_sizes,
                         void *stream);

/*
 * PQIndex#batchQuery over an index whose rows are sharded across processes (one process per GPU).
 * The reference merges per-range heaps in GroupedIndex#query (G/Index.scala:273-281) with
 * TopKHeap#merge (G/TopKHeap.scala:44-53); this entry point is that merge across GPUs:
 *   1. this rank scans ITS row shard (`ix`; global row id = local id + row_offset) for the queries of
 *      its query group;
 *   2. row_comm->allgather exchanges the k candidates per query between the ranks that hold the row
 *      shards of one copy of the index (80 B per query, rank and k = 10); every rank merges them by
 *      (distance, id);
 *   3. query_comm->allgather (optional) assembles the batch when the ranks are also split into query
 *      groups: group g = query_comm->rank answers queries [g*per, (g+1)*per), per = ceil(nq / groups).
 * row_comm / query_comm may be NULL (or world == 1) when that dimension is not split.  EVERY rank
 * passes the same batch and receives the whole answer.  The hooks receive device buffers and the
 * stream; a failing hook gives GULON_ECOMM.  _dev: device pointers, asynchronous on `stream`.
 * The host form copies only this rank's query slice to the device.
 */
int gulon_pq_query_sharded_dev(gulon_index_t ix, const gulon_comm_t *row_comm,
                               const gulon_comm_t *query_comm, const float *dqueries, int64_t nq,
                               int64_t ldq, int32_t k, int32_t normalize, int64_t row_offset,
                               int32_t *d_ids, float *d_dists, int32_t *d_sizes, void *stream);
int gulon_pq_query_sharded(gulon_index_t ix, const gulon_comm_t *row_comm,
                           const gulon_comm_t *query_comm, const float *queries, int64_t nq,
                           int64_t ldq, int32_t k, int32_t normalize, int64_t row_offset,
                           int32_t *out_ids, float *out_dists, int32_t *out_sizes);

/*
 * GroupedIndex#query, G/Index.scala:267-283, for a whole batch in ONE launch: the host decides which
 * partitions every query probes (GroupedIndex#searchSpace, :285-299, e.g. with gulon_exact_topk over
 * the coarse centroids) and hands over the work list of (query, partition, probe rank) triples; one
 * CTA per pair subtracts the partition's centroid from the query (MathUtils.subtract), rebuilds the
 * lookup table (Index.prepareQuery), scans the partition's rows [bounds[p], bounds[p+1]) and keeps the
 * pair's k best; the pairs of a query are merged by (distance, id) (TopKHeap#merge).  Row ids are
 * positions in the grouped order.  d_centroids [n_partitions][D]; d_bounds [n_partitions + 1];
 * pair_slot < slots is the place of the pair among its query's probes.  Queries must already be
 * normalised for a cosine index.  1 <= k <= 1024.
 */
int gulon_grouped_query_dev(gulon_index_t ix, const float *dqueries, int64_t nq, int64_t ldq,
                            const float *d_centroids, int32_t n_partitions, const int32_t *d_bounds,
                            const int32_t *d_pair_query, const int32_t *d_pair_partition,
                            const int32_t *d_pair_slot, int64_t n_pairs, int32_t slots, int32_t k,
                            int32_t *d_ids, float *d_dists, int32_t *d_sizes, void *stream);

/* Index.exactNearestNeighbours, G/Index.scala:209-229 (+ MathUtils.distanceSq, G/MathUtils.scala:85-95). */
int gulon_exact_topk(gulon_points_t p, const float *queries, int64_t nq, int64_t ldq, int32_t k,
                     int64_t from, int64_t until, int32_t *out_ids, float *out_dists,
                     int32_t *out_sizes);

/*
 * Exact fp32 re-rank of PQ candidates: for each query the candidate rows cand_ids[q][0..R) (ids < 0
 * are skipped) are scored with MathUtils.distanceSq (G/MathUtils.scala:85-95) against the raw
 * vectors and the k best kept, (distance, id) ascending.  This is the exact-distance step of the
 * recall harness (G/Tests.scala:24-37) used as a production re-rank.  Host buffers.
 */
int gulon_rerank(gulon_points_t p, const float *queries, int64_t nq, int64_t ldq,
                 const int32_t *cand_ids, int32_t R, int32_t k, int32_t *out_ids,
                 float *out_dists, int32_t *out_sizes);

/* Device-resident re-rank.  cand ids are GLOBAL row ids; this device's points are global rows
 * [id_lo, id_lo + N): candidates outside (another shard's rows, or -1) are skipped.  One fused kernel
 * per query (exact distances + top-k in shared memory); R <= 1024. */
int gulon_rerank_dev(gulon_points_t p, const float *dqueries, int64_t nq, int64_t ldq,
                     const int32_t *d_cand_ids, int32_t R, int32_t k, int64_t id_lo, int32_t *d_ids,
                     float *d_dists, int32_t *d_sizes, void *stream);

/*
 * The re-ranked query as ONE call (BASELINE configs[4]: "PQ candidates + brute-force fp32 re-rank"):
 *   1. PQIndex#batchQuery for the R best PQ candidates of every query (G/Index.scala:414-440); with
 *      row_comm the index is row-sharded and the candidates are the GLOBAL top-R (as
 *      gulon_pq_query_sharded);
 *   2. MathUtils.distanceSq (G/MathUtils.scala:85-95) between the query and the raw vector of every
 *      candidate -- Index.exactNearestNeighbours restricted to the candidates (G/Index.scala:209-229),
 *      the exact-distance step of G/Tests.scala:24-37 -- computed by the rank that owns the row
 *      (`points` are the raw vectors of exactly the rows `ix` encodes);
 *   3. the k best by (exact distance, id); sharded: one more all-gather + merge.
 * Every rank passes the same batch and receives the whole answer.  1 <= k <= R <= 1024.
 */
int gulon_pq_rerank_query_dev(gulon_index_t ix, gulon_points_t points, const gulon_comm_t *row_comm,
                              const float *dqueries, int64_t nq, int64_t ldq, int32_t k, int32_t R,
                              int32_t normalize, int64_t row_offset, int32_t *d_ids, float *d_dists,
                              int32_t *d_sizes, void *stream);
int gulon_pq_rerank_query(gulon_index_t ix, gulon_points_t points, const gulon_comm_t *row_comm,
                          const float *queries, int64_t nq, int64_t ldq, int32_t k, int32_t R,
                          int32_t normalize, int64_t row_offset, int32_t *out_ids, float *out_dists,
                          int32_t *out_sizes);

/* ---- synthetic data (benchmark / test tooling; not part of the reference path) --------------- */
/*
 * The data sets of the benchmark configurations as a pure function of (seed, stream,
 * row, column): gulon_b200/csrc/synth_spec.h.  The CPU twin (oracle/synth.c) gives identical bits, so
 * bench.py's reference arm rebuilds on host cores the very index the GPU arm measures.  Tables:
 * d_centres [centres][latent or D], d_map [latent][D].  Rows [lo, lo+n) of stream `stream_id` ->
 * d_out [n][ld].
 */
typedef struct gulon_synth_params {
  uint64_t seed;
  int32_t D, centres, latent, nonneg;
  float noise, eps, span, inv_sqrt_latent;
} gulon_synth_params_t;
int gulon_synth_tables_dev(const gulon_synth_params_t *prm, float *d_centres, float *d_map, void *stream);
int gulon_synth_rows_dev(const gulon_synth_params_t *prm, int64_t stream_id, int64_t lo, int64_t n,
                         const float *d_centres, const float *d_map, float *d_out, int64_t ld,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GULON_B200_H */
