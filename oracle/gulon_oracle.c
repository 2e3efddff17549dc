/*
 * gulon_oracle.c -- CPU restatement of tixxit/gulon's product-quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gulon_b200/ may include, link, load or call this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker or as the reported CPU baseline.
 *
 * PARITY UNPINNED: the reference is Scala on the JVM; this image has no JDK/scala/sbt, so the
 * reference itself cannot run here, and its own tests hold no golden vectors or known answers
 * for this path (all ScalaCheck properties).  This restatement is pinned only by
 *   (a) the JDK-documented java.util.Random known answers (see tests/test_oracle_kat.py),
 *   (b) an independent numpy-float32 restatement (oracle/np_oracle.py) that must agree bit for
 *       bit on random inputs, and
 *   (c) re-statements of the reference's own property tests.
 * Fidelity therefore rests on line-by-line source correspondence; each function cites the
 * reference lines it follows.  Path alias: G/ = core/src/main/scala/net/tixxit/gulon/.
 *
 * Numeric model (JVM): strict IEEE-754 binary32, round-to-nearest-even, no FMA contraction, no
 * reassociation.  Build with -O2 -ffp-contract=off -fno-fast-math -fexcess-precision=standard
 * on x86-64 (SSE scalar float), see oracle/Makefile.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GO_TIE_LITERAL 0   /* reference: rng.nextBoolean() on exact ties (G/KMeans.scala:47,90) */
#define GO_TIE_LOWEST  1   /* canonical: keep the lowest index on exact ties                     */

#define GO_TOPK_LITERAL   0 /* reference TopKHeap, literal binary-heap behaviour                 */
#define GO_TOPK_CANONICAL 1 /* (distance asc, id asc) lexicographic                              */

/* ------------------------------------------------------------------------------------------ */
/* java.util.Random (JDK javadoc algorithm); scala.util.Random(seed:Int) wraps it.            */
/* Call sites: G/KMeans.scala:28,47,71,90,189-191; G/Tests.scala:82-84.                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t s; } go_jrandom;

#define JR_MULT 0x5DEECE66DULL
#define JR_MASK ((1ULL << 48) - 1)

void go_jr_init(go_jrandom *r, int64_t seed) { r->s = ((uint64_t)seed ^ JR_MULT) & JR_MASK; }

static inline int32_t jr_next(go_jrandom *r, int bits) {
  r->s = (r->s * JR_MULT + 0xBULL) & JR_MASK;
  return (int32_t)(int64_t)(r->s >> (48 - bits)); /* (int)(seed >>> (48 - bits)) */
}

int32_t go_jr_next_int(go_jrandom *r) { return jr_next(r, 32); }

int go_jr_next_boolean(go_jrandom *r) { return jr_next(r, 1) != 0; }

float go_jr_next_float(go_jrandom *r) { return (float)jr_next(r, 24) / (float)(1 << 24); }

int32_t go_jr_next_int_bound(go_jrandom *r, int32_t bound) {
  /* Random.nextInt(int bound): power-of-two shortcut, else rejection loop with int wrap-around */
  int32_t rr = jr_next(r, 31);
  int32_t m = bound - 1;
  if ((bound & m) == 0) {
    rr = (int32_t)(((int64_t)bound * (int64_t)rr) >> 31);
  } else {
    int32_t u = rr;
    for (;;) {
      rr = u % bound;
      /* u - rr + m < 0 evaluated with 32-bit wrap */
      int32_t t = (int32_t)((uint32_t)u - (uint32_t)rr + (uint32_t)m);
      if (t >= 0) break;
      u = jr_next(r, 31);
    }
  }
  return rr;
}

/* ------------------------------------------------------------------------------------------ */
/* Vectors.subvectors split rule, G/Vectors.scala:84-104.                                      */
/* ------------------------------------------------------------------------------------------ */
int go_subvectors(int D, int M, int32_t *from, int32_t *dim) {
  if (M <= 0) return -1;
  int ideal = (D + M - 1) / M;
  int shortfall = ideal * M - D;
  int full = M - shortfall;
  for (int i = 0; i < M; i++) {
    if (i < full) { from[i] = i * ideal; dim[i] = ideal; }
    else { from[i] = full * ideal + (i - full) * (ideal - 1); dim[i] = ideal - 1; }
  }
  return ideal;
}

/* KMeans.apply offsets, G/KMeans.scala:170-186: off_k = sum_j c_kj^2, sequential, unfused.    */
void go_offsets(const float *C, int K, int dim, int64_t ldc, float *off) {
  for (int k = 0; k < K; k++) {
    const float *c = C + (int64_t)k * ldc;
    float s = 0.0f;
    for (int j = 0; j < dim; j++) { float x = c[j]; s += x * x; }
    off[k] = s;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* KMeans.assign, G/KMeans.scala:24-55 (ranged) and :70-98 (whole array): one RNG per call.    */
/* `assign` is in/out: rows where no centroid is accepted keep their previous content.         */
/* tie_stats[0] += tie events (rng draws in literal mode); tie_stats[1] += rows whose literal  */
/* result differs from the lowest-index result.                                                */
/* ------------------------------------------------------------------------------------------ */
void go_assign_range(const float *X, int64_t ld, int from, int dim, const float *C, int64_t ldc,
                     const float *off, int K, int64_t start, int64_t end, int tie_mode,
                     int32_t *assign, int64_t *tie_stats) {
  go_jrandom rng;
  go_jr_init(&rng, 0);
  int64_t events = 0, flips = 0;
  for (int64_t i = start; i < end; i++) {
    const float *row = X + i * ld + from;
    float min = FLT_MAX;
    int lit = -1, low = -1;
    for (int k = 0; k < K; k++) {
      const float *c = C + (int64_t)k * ldc;
      float d = 0.0f;
      for (int j = 0; j < dim; j++) d += row[j] * c[j];
      d = off[k] - 2 * d;
      if (d < min) { lit = k; low = k; min = d; }
      else if (d == min) {
        events++;
        if (go_jr_next_boolean(&rng)) lit = k; /* min unchanged */
      }
    }
    if (lit != low) flips++;
    int pick = (tie_mode == GO_TIE_LITERAL) ? lit : low;
    if (pick >= 0) assign[i] = pick;
  }
  if (tie_stats) {
#ifdef _OPENMP
#pragma omp atomic
#endif
    tie_stats[0] += events;
#ifdef _OPENMP
#pragma omp atomic
#endif
    tie_stats[1] += flips;
  }
}

/* KMeans.parAssign, G/KMeans.scala:57-68: 25 000-row batches, each with a fresh Random(0).    */
/* batch <= 0 means the whole-array form (one RNG for all N rows, G/KMeans.scala:70-98).       */
void go_assign(const float *X, int64_t N, int64_t ld, int from, int dim, const float *C,
               int64_t ldc, int K, int64_t batch, int tie_mode, int nthreads, int32_t *assign,
               int64_t *tie_stats) {
  float *off = (float *)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
  go_offsets(C, K, dim, ldc, off);
  if (batch <= 0) {
    go_assign_range(X, ld, from, dim, C, ldc, off, K, 0, N, tie_mode, assign, tie_stats);
  } else {
    int64_t nb = (N + batch - 1) / batch;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t b = 0; b < nb; b++) {
      int64_t s = b * batch, e = s + batch < N ? s + batch : N;
      go_assign_range(X, ld, from, dim, C, ldc, off, K, s, e, tie_mode, assign, tie_stats);
    }
  }
  free(off);
}

/* KMeans.fromAssignment, G/KMeans.scala:198-226: running mean, rows in order.                 */
void go_from_assignment(const float *X, int64_t N, int64_t ld, int from, int dim,
                        const int32_t *assign, int K, float *C, int64_t ldc, int32_t *counts_out) {
  int32_t *counts = (int32_t *)calloc((size_t)(K > 0 ? K : 1), sizeof(int32_t));
  for (int k = 0; k < K; k++) memset(C + (int64_t)k * ldc, 0, sizeof(float) * (size_t)dim);
  for (int64_t i = 0; i < N; i++) {
    const float *v = X + i * ld + from;
    int k = assign[i];
    float *c = C + (int64_t)k * ldc;
    int32_t n = counts[k] + 1;
    float fn = (float)n; /* Int -> Float conversion of the divisor */
    for (int j = 0; j < dim; j++) {
      float p = c[j];
      c[j] = p + ((v[j] - p) / fn);
    }
    counts[k] = n;
  }
  if (counts_out) memcpy(counts_out, counts, sizeof(int32_t) * (size_t)K);
  free(counts);
}

/* KMeans.init, G/KMeans.scala:188-196: k rows sampled WITH replacement, Random(seed).         */
void go_kmeans_init(const float *X, int64_t N, int64_t ld, int from, int dim, int K, int32_t seed,
                    float *C, int64_t ldc, int32_t *rows_out) {
  go_jrandom rng;
  go_jr_init(&rng, (int64_t)seed);
  for (int k = 0; k < K; k++) {
    int32_t i = go_jr_next_int_bound(&rng, (int32_t)N);
    if (rows_out) rows_out[k] = i;
    memcpy(C + (int64_t)k * ldc, X + (int64_t)i * ld + from, sizeof(float) * (size_t)dim);
  }
}

/* MathUtils.distance(x, y), G/MathUtils.scala:85-98: (float)sqrt((double)sum (y-x)^2)         */
static float go_distance2(const float *x, const float *y, int dim) {
  float s = 0.0f;
  for (int i = 0; i < dim; i++) { float dx = y[i] - x[i]; s += dx * dx; }
  return (float)sqrt((double)s);
}

/* KMeans.stepSize + SummaryStatsBuilder, G/KMeans.scala:160-168, G/MathUtils.scala:46-57.     */
static void go_step_size(const float *prev, const float *next, int K, int dim, int64_t ldc,
                         float *mean, float *stddev) {
  float m = 0.0f, s = 0.0f;
  int n = 0;
  for (int i = 0; i < K; i++) {
    float x = go_distance2(prev + (int64_t)i * ldc, next + (int64_t)i * ldc, dim);
    n += 1;
    float m0 = m;
    m = m0 + (x - m0) / (float)n;
    s = s + (x - m0) * (x - m);
  }
  *mean = m;
  *stddev = n > 0 ? (float)sqrt((double)(s / (float)n)) : 0.0f;
}

/*
 * KMeans.computeClusters, G/KMeans.scala:134-157.  init -> parAssign; then for i = 0..maxIter
 * inclusive: fromAssignment -> parAssign -> converged = Arrays.equals(prev, next) -> stop.
 * Returns the last UPDATED centroids.  report[i] = {iteration index, step mean, step stddev,
 * converged} for every loop pass (report may be NULL; capacity max_iter + 1 rows of 4 floats).
 * n_updates receives the number of fromAssignment calls performed.
 */
void go_compute_clusters(const float *X, int64_t N, int64_t ld, int from, int dim, int K,
                         int max_iter, int32_t seed, int tie_mode, int nthreads, float *C,
                         int64_t ldc, int32_t *n_updates, int32_t *converged_out, float *report,
                         int32_t *final_assign, int64_t *tie_stats) {
  float *prevC = (float *)malloc(sizeof(float) * (size_t)K * (size_t)ldc);
  int32_t *prev = (int32_t *)calloc((size_t)(N > 0 ? N : 1), sizeof(int32_t));
  int32_t *next = (int32_t *)calloc((size_t)(N > 0 ? N : 1), sizeof(int32_t));
  go_kmeans_init(X, N, ld, from, dim, K, seed, C, ldc, NULL);
  go_assign(X, N, ld, from, dim, C, ldc, K, 25000, tie_mode, nthreads, prev, tie_stats);
  int i = 0, updates = 0, conv = 0;
  while (i <= max_iter) {
    memcpy(prevC, C, sizeof(float) * (size_t)K * (size_t)ldc);
    go_from_assignment(X, N, ld, from, dim, prev, K, C, ldc, NULL);
    memset(next, 0, sizeof(int32_t) * (size_t)N); /* parAssign allocates a fresh array */
    go_assign(X, N, ld, from, dim, C, ldc, K, 25000, tie_mode, nthreads, next, tie_stats);
    conv = memcmp(prev, next, sizeof(int32_t) * (size_t)N) == 0;
    if (report) {
      float mean, sd;
      go_step_size(prevC, C, K, dim, ldc, &mean, &sd);
      report[4 * updates + 0] = (float)i;
      report[4 * updates + 1] = mean;
      report[4 * updates + 2] = sd;
      report[4 * updates + 3] = (float)conv;
    }
    updates++;
    int32_t *t = prev; prev = next; next = t;
    i = conv ? max_iter + 1 : i + 1;
  }
  if (n_updates) *n_updates = updates;
  if (converged_out) *converged_out = conv;
  if (final_assign) memcpy(final_assign, prev, sizeof(int32_t) * (size_t)N);
  free(prevC); free(prev); free(next);
}

/*
 * ProductQuantizer.apply / fromSubvectors, G/ProductQuantizer.scala:121-153: M independent
 * computeClusters over the column windows, seed = subspace index (:139), run concurrently.
 * codebook layout: float [M][K][dmax], dmax = ceil(D/M); subspace m uses the first dim[m] floats.
 */
void go_pq_train(const float *X, int64_t N, int64_t ld, int D, int M, int K, int max_iter,
                 int tie_mode, int nthreads, float *codebook, int32_t *n_updates,
                 int32_t *converged) {
  int32_t *from = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  int32_t *dim = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  int dmax = go_subvectors(D, M, from, dim);
  (void)nthreads;
#ifdef _OPENMP
  int outer = nthreads > 0 ? (nthreads < M ? nthreads : M) : 1;
  int inner = nthreads > 0 ? (nthreads / outer > 0 ? nthreads / outer : 1) : 1;
  omp_set_max_active_levels(2);
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer)
#else
  int inner = 1;
#endif
  for (int m = 0; m < M; m++) {
    go_compute_clusters(X, N, ld, from[m], dim[m], K, max_iter, m, tie_mode, inner,
                        codebook + (int64_t)m * K * dmax, dmax, n_updates ? n_updates + m : NULL,
                        converged ? converged + m : NULL, NULL, NULL, NULL);
  }
  free(from); free(dim);
}

/* Index.prepareQuery, G/Index.scala:352-383: LUT[q][m][i] = sum_k fl(d*d), d = q[k+from]-c[k]. */
void go_prepare_query(const float *queries, int64_t Q, int64_t ldq, int D, int M, int K,
                      const float *codebook, float *lut) {
  int32_t *from = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  int32_t *dim = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  int dmax = go_subvectors(D, M, from, dim);
  for (int m = 0; m < M; m++)
    for (int i = 0; i < K; i++) {
      const float *c = codebook + ((int64_t)m * K + i) * dmax;
      for (int64_t q = 0; q < Q; q++) {
        const float *query = queries + q * ldq + from[m];
        float sumSq = 0.0f;
        for (int k = 0; k < dim[m]; k++) { float d = query[k] - c[k]; sumSq += d * d; }
        lut[(q * M + m) * K + i] = sumSq;
      }
    }
  free(from); free(dim);
}

/* ------------------------------------------------------------------------------------------ */
/* TopKHeap, G/TopKHeap.scala:3-94 (literal).                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t *keys; float *values; int cap; int size; } go_heap;

go_heap *go_heap_new(int k) {
  go_heap *h = (go_heap *)malloc(sizeof(go_heap));
  h->keys = (int32_t *)calloc((size_t)(k > 0 ? k : 1), sizeof(int32_t));
  h->values = (float *)calloc((size_t)(k > 0 ? k : 1), sizeof(float));
  h->cap = k; h->size = 0;
  return h;
}
void go_heap_free(go_heap *h) { free(h->keys); free(h->values); free(h); }
int go_heap_size(const go_heap *h) { return h->size; }
void go_heap_raw(const go_heap *h, int32_t *keys, float *values) {
  memcpy(keys, h->keys, sizeof(int32_t) * (size_t)h->size);
  memcpy(values, h->values, sizeof(float) * (size_t)h->size);
}
static void heap_swap(go_heap *h, int i, int j) {
  int32_t tk = h->keys[i]; float tv = h->values[i];
  h->keys[i] = h->keys[j]; h->values[i] = h->values[j];
  h->keys[j] = tk; h->values[j] = tv;
}
static void heap_up(go_heap *h, int i) {
  while (i > 0) {
    int p = (i - 1) / 2;
    if (h->values[i] > h->values[p]) { heap_swap(h, i, p); i = p; } else break;
  }
}
static void heap_down(go_heap *h, int i) {
  for (;;) {
    int top = i, lc = 2 * i + 1, rc = 2 * i + 2;
    if (lc < h->size && h->values[top] < h->values[lc]) top = lc;
    if (rc < h->size && h->values[top] < h->values[rc]) top = rc;
    if (top != i) { heap_swap(h, i, top); i = top; } else break;
  }
}
/* delete(), :57-67; returns -1 (and leaves the heap alone) where the reference throws. */
int go_heap_delete(go_heap *h, int32_t *removed) {
  if (h->size <= 0) return -1;
  h->size -= 1;
  if (removed) *removed = h->keys[0];
  h->keys[0] = h->keys[h->size];
  h->values[0] = h->values[h->size];
  heap_down(h, 0);
  return 0;
}
/* update(), :69-79 */
void go_heap_update(go_heap *h, int32_t k, float v) {
  if (h->size == h->cap && h->cap > 0 && h->values[0] > v) go_heap_delete(h, NULL);
  if (h->size < h->cap) {
    h->keys[h->size] = k; h->values[h->size] = v;
    heap_up(h, h->size);
    h->size += 1;
  }
}
/* merge(), :44-53: replay the other heap's ARRAY order through update */
void go_heap_merge(go_heap *h, const go_heap *that) {
  for (int i = 0; i < that->size; i++) go_heap_update(h, that->keys[i], that->values[i]);
}
/* Result.fromHeap, G/Index.scala:83-94 / deleteAll, G/TopKHeap.scala:81-89: pop max, fill from
 * the back => ascending distance.  Destroys the heap content. Returns the size. */
int go_heap_drain(go_heap *h, int32_t *ids, float *dists) {
  int n = h->size;
  for (int i = n - 1; i >= 0; i--) {
    ids[i] = h->keys[0];
    if (dists) dists[i] = h->values[0];
    go_heap_delete(h, NULL);
  }
  return n;
}

/* Canonical bounded top-k: sorted ascending by (value, key); insert keeps lexicographic order.
 * The order on values is TOTAL: every NaN ranks after +inf (all NaNs equal).  The reference heap has
 * no defined behaviour for NaN (`root > v` is false both ways, G/TopKHeap.scala:69-79: what it keeps
 * depends on the insertion order); the canonical rule, which is the GPU library's, must be a total
 * order or the result would depend on the scan order too. */
typedef struct { int32_t *keys; float *values; int cap; int size; } go_topk;
static inline int topk_less(float a, int32_t ka, float b, int32_t kb) {
  const int na = a != a, nb = b != b;
  if (na || nb) return na != nb ? nb : ka < kb;
  return a < b || (a == b && ka < kb);
}
static void topk_insert(go_topk *t, int32_t key, float v) {
  if (t->cap <= 0) return;
  if (t->size == t->cap) {
    float lv = t->values[t->size - 1]; int32_t lk = t->keys[t->size - 1];
    if (!topk_less(v, key, lv, lk)) return;
    t->size -= 1;
  }
  int i = t->size;
  while (i > 0 && topk_less(v, key, t->values[i - 1], t->keys[i - 1])) {
    t->values[i] = t->values[i - 1]; t->keys[i] = t->keys[i - 1]; i--;
  }
  t->values[i] = v; t->keys[i] = key; t->size += 1;
}

#define GO_CODE_T uint8_t
#define GO_NAME(f) f
#include "go_codes.inc"
#undef GO_CODE_T
#undef GO_NAME
#define GO_CODE_T uint16_t
#define GO_NAME(f) f##16
#include "go_codes.inc"
#undef GO_CODE_T
#undef GO_NAME

/* MathUtils.distanceSq(x, y), G/MathUtils.scala:85-95: dx = y_i - x_i; sumSq += dx*dx.        */
float go_distance_sq(const float *x, const float *y, int dim) {
  float s = 0.0f;
  for (int i = 0; i < dim; i++) { float dx = y[i] - x[i]; s += dx * dx; }
  return s;
}

/* Index.exactNearestNeighbours, G/Index.scala:209-229: heap.update(i, distanceSq(vectors(i), q)) */
void go_exact_nn(const float *X, int64_t ld, int D, int64_t from, int64_t until,
                 const float *queries, int64_t Q, int64_t ldq, int k, int topk_mode, int nthreads,
                 int32_t *ids, float *dists, int32_t *sizes) {
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int64_t q = 0; q < Q; q++) {
    const float *query = queries + q * ldq;
    go_heap *h = go_heap_new(k);
    go_topk t;
    t.keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
    t.values = (float *)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
    t.cap = k; t.size = 0;
    for (int64_t i = from; i < until; i++) {
      float d = go_distance_sq(X + i * ld, query, D);
      if (topk_mode == GO_TOPK_LITERAL) go_heap_update(h, (int32_t)i, d);
      else topk_insert(&t, (int32_t)i, d);
    }
    if (topk_mode == GO_TOPK_LITERAL) {
      sizes[q] = go_heap_drain(h, ids + q * k, dists + q * k);
    } else {
      sizes[q] = t.size;
      memcpy(ids + q * k, t.keys, sizeof(int32_t) * (size_t)t.size);
      memcpy(dists + q * k, t.values, sizeof(float) * (size_t)t.size);
    }
    go_heap_free(h); free(t.keys); free(t.values);
  }
}

/* MathUtils.normalize, G/MathUtils.scala:100-120: d = (float)sqrt((double)sum x^2); y = x / d  */
void go_normalize(const float *X, int64_t N, int64_t ld, int D, float *out, int64_t ldo) {
  for (int64_t i = 0; i < N; i++) {
    const float *x = X + i * ld;
    float sum = 0.0f;
    for (int j = 0; j < D; j++) { float v = x[j]; sum += v * v; }
    float d = (float)sqrt((double)sum);
    for (int j = 0; j < D; j++) out[i * ldo + j] = x[j] / d;
  }
}

/*
 * GroupedIndex.query + searchSpace, G/Index.scala:267-299.  strategy 0 = LimitGroups(limit),
 * 1 = LimitVectors(limit).  centroids [P][D]; offsets [P-1] (group boundaries); the residual PQ
 * index is (codebook, codes).  `query` must already be normalised if the metric is cosine.
 */
void go_grouped_query(const float *query, int D, const float *centroids, int P,
                      const int32_t *offsets, int strategy, int limit, const float *codebook,
                      int M, int K, const uint8_t *codes, int64_t N, int64_t plane_stride, int k,
                      int topk_mode, int32_t *ids, float *dists, int32_t *size,
                      int32_t *probed, int32_t *n_probed) {
  int kk = strategy == 0 ? limit : P;
  int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(kk > 0 ? kk : 1));
  float *odist = (float *)malloc(sizeof(float) * (size_t)(kk > 0 ? kk : 1));
  int32_t osz = 0;
  go_exact_nn(centroids, D, D, 0, P, query, 1, D, kk, topk_mode, 1, order, odist, &osz);
  free(odist);
  int np = osz;
  if (strategy == 1) {
    int i = 0; int64_t count = 0;
    while (i < osz && count < limit) {
      int c = order[i];
      int64_t s = c == 0 ? 0 : offsets[c - 1];
      int64_t e = c == P - 1 ? N : offsets[c];
      count += e - s; i++;
    }
    np = i;
  }
  go_heap *heap = go_heap_new(k);
  go_topk t;
  t.keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
  t.values = (float *)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
  t.cap = k; t.size = 0;
  float *residual = (float *)malloc(sizeof(float) * (size_t)D);
  float *lut = (float *)malloc(sizeof(float) * (size_t)M * (size_t)K);
  int32_t *pids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
  float *pds = (float *)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
  for (int i = 0; i < np; i++) {
    int c = order[i];
    int64_t s = c == 0 ? 0 : offsets[c - 1];
    int64_t e = c == P - 1 ? N : offsets[c];
    for (int j = 0; j < D; j++) residual[j] = query[j] - centroids[(int64_t)c * D + j];
    go_prepare_query(residual, 1, D, D, M, K, codebook, lut);
    if (topk_mode == GO_TOPK_LITERAL) {
      /* vectorIndex.query(k, residual, from, until) returns a heap; heap.merge replays its
         array order.  Rebuild that heap literally. */
      go_heap *ph = go_heap_new(k);
      float ds[4096];
      for (int64_t r0 = s; r0 < e;) {
        int bs = (int)(e - r0 < 4096 ? e - r0 : 4096);
        for (int r = 0; r < bs; r++) ds[r] = 0.0f;
        for (int j = 0; j < M; j++)
          for (int r = 0; r < bs; r++)
            ds[r] += lut[(int64_t)j * K + codes[(int64_t)j * plane_stride + r0 + r]];
        for (int r = 0; r < bs; r++) go_heap_update(ph, (int32_t)(r0 + r), ds[r]);
        r0 += bs;
      }
      go_heap_merge(heap, ph);
      go_heap_free(ph);
    } else {
      int32_t psz = 0;
      go_batch_query(lut, 1, M, K, codes, plane_stride, s, e, k, GO_TOPK_CANONICAL, 1, pids, pds,
                     &psz);
      for (int r = 0; r < psz; r++) topk_insert(&t, pids[r], pds[r]);
    }
    if (probed) probed[i] = c;
  }
  if (n_probed) *n_probed = np;
  if (topk_mode == GO_TOPK_LITERAL) {
    *size = go_heap_drain(heap, ids, dists);
  } else {
    *size = t.size;
    memcpy(ids, t.keys, sizeof(int32_t) * (size_t)t.size);
    memcpy(dists, t.values, sizeof(float) * (size_t)t.size);
  }
  go_heap_free(heap); free(t.keys); free(t.values);
  free(residual); free(lut); free(pids); free(pds); free(order);
}

/* Objective used by T/KMeansSpec.scala:40-57: sum of squared distances to assigned centroid.  */
double go_objective(const float *X, int64_t N, int64_t ld, int from, int dim, const float *C,
                    int64_t ldc, const int32_t *assign) {
  double tot = 0.0;
  for (int64_t i = 0; i < N; i++)
    tot += (double)go_distance_sq(X + i * ld + from, C + (int64_t)assign[i] * ldc, dim);
  return tot;
}

int go_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
