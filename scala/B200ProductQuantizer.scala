// Facade with the reference's signatures over libgulon_b200.so -- source a gulon maintainer adds.
// NOT compiled in this image (no JDK / scala / sbt); see INTEGRATION.md.  Every method names the
// reference method it stands in for; the Python mirror (gulon_b200/quantizer.py, index.py) is the
// executable twin the parity tests drive.
package net.tixxit.gulon.b200

import cats.effect.{ContextShift, IO}
import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import net.tixxit.gulon._
import GulonNative._
import scala.util.Using

object B200ProductQuantizer {

  /** gulon_codebook_info + gulon_codebook_export -> ProductQuantizer(numClusters, Vector[Quantizer(from,
    * KMeans(dim, centroids))]), the windows given by the Vectors.subvectors split rule. */
  private[b200] def fromCodebookHandle(cb: MemorySegment, arena: Arena): ProductQuantizer = {
    val info = arena.allocate(JAVA_INT, 4)               // D, M, K, dmax
    check(codebookInfo.invoke(cb, info, info.asSlice(4), info.asSlice(8), info.asSlice(12)).asInstanceOf[Int])
    val d = info.getAtIndex(JAVA_INT, 0); val m = info.getAtIndex(JAVA_INT, 1)
    val k = info.getAtIndex(JAVA_INT, 2); val dmax = info.getAtIndex(JAVA_INT, 3)
    val from = arena.allocate(JAVA_INT, m.toLong); val dim = arena.allocate(JAVA_INT, m.toLong)
    check(subvectors.invoke(d, m, from, dim).asInstanceOf[Int])
    val flat = arena.allocate(JAVA_FLOAT, m.toLong * k * dmax)
    check(codebookExport.invoke(cb, flat).asInstanceOf[Int])
    val quantizers = Vector.tabulate(m) { q =>
      val w = dim.getAtIndex(JAVA_INT, q.toLong)
      val centroids = Array.tabulate(k) { c =>
        val row = new Array[Float](w)
        MemorySegment.copy(flat, JAVA_FLOAT, 4L * ((q.toLong * k + c) * dmax), row, 0, w)
        row
      }
      ProductQuantizer.Quantizer(from.getAtIndex(JAVA_INT, q.toLong), KMeans(w, centroids))
    }
    ProductQuantizer(k, quantizers)
  }

  /** ProductQuantizer -> gulon_codebook_create (float[M][K][dmax], zero padded). */
  private[b200] def toCodebookHandle(pq: ProductQuantizer, arena: Arena): MemorySegment = {
    val m = pq.quantizers.size; val k = pq.numClusters
    val dmax = (pq.dimension + m - 1) / m
    val flat = arena.allocate(JAVA_FLOAT, m.toLong * k * dmax)
    flat.fill(0.toByte)
    pq.quantizers.zipWithIndex.foreach { case (q, i) =>
      var c = 0
      while (c < k) {
        val row = q.clusters.centroids(c)
        MemorySegment.copy(row, 0, flat, JAVA_FLOAT, 4L * ((i.toLong * k + c) * dmax), row.length)
        c += 1
      }
    }
    val out = arena.allocate(ADDRESS)
    check(codebookCreate.invoke(pq.dimension, m, k, flat, out).asInstanceOf[Int])
    out.get(ADDRESS, 0)
  }

  /** The ProgressReport upcall: KMeans.ProgressReport per quantizer (G/KMeans.scala:119-127) folded into
    * ProductQuantizer.ProgressReport (G/ProductQuantizer.scala:113-119) exactly as fromSubvectors does
    * with its Ref, then handed to config.report on the calling thread. */
  private final class Reports(config: ProductQuantizer.Config) extends ProgressSink {
    private var reports = Vector.fill(config.numQuantizers)(KMeans.ProgressReport.init(config.maxIterations))
    def onReport(user: MemorySegment, r0: MemorySegment): Unit = {
      val r = r0.reinterpret(ProgressLayout.byteSize)
      val stats = SummaryStats(r.get(JAVA_INT, 24), r.get(JAVA_FLOAT, 12), r.get(JAVA_FLOAT, 28))
      val report = KMeans.ProgressReport(r.get(JAVA_INT, 4), r.get(JAVA_INT, 8), stats, r.get(JAVA_INT, 20) != 0)
      reports = reports.updated(r.get(JAVA_INT, 0), report)
      config.report(ProductQuantizer.ProgressReport(reports)).unsafeRunSync()
    }
  }

  /** ProductQuantizer.apply(vectors, config), G/ProductQuantizer.scala:150-153: M independent k-means with
    * seed = quantizer index, all M windows trained together on the device.  Ties to the lowest centroid
    * index; the literal running-mean update (bit-exact with the reference given equal assignments). */
  def apply(vectors: Matrix, config: ProductQuantizer.Config)
           (implicit cs: ContextShift[IO]): IO[ProductQuantizer] = IO.shift *> IO.delay {
    Using.resource(Arena.ofConfined()) { arena =>
      val flat = flatten(vectors.data, vectors.cols, arena)
      val pts = arena.allocate(ADDRESS); val cb = arena.allocate(ADDRESS)
      check(pointsCreate.invoke(flat, vectors.rows.toLong, vectors.cols, vectors.cols.toLong, pts).asInstanceOf[Int])
      try {
        check(pqTrain.invoke(pts.get(ADDRESS, 0), config.numQuantizers, config.numClusters,
                             config.maxIterations, TieLowest, UpdateRunningMean,
                             MemorySegment.NULL, 0L, 0L, progressStub(new Reports(config), arena),
                             MemorySegment.NULL, cb).asInstanceOf[Int])
        try fromCodebookHandle(cb.get(ADDRESS, 0), arena)
        finally codebookDestroy.invoke(cb.get(ADDRESS, 0))
      } finally pointsDestroy.invoke(pts.get(ADDRESS, 0))
    }
  }

  /** ProductQuantizer#encode, G/ProductQuantizer.scala:25-35: one assign per quantizer, then
    * coder.buildCode.  The library returns plane-major ids (uint8 [M][N], or uint16 for the BytePlus
    * coders); each plane is packed with the reference's own coder so the EncodedMatrix is what
    * `encode` would have produced (including Coder2 / Coder4 bit packing). */
  def encode(pq: ProductQuantizer, vectors: Matrix)
            (implicit cs: ContextShift[IO]): IO[EncodedMatrix] = IO.shift *> IO.delay {
    Using.resource(Arena.ofConfined()) { arena =>
      val n = vectors.rows; val m = pq.quantizers.size
      val coder = pq.coderFactory(n)
      val cb = toCodebookHandle(pq, arena)
      try {
        val flat = flatten(vectors.data, vectors.cols, arena)
        val wide = pq.numClusters > 256
        val codes = arena.allocate((if (wide) 2L else 1L) * m * math.max(1, n), 64)
        val call = if (wide) pqEncode16 else pqEncode
        check(call.invoke(cb, flat, n.toLong, vectors.cols.toLong, TieLowest, codes).asInstanceOf[Int])
        val planes = Vector.tabulate(m) { q =>
          val idx = new Array[Int](n)
          var i = 0
          while (i < n) {
            idx(i) = if (wide) codes.get(JAVA_SHORT, 2L * (q.toLong * n + i)) & 0xffff
                     else codes.get(JAVA_BYTE, q.toLong * n + i) & 0xff
            i += 1
          }
          coder.buildCode(idx)
        }
        EncodedMatrix(coder)(planes)
      } finally codebookDestroy.invoke(cb)
    }
  }
}

/** PQIndex(productQuantizer, data), G/Index.scala:385-441, with the code planes resident in HBM.
  * Owns a codebook and an index handle; `close()` frees them. */
final class B200PQIndex(val productQuantizer: ProductQuantizer, val data: EncodedMatrix) extends AutoCloseable {
  private val arena = Arena.ofShared()
  private val m = productQuantizer.quantizers.size
  private val n = data.length
  private val wide = productQuantizer.numClusters > 256
  private val cb: MemorySegment = B200ProductQuantizer.toCodebookHandle(productQuantizer, arena)
  private val handle: MemorySegment = {
    // unpack every plane to one id per (quantizer, row): what the kernels read
    val codes = arena.allocate((if (wide) 2L else 1L) * m * math.max(1, n), 64)
    var q = 0
    while (q < m) {
      val code = data.encodings(q)
      var i = 0
      while (i < n) {
        val id = data.coder.getIndex(code, i)
        if (wide) codes.set(JAVA_SHORT, 2L * (q.toLong * n + i), id.toShort)
        else codes.set(JAVA_BYTE, q.toLong * n + i, id.toByte)
        i += 1
      }
      q += 1
    }
    val out = arena.allocate(ADDRESS)
    val call = if (wide) indexCreate16 else indexCreate
    check(call.invoke(cb, codes, n.toLong, n.toLong, out).asInstanceOf[Int])
    out.get(ADDRESS, 0)
  }

  def dimension: Int = productQuantizer.dimension

  /** ids / distances [Q][k] ascending -> one TopKHeap per query.  A max-heap laid out in DEscending
    * order is a valid heap (every parent >= its children), so `fromHeap` / `merge` / `update` of the
    * reference keep working on the result (G/TopKHeap.scala:3-94, G/Index.scala:83-94). */
  private def heaps(k: Int, nq: Int, ids: MemorySegment, dists: MemorySegment, sizes: MemorySegment): Vector[TopKHeap] =
    Vector.tabulate(nq) { q =>
      val sz = sizes.getAtIndex(JAVA_INT, q.toLong)
      val keys = new Array[Int](k); val values = new Array[Float](k)
      var i = 0
      while (i < sz) {
        keys(i) = ids.getAtIndex(JAVA_INT, q.toLong * k + (sz - 1 - i))
        values(i) = dists.getAtIndex(JAVA_FLOAT, q.toLong * k + (sz - 1 - i))
        i += 1
      }
      val heap = new TopKHeap(keys, values)
      heap.size = sz
      heap
    }

  /** PQIndex#batchQuery(k, vectors, from, until), G/Index.scala:414-440 (same require(...) behaviour:
    * GULON_EINVAL -> IllegalArgumentException). */
  def batchQuery(k: Int, vectors: Matrix, from: Int, until: Int): Vector[TopKHeap] =
    Using.resource(Arena.ofConfined()) { a =>
      val nq = vectors.rows
      val q = flatten(vectors.data, vectors.cols, a)
      val ids = a.allocate(JAVA_INT, math.max(1L, nq.toLong * k))
      val dists = a.allocate(JAVA_FLOAT, math.max(1L, nq.toLong * k))
      val sizes = a.allocate(JAVA_INT, math.max(1L, nq.toLong))
      check(pqQuery.invoke(handle, q, nq.toLong, vectors.cols.toLong, k, from.toLong, until.toLong, 0, 0L,
                           ids, dists, sizes).asInstanceOf[Int])
      heaps(k, nq, ids, dists, sizes)
    }

  /** PQIndex#query(k, query, from, until), G/Index.scala:411-412. */
  def query(k: Int, query: Array[Float], from: Int, until: Int): TopKHeap =
    batchQuery(k, Matrix(1, query.length, Array(query)), from, until).head

  /** Row-sharded form (one process per GPU; this index holds rows [rowOffset, rowOffset + length) of the
    * whole): gulon_pq_query_sharded drives local scan -> allgather hook -> (distance, id) merge.
    * `rowComm` / `queryComm` come from GulonNative.comm(...); either may be MemorySegment.NULL. */
  def batchQuerySharded(k: Int, vectors: Matrix, rowComm: MemorySegment, queryComm: MemorySegment,
                        rowOffset: Long): Vector[TopKHeap] =
    Using.resource(Arena.ofConfined()) { a =>
      val nq = vectors.rows
      val q = flatten(vectors.data, vectors.cols, a)
      val ids = a.allocate(JAVA_INT, math.max(1L, nq.toLong * k))
      val dists = a.allocate(JAVA_FLOAT, math.max(1L, nq.toLong * k))
      val sizes = a.allocate(JAVA_INT, math.max(1L, nq.toLong))
      check(pqQuerySharded.invoke(handle, rowComm, queryComm, q, nq.toLong, vectors.cols.toLong, k, 0,
                                  rowOffset, ids, dists, sizes).asInstanceOf[Int])
      heaps(k, nq, ids, dists, sizes)
    }

  def close(): Unit = {
    indexDestroy.invoke(handle)
    codebookDestroy.invoke(cb)
    arena.close()
  }
}
