// Not compiled in this image (no JDK/sbt); see INTEGRATION.md.
package net.tixxit.gulon.b200

import cats.effect.{ContextShift, IO}
import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import net.tixxit.gulon._
import GulonNative._
import scala.util.Using

// Drop-in for ProductQuantizer.apply / #encode and PQIndex#batchQuery: same signatures, IO-wrapped
// blocking native calls (the reference's parallel sections need a ContextShift; here the
// parallelism is on the device, so the IO just shifts to a blocking pool).
object B200ProductQuantizer {
  def apply(vectors: Matrix, config: ProductQuantizer.Config)
           (implicit cs: ContextShift[IO]): IO[ProductQuantizer] = IO.shift *> IO.delay {
    Using.resource(Arena.ofConfined()) { arena =>
      val flat = arena.allocate(4L * vectors.rows * vectors.cols, 64)
      var i = 0
      while (i < vectors.rows) {                       // flatten the jagged Array[Array[Float]] once
        MemorySegment.copy(vectors.data(i), 0, flat, JAVA_FLOAT, 4L * i * vectors.cols, vectors.cols)
        i += 1
      }
      val pts = arena.allocate(ADDRESS); val cb = arena.allocate(ADDRESS)
      check(pointsCreate.invoke(flat, vectors.rows.toLong, vectors.cols, vectors.cols.toLong, pts).asInstanceOf[Int])
      try check(pqTrain.invoke(pts.get(ADDRESS, 0), config.numQuantizers, config.numClusters,
                               config.maxIterations, /*GULON_TIE_LOWEST*/ 1, /*RUNNING_MEAN*/ 0,
                               MemorySegment.NULL, 0L, 0L, progressStub(config.report), MemorySegment.NULL,
                               cb).asInstanceOf[Int])
      finally pointsDestroy.invoke(pts.get(ADDRESS, 0))
      fromCodebookHandle(cb.get(ADDRESS, 0))           // gulon_codebook_export -> Vector[Quantizer(from, KMeans)]
    }
  }
}

final class B200PQIndex(pq: ProductQuantizer, data: EncodedMatrix) {
  // gulon_codebook_create + gulon_index_create once; codes stay resident in HBM
  def batchQuery(k: Int, vectors: Matrix, from: Int, until: Int): Vector[TopKHeap] = {
    // gulon_pq_query(handle, flatQueries, Q, D, k, from, until, normalize = 0, idOffset = 0, ids, dists, sizes)
    // -> rebuild TopKHeap(keys, values, size) per query from ids/dists (ascending)
    ???
  }
}
