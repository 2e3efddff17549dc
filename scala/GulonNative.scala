// Panama (JDK 22 java.lang.foreign) binding of include/gulon_b200.h -- source a gulon maintainer adds
// under core/src/main/scala/net/tixxit/gulon/b200/.  NOT compiled in this image (no JDK / scala / sbt:
// `bench.py` records the probe in `jvm_probe`); the executable mirror of exactly these symbols is
// gulon_b200/_native.py (ctypes), which the tests drive.  See INTEGRATION.md.
package net.tixxit.gulon.b200

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.{MethodHandle, MethodHandles, MethodType}

object GulonNative {
  private val linker = Linker.nativeLinker()
  private val lib    = SymbolLookup.libraryLookup("libgulon_b200.so", Arena.global())
  private def fn(name: String, res: MemoryLayout, args: MemoryLayout*): MethodHandle =
    linker.downcallHandle(lib.find(name).get, FunctionDescriptor.of(res, args: _*))

  // ---- library / device
  val lastError      = fn("gulon_last_error", ADDRESS)
  val init           = fn("gulon_init", JAVA_INT, ADDRESS, JAVA_INT)
  val shutdown       = fn("gulon_shutdown", JAVA_INT)
  val setDevice      = fn("gulon_set_device", JAVA_INT, JAVA_INT)
  val setOption      = fn("gulon_set_option", JAVA_INT, ADDRESS, JAVA_LONG)
  val subvectors     = fn("gulon_subvectors", JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS)
  // ---- Matrix
  val pointsCreate   = fn("gulon_points_create", JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS)
  val pointsDestroy  = fn("gulon_points_destroy", JAVA_INT, ADDRESS)
  val normalize      = fn("gulon_normalize", JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS, JAVA_LONG)
  // ---- KMeans
  val kmeansAssign   = fn("gulon_kmeans_assign", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT,
                          JAVA_LONG, JAVA_INT, ADDRESS)
  val kmeansUpdate   = fn("gulon_kmeans_update", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT,
                          JAVA_INT, ADDRESS, ADDRESS)
  val kmeansInit     = fn("gulon_kmeans_init", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT,
                          ADDRESS, ADDRESS)
  val kmeansTrain    = fn("gulon_kmeans_train", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT,
                          JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS,
                          ADDRESS, ADDRESS, ADDRESS)
  // ---- ProductQuantizer
  val pqTrain        = fn("gulon_pq_train", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT,
                          JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val codebookCreate = fn("gulon_codebook_create", JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS)
  val codebookInfo   = fn("gulon_codebook_info", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val codebookExport = fn("gulon_codebook_export", JAVA_INT, ADDRESS, ADDRESS)
  val codebookDestroy = fn("gulon_codebook_destroy", JAVA_INT, ADDRESS)
  val pqEncode       = fn("gulon_pq_encode", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS)
  val pqDecode       = fn("gulon_pq_decode", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG)
  // more than 256 clusters (BytePlus coders, G/Coder.scala:142-168): ids cross unpacked, one Short each
  val pqEncode16     = fn("gulon_pq_encode16", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS)
  val pqDecode16     = fn("gulon_pq_decode16", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG)
  // ---- Index.PQIndex
  val indexCreate    = fn("gulon_index_create", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
  val indexCreate16  = fn("gulon_index_create16", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
  val indexDestroy   = fn("gulon_index_destroy", JAVA_INT, ADDRESS)
  val prepareQuery   = fn("gulon_prepare_query", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
  val pqQuery        = fn("gulon_pq_query", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT,
                          JAVA_LONG, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val pqQuerySharded = fn("gulon_pq_query_sharded", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG,
                          JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val pqRerankQuery  = fn("gulon_pq_rerank_query", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG,
                          JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val exactTopk      = fn("gulon_exact_topk", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT,
                          JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val rerank         = fn("gulon_rerank", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_INT,
                          JAVA_INT, ADDRESS, ADDRESS, ADDRESS)

  final val TieLowest = 1
  final val UpdateRunningMean = 0
  final val UpdateSum = 1

  def check(rc: Int): Unit = if (rc < 0) {
    val msg = lastError.invoke().asInstanceOf[MemorySegment].reinterpret(1024).getString(0)
    rc match {
      case -1 => throw new IllegalArgumentException(msg)   // GULON_EINVAL  (require(...))
      case -7 => throw new IllegalStateException(msg)      // GULON_ESTATE
      case -3 => throw new OutOfMemoryError(msg)           // GULON_ENOMEM
      case _  => throw new RuntimeException(s"gulon_b200 error $rc: $msg")
    }
  }

  // ---- gulon_progress_t (7 x 4 bytes + pad): quantizer, num_iterations, max_iterations, step_mean,
  //      step_stddev, converged, step_count, step_s
  val ProgressLayout: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("quantizer"), JAVA_INT.withName("num_iterations"), JAVA_INT.withName("max_iterations"),
    JAVA_FLOAT.withName("step_mean"), JAVA_FLOAT.withName("step_stddev"), JAVA_INT.withName("converged"),
    JAVA_INT.withName("step_count"), JAVA_FLOAT.withName("step_s"))

  /** The receiver of gulon_progress_fn upcalls; `onReport` runs synchronously on the calling thread. */
  trait ProgressSink { def onReport(user: MemorySegment, report: MemorySegment): Unit }

  private val onReportMH: MethodHandle = MethodHandles.lookup().findVirtual(
    classOf[ProgressSink], "onReport",
    MethodType.methodType(java.lang.Void.TYPE, classOf[MemorySegment], classOf[MemorySegment]))

  /** gulon_progress_fn: void (*)(void *user, const gulon_progress_t *report). */
  def progressStub(sink: ProgressSink, arena: Arena): MemorySegment =
    linker.upcallStub(onReportMH.bindTo(sink), FunctionDescriptor.ofVoid(ADDRESS, ADDRESS), arena)

  // ---- gulon_comm_t: { int32 rank, world; 3 hooks; void *user; 2 optional hooks }
  val CommLayout: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("rank"), JAVA_INT.withName("world"),
    ADDRESS.withName("allreduce_sum_f32"), ADDRESS.withName("allreduce_sum_i32"), ADDRESS.withName("allgather"),
    ADDRESS.withName("user"), ADDRESS.withName("allreduce_sum_i64"), ADDRESS.withName("allreduce_max_f32"))

  /** What a JVM host wires to its collective library (NCCL through its own binding): device
    * pointers and the cudaStream_t the library works on; return 0 on success. */
  trait Collectives {
    def rank: Int
    def world: Int
    def allReduceSumF32(buf: MemorySegment, n: Long, stream: MemorySegment): Int
    def allReduceSumI32(buf: MemorySegment, n: Long, stream: MemorySegment): Int
    def allReduceSumI64(buf: MemorySegment, n: Long, stream: MemorySegment): Int
    def allReduceMaxF32(buf: MemorySegment, n: Long, stream: MemorySegment): Int
    def allGather(send: MemorySegment, recv: MemorySegment, bytesPerRank: Long, stream: MemorySegment): Int
  }

  private final class CommHooks(c: Collectives) {
    def arF32(u: MemorySegment, b: MemorySegment, n: Long, s: MemorySegment): Int = c.allReduceSumF32(b, n, s)
    def arI32(u: MemorySegment, b: MemorySegment, n: Long, s: MemorySegment): Int = c.allReduceSumI32(b, n, s)
    def arI64(u: MemorySegment, b: MemorySegment, n: Long, s: MemorySegment): Int = c.allReduceSumI64(b, n, s)
    def arMax(u: MemorySegment, b: MemorySegment, n: Long, s: MemorySegment): Int = c.allReduceMaxF32(b, n, s)
    def ag(u: MemorySegment, a: MemorySegment, b: MemorySegment, n: Long, s: MemorySegment): Int =
      c.allGather(a, b, n, s)
  }

  /** A gulon_comm_t in `arena` whose hooks call `c`. */
  def comm(c: Collectives, arena: Arena): MemorySegment = {
    val hooks = new CommHooks(c)
    val lk = MethodHandles.lookup()
    val ms = classOf[MemorySegment]
    val arT = MethodType.methodType(java.lang.Integer.TYPE, ms, ms, java.lang.Long.TYPE, ms)
    val agT = MethodType.methodType(java.lang.Integer.TYPE, ms, ms, ms, java.lang.Long.TYPE, ms)
    val arD = FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
    val agD = FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
    def stub(name: String, t: MethodType, d: FunctionDescriptor): MemorySegment =
      linker.upcallStub(lk.findVirtual(classOf[CommHooks], name, t).bindTo(hooks), d, arena)
    val seg = arena.allocate(CommLayout)
    seg.set(JAVA_INT, 0, c.rank)
    seg.set(JAVA_INT, 4, c.world)
    seg.set(ADDRESS, 8, stub("arF32", arT, arD))
    seg.set(ADDRESS, 16, stub("arI32", arT, arD))
    seg.set(ADDRESS, 24, stub("ag", agT, agD))
    seg.set(ADDRESS, 32, MemorySegment.NULL)
    seg.set(ADDRESS, 40, stub("arI64", arT, arD))
    seg.set(ADDRESS, 48, stub("arMax", arT, arD))
    seg
  }

  /** Matrix.data (jagged Array[Array[Float]], G/Matrix.scala:3) -> one float[rows][cols] block. */
  def flatten(rows: Array[Array[Float]], cols: Int, arena: Arena): MemorySegment = {
    val flat = arena.allocate(4L * math.max(1, rows.length) * cols, 64)
    var i = 0
    while (i < rows.length) {
      MemorySegment.copy(rows(i), 0, flat, JAVA_FLOAT, 4L * i * cols, cols)
      i += 1
    }
    flat
  }
}
