// Not compiled in this image (no JDK/sbt); see INTEGRATION.md.
package net.tixxit.gulon.b200

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.MethodHandle

object GulonNative {
  private val linker = Linker.nativeLinker()
  private val lib    = SymbolLookup.libraryLookup("libgulon_b200.so", Arena.global())
  private def fn(name: String, res: MemoryLayout, args: MemoryLayout*): MethodHandle =
    linker.downcallHandle(lib.find(name).get, FunctionDescriptor.of(res, args: _*))

  val lastError     = fn("gulon_last_error", ADDRESS)
  val pointsCreate  = fn("gulon_points_create", JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS)
  val pointsDestroy = fn("gulon_points_destroy", JAVA_INT, ADDRESS)
  val pqTrain       = fn("gulon_pq_train", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT,
                         JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val codebookCreate = fn("gulon_codebook_create", JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS)
  val codebookExport = fn("gulon_codebook_export", JAVA_INT, ADDRESS, ADDRESS)
  val pqEncode      = fn("gulon_pq_encode", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS)
  val indexCreate   = fn("gulon_index_create", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
  // more than 256 clusters (BytePlus coders, G/Coder.scala:142-168): ids cross unpacked, one Short each
  val pqEncode16    = fn("gulon_pq_encode16", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS)
  val pqDecode16    = fn("gulon_pq_decode16", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG)
  val indexCreate16 = fn("gulon_index_create16", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
  val setOption     = fn("gulon_set_option", JAVA_INT, ADDRESS, JAVA_LONG)
  val pqQuery       = fn("gulon_pq_query", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT,
                         JAVA_LONG, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)

  def check(rc: Int): Unit = if (rc < 0) {
    val msg = lastError.invoke().asInstanceOf[MemorySegment].reinterpret(1024).getString(0)
    rc match {
      case -1 => throw new IllegalArgumentException(msg)   // GULON_EINVAL  (require(...))
      case -7 => throw new IllegalStateException(msg)      // GULON_ESTATE
      case _  => throw new RuntimeException(s"gulon_b200 error $rc: $msg")
    }
  }
}
