/*
 * NativePcaPool -- what `class VariantsPcaDriver` holds on a multi-GPU host: ONE vpca_pool (include/vpca.h) = one
 * vpca_ctx per GPU of the box, shared by all task threads of the driver JVM (`local[*]`; the reference's process
 * model, VariantsPca.scala:38-50).  Spark partition p is served by GPU p % nGpus.  The contexts are wired together
 * inside libvpca (vpca_gram_set_peers_local: peer access between the devices, no IPC, no second process), so every
 * commit is added straight into the owners of its Gram rows over NVLink and `reduceAndFinalize` -- the
 * `reduceByKey(_ + _)` of VariantsPca.scala:190 -- is only the closing barrier, the all-gather of the row bands and the
 * symmetrize.
 *
 * accumulate* / commit / abort are safe from many task threads at once; reset / reduceAndFinalize / getGram* /
 * computePca are driver-side.  spark.speculation must stay off (a partition id belongs to one task at a time).
 */
package com.google.cloud.genomics.spark.examples

import java.nio.ByteBuffer

object NativePcaPool {
  System.loadLibrary("vpca_jni")

  @native def create(nSamples: Int, nGpus: Int, dtype: Int, numPc: Int, maxMultiplicity: Int, partitionsInFlight: Int,
                     stagingLanes: Int): Long
  @native def destroy(pool: Long): Unit
  @native def size(pool: Long): Int
  /** The NativePca handle of the GPU that serves `partitionId` (partitionId % size): for the single-context entry points
   *  (NativePca.joinRows / accumulateJoined); the pool keeps owning it. */
  @native def ctx(pool: Long, partitionId: Long): Long
  @native def reset(pool: Long): Unit
  @native def accumulateCalls(pool: Long, partitionId: Long, offsets: Array[Long], sampleIdx: Array[Int], nv: Long): Unit
  @native def accumulateCallsU16(pool: Long, partitionId: Long, offsets: Array[Long], sampleIdx: Array[Short], nv: Long): Unit
  @native def accumulateCallsDirect(pool: Long, partitionId: Long, offsets: ByteBuffer, sampleIdx: ByteBuffer, nv: Long,
                                    idxBytes: Int): Unit
  @native def accumulateBits(pool: Long, partitionId: Long, bits: Array[Byte], nv: Long, strideBytes: Long): Unit
  @native def accumulateBitsDirect(pool: Long, partitionId: Long, bits: ByteBuffer, nv: Long, strideBytes: Long): Unit
  @native def accumulateBed(pool: Long, partitionId: Long, rows: Array[Byte], nv: Long, strideBytes: Long,
                            countedAllele: Int): Unit
  @native def commit(pool: Long, partitionId: Long): Unit
  @native def abort(pool: Long, partitionId: Long): Unit
  @native def reduceAndFinalize(pool: Long): Unit
  @native def getGram(pool: Long, nSamples: Int, out: Array[Int]): Unit
  @native def getGramRows(pool: Long, nSamples: Int, row0: Int, rows: Int, out: Array[Int]): Unit
  @native def computePca(pool: Long, nSamples: Int, k: Int, vecs: Array[Double], evals: Array[Double]): Int

  // ---- driver-side cache: one pool per (cohort size, GPU count) for the life of the JVM ----------------------------
  private val pools = scala.collection.mutable.Map[(Int, Int, Int), Long]()

  /** The pool for this cohort, created on first use; tasks of one analysis share it (VariantsPca.scala:184-189). */
  def get(nSamples: Int, nGpus: Int, numPc: Int, taskThreads: Int = Runtime.getRuntime.availableProcessors): Long =
    pools.synchronized {
      pools.getOrElseUpdate((nSamples, nGpus, numPc), {
        // partitionsInFlight: every task thread may hold one partition open on any GPU; two staging lanes per GPU
        // overlap the H2D copy of one task with the Gram kernel of another
        val h = create(nSamples, nGpus, NativePca.DTYPE_I8, math.max(2, numPc), 2, taskThreads + 2, 2)
        sys.addShutdownHook(destroy(h))
        h
      })
    }

  /** Pack a batch of RDD[Seq[Int]] rows (VariantsPca.scala:153-168) into CSR primitives and stage it. */
  def accumulateRows(pool: Long, partitionId: Long, rows: Seq[Seq[Int]], nSamples: Int): Unit = {
    val offsets = rows.scanLeft(0L)(_ + _.size).toArray
    if (nSamples <= 65536) {
      val idx = new Array[Short](offsets.last.toInt)
      var w = 0
      rows.foreach(_.foreach { s => idx(w) = s.toShort; w += 1 })  // toShort keeps the low 16 bits = uint16 on the wire
      accumulateCallsU16(pool, partitionId, offsets, idx, rows.size)
    } else {
      accumulateCalls(pool, partitionId, offsets, rows.iterator.flatten.toArray, rows.size)
    }
  }
}
