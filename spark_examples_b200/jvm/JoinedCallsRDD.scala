/*
 * JoinedCallsRDD -- the `RDD[Seq[Int]]` that `getCallsRdd` returns (VariantsPca.scala:153) on its 2-dataset (join, :159)
 * and N-dataset (merge, :160) branches when keying, join / merge and the flattening of the calls ran on the GPU
 * (NativePca.joinRows -> csrc/join.cu).  It IS an RDD of call rows -- any caller may map / collect / count it: the rows
 * are fetched from the device on first use (NativePca.joinFetch) -- and it additionally carries the context handle, so
 * that `getSimilarityMatrix` (:182) can feed the joined rows to the encoder without them ever crossing PCIe
 * (NativePca.accumulateJoined).  Same device as GramRDD: the public signature of the reference stays as it is.
 */
package com.google.cloud.genomics.spark.examples

import org.apache.spark.{Partition, SparkContext, TaskContext}
import org.apache.spark.rdd.RDD

private[examples] case class JoinedRows(index: Int) extends Partition

class JoinedCallsRDD(sc: SparkContext, val handle: Long, val rows: Long, val calls: Long)
    extends RDD[Seq[Int]](sc, Nil) {

  override protected def getPartitions: Array[Partition] = Array(JoinedRows(0))

  // local[*] only (SURVEY 8b "process model"): the handle is a pointer into this JVM
  override def compute(split: Partition, context: TaskContext): Iterator[Seq[Int]] = {
    val offsets = new Array[Long]((rows + 1).toInt)
    val idx = new Array[Int](math.max(1L, calls).toInt)
    NativePca.joinFetch(handle, offsets, idx)
    Iterator.range(0, rows.toInt)
      .map(r => idx.slice(offsets(r).toInt, offsets(r + 1).toInt).toSeq)
      .filter(_.nonEmpty)                                  // VariantsPca.scala:166
  }
}

object JoinedCallsRDD {
  /** What `joinDatasets` / `mergeDatasets` shuffle, laid out for NativePca.joinRows: per variant (datasets in order) the
   *  bytes getVariantKey would hash (VariantsPca.scala:65-73) and the callset indices with variation (:56-60, :164). */
  def build(sc: SparkContext, handle: Long, datasets: List[RDD[Variant]], indexes: Map[String, Int],
            variantSetCount: Int): JoinedCallsRDD = {
    val join = variantSetCount == 2
    val perDataset = (if (join) datasets.take(2) else datasets).map(_.map { v =>
      val key = new java.io.ByteArrayOutputStream()
      def putLong(x: Long): Unit = (0 until 8).foreach(b => key.write(((x >>> (8 * b)) & 0xff).toInt))   // Guava: little-endian
      key.write(v.contig.getBytes("UTF-8")); putLong(v.start); putLong(v.end)
      key.write(v.referenceBases.getBytes("UTF-8"))
      key.write(v.alternateBases.map(_.mkString("")).getOrElse("").getBytes("UTF-8"))
      val carriers = VariantsPcaDriver.extractCallInfo(v, indexes).filter(_.hasVariation).map(_.callsetId)
      (key.toByteArray, carriers.toArray)
    }.collect())
    val all = perDataset.flatten
    val keyOffsets = all.scanLeft(0L)(_ + _._1.length).toArray
    val offsets = all.scanLeft(0L)(_ + _._2.length).toArray
    val rows = NativePca.joinRows(handle, if (join) 0 else 1, variantSetCount, perDataset.head.length.toLong,
      all.flatMap(_._1).toArray, keyOffsets, offsets, all.flatMap(_._2).toArray, all.length.toLong)
    new JoinedCallsRDD(sc, handle, rows, NativePca.joinCallCount(handle))
  }
}
