/*
 * GramRDD -- the `RDD[((Int, Int), Int)]` that `getSimilarityMatrix` returns (VariantsPca.scala:182), backed by the
 * similarity matrix that is resident on the GPUs.  It IS an RDD of ((row, col), count) records with all N^2 keys
 * present, exactly like the reference's `reduceByKey` output (:189-190) -- any caller may map / collect / save it --
 * and it additionally carries the pool handle, so that `computePca` (:198) can run on the device-resident matrix
 * without the N^2 records ever crossing PCIe.  This is what keeps both public signatures of the reference unchanged.
 */
package com.google.cloud.genomics.spark.examples

import org.apache.spark.{Partition, SparkContext, TaskContext}
import org.apache.spark.rdd.RDD

private[examples] case class GramBand(index: Int, row0: Int, rows: Int) extends Partition

class GramRDD(sc: SparkContext, val pool: Long, val nSamples: Int, bands: Int = 8)
    extends RDD[((Int, Int), Int)](sc, Nil) {

  override protected def getPartitions: Array[Partition] = {
    val per = (nSamples + bands - 1) / bands
    (0 until bands).map(b => GramBand(b, math.min(nSamples, b * per), math.max(0, math.min(per, nSamples - b * per))))
      .filter(_.rows > 0).zipWithIndex.map { case (g, i) => g.copy(index = i): Partition }.toArray
  }

  // local[*] only (SURVEY 8b "process model"): the pool handle is a pointer into this JVM
  override def compute(split: Partition, context: TaskContext): Iterator[((Int, Int), Int)] = {
    val band = split.asInstanceOf[GramBand]
    val cells = new Array[Int](band.rows * nSamples)
    NativePcaPool.getGramRows(pool, nSamples, band.row0, band.rows, cells)
    for (r <- Iterator.range(0, band.rows); c <- Iterator.range(0, nSamples))
      yield ((band.row0 + r, c), cells(r * nSamples + c))
  }
}
