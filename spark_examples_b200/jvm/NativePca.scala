/*
 * NativePca -- JNI binding of libvpca.so (include/vpca.h), 1:1 with the C ABI.
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no JVM, scalac or jni.h here (SURVEY.md 8c).  This file and
 * vpca_jni.c are the binding a maintainer of googlegenomics/spark-examples adds; spark_examples_b200/native.py is the
 * same binding over ctypes and is what the tests exercise.
 */
package com.google.cloud.genomics.spark.examples

object NativePca {
  System.loadLibrary("vpca_jni") // links against libvpca.so

  val DTYPE_I8 = 0
  val DTYPE_BF16 = 1
  val DTYPE_E2M1 = 2 // packed 4-bit cells (tcgen05 kind::mxf4 with unit scales; exact for 0/1/2)

  // every native method throws RuntimeException(vpca_last_error) on a negative vpca_status
  @native def create(nSamples: Int, device: Int, dtype: Int, numPc: Int, maxMultiplicity: Int,
                     partitionsInFlight: Int): Long
  @native def destroy(handle: Long): Unit
  @native def reset(handle: Long): Unit
  /** offsets: nv + 1 entries; sampleIdx: the concatenated rows of one batch of RDD[Seq[Int]] (VariantsPca.scala:153-168). */
  @native def accumulateCalls(handle: Long, partitionId: Long, offsets: Array[Long], sampleIdx: Array[Int], nv: Long): Unit
  /** Same rows with 16-bit sample indices (N <= 65536): half the PCIe bytes. */
  @native def accumulateCallsU16(handle: Long, partitionId: Long, offsets: Array[Long], sampleIdx: Array[Short], nv: Long): Unit
  /** One N-bit bitmap per variant, bit s (LSB first) = hasVariation of sample s; rows strideBytes apart. */
  @native def accumulateBits(handle: Long, partitionId: Long, bits: Array[Byte], nv: Long, strideBytes: Long): Unit
  /** PLINK .bed rows as on disk (2 bits per sample); countedAllele 1 = A1, 2 = A2. */
  @native def accumulateBed(handle: Long, partitionId: Long, rows: Array[Byte], nv: Long, strideBytes: Long,
                            countedAllele: Int): Unit
  @native def commit(handle: Long, partitionId: Long): Unit
  @native def abort(handle: Long, partitionId: Long): Unit
  @native def finalizeGram(handle: Long): Unit
  /** Row-major N x N (the collected RDD[((Int, Int), Int)] of VariantsPca.scala:182-191 in key order). */
  @native def getGram(handle: Long, out: Array[Int]): Unit
  @native def setGram(handle: Long, gram: Array[Int]): Unit
  /** Device address of the int32 Gram, for the NCCL all-reduce of NativePcaPool (INTEGRATION.md section 3). */
  @native def gramDevicePtr(handle: Long): Long
  /** vecs: N x k column-major -- the layout of `pca.toArray` (VariantsPca.scala:227); returns nonZeroRows (:207). */
  @native def computePca(handle: Long, k: Int, vecs: Array[Double], evals: Array[Double]): Int
}
