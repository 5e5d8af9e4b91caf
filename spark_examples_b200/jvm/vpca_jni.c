/*
 * vpca_jni.c -- JNI shim between com.google.cloud.genomics.spark.examples.NativePca and libvpca.so.
 * NOT COMPILED HERE (no jni.h in the image); build on a JVM host with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude vpca_jni.c -L. -lvpca -o libvpca_jni.so
 * Primitive arrays are pinned with Get/ReleasePrimitiveArrayCritical for the duration of one call only; libvpca has
 * finished reading them when the call returns (vpca.h "Conventions"), and no JVM reference is retained.
 */
#include <jni.h>
#include <stdint.h>

#include <stddef.h>

#include "vpca.h"

#define CLS(name) Java_com_google_cloud_genomics_spark_examples_NativePca_00024_##name

static void throw_last(JNIEnv* env, vpca_ctx* ctx) {
    jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (ex != NULL) (*env)->ThrowNew(env, ex, vpca_last_error(ctx));
}

JNIEXPORT jlong JNICALL CLS(create)(JNIEnv* env, jobject self, jint n, jint device, jint dtype, jint numPc, jint maxMult,
                                    jint inFlight) {
    vpca_config cfg = {0};
    cfg.struct_size = sizeof(cfg);
    cfg.n_samples = n;
    cfg.device = device;
    cfg.dtype = dtype;
    cfg.num_pc = numPc;
    cfg.max_multiplicity = maxMult;
    cfg.partitions_in_flight = inFlight;
    vpca_ctx* ctx = NULL;
    if (vpca_create(&cfg, &ctx) != VPCA_OK) {
        throw_last(env, NULL);
        return 0;
    }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL CLS(destroy)(JNIEnv* env, jobject self, jlong h) { vpca_destroy((vpca_ctx*)(intptr_t)h); }

JNIEXPORT void JNICALL CLS(reset)(JNIEnv* env, jobject self, jlong h) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    if (vpca_reset(ctx) != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(accumulateCalls)(JNIEnv* env, jobject self, jlong h, jlong pid, jlongArray offsets,
                                            jintArray idx, jlong nv) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    jlong* off = (*env)->GetPrimitiveArrayCritical(env, offsets, NULL);
    jint* ix = (*env)->GetPrimitiveArrayCritical(env, idx, NULL);
    int rc = VPCA_ERR_NOMEM;
    if (off != NULL && ix != NULL) rc = vpca_accumulate_calls(ctx, pid, (const int64_t*)off, (const int32_t*)ix, nv);
    if (ix != NULL) (*env)->ReleasePrimitiveArrayCritical(env, idx, ix, JNI_ABORT);
    if (off != NULL) (*env)->ReleasePrimitiveArrayCritical(env, offsets, off, JNI_ABORT);
    if (rc != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(accumulateCallsU16)(JNIEnv* env, jobject self, jlong h, jlong pid, jlongArray offsets,
                                               jshortArray idx, jlong nv) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    jlong* off = (*env)->GetPrimitiveArrayCritical(env, offsets, NULL);
    jshort* ix = (*env)->GetPrimitiveArrayCritical(env, idx, NULL);
    int rc = VPCA_ERR_NOMEM;
    if (off != NULL && ix != NULL) rc = vpca_accumulate_calls_u16(ctx, pid, (const int64_t*)off, (const uint16_t*)ix, nv);
    if (ix != NULL) (*env)->ReleasePrimitiveArrayCritical(env, idx, ix, JNI_ABORT);
    if (off != NULL) (*env)->ReleasePrimitiveArrayCritical(env, offsets, off, JNI_ABORT);
    if (rc != VPCA_OK) throw_last(env, ctx);
}

/* packed rows: mode 0 = bitmaps (vpca_accumulate_bits), 1 / 2 = PLINK .bed rows counting A1 / A2 */
static void accumulate_packed(JNIEnv* env, jlong h, jlong pid, jbyteArray rows, jlong nv, jlong stride, int mode) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    jbyte* p = (*env)->GetPrimitiveArrayCritical(env, rows, NULL);
    int rc = VPCA_ERR_NOMEM;
    if (p != NULL)
        rc = mode == 0 ? vpca_accumulate_bits(ctx, pid, (const uint8_t*)p, nv, stride)
                       : vpca_accumulate_bed(ctx, pid, (const uint8_t*)p, nv, stride, mode);
    if (p != NULL) (*env)->ReleasePrimitiveArrayCritical(env, rows, p, JNI_ABORT);
    if (rc != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(accumulateBits)(JNIEnv* env, jobject self, jlong h, jlong pid, jbyteArray bits, jlong nv,
                                           jlong stride) {
    accumulate_packed(env, h, pid, bits, nv, stride, 0);
}

JNIEXPORT void JNICALL CLS(accumulateBed)(JNIEnv* env, jobject self, jlong h, jlong pid, jbyteArray rows, jlong nv,
                                          jlong stride, jint counted) {
    if (counted != 1 && counted != 2) counted = 1;
    accumulate_packed(env, h, pid, rows, nv, stride, counted);
}

JNIEXPORT jlong JNICALL CLS(gramDevicePtr)(JNIEnv* env, jobject self, jlong h) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    void* p = NULL;
    if (vpca_gram_device_ptr(ctx, &p) != VPCA_OK) throw_last(env, ctx);
    return (jlong)(intptr_t)p;
}

JNIEXPORT void JNICALL CLS(commit)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    if (vpca_commit(ctx, pid) != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(abort)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    if (vpca_abort(ctx, pid) != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(finalizeGram)(JNIEnv* env, jobject self, jlong h) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    if (vpca_finalize_gram(ctx) != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(getGram)(JNIEnv* env, jobject self, jlong h, jintArray out) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    jint* p = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
    int rc = p ? vpca_get_gram(ctx, (int32_t*)p) : VPCA_ERR_NOMEM;
    if (p) (*env)->ReleasePrimitiveArrayCritical(env, out, p, 0);
    if (rc != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT void JNICALL CLS(setGram)(JNIEnv* env, jobject self, jlong h, jintArray gram) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    jint* p = (*env)->GetPrimitiveArrayCritical(env, gram, NULL);
    int rc = p ? vpca_set_gram(ctx, (const int32_t*)p) : VPCA_ERR_NOMEM;
    if (p) (*env)->ReleasePrimitiveArrayCritical(env, gram, p, JNI_ABORT);
    if (rc != VPCA_OK) throw_last(env, ctx);
}

JNIEXPORT jint JNICALL CLS(computePca)(JNIEnv* env, jobject self, jlong h, jint k, jdoubleArray vecs, jdoubleArray evals) {
    vpca_ctx* ctx = (vpca_ctx*)(intptr_t)h;
    jdouble* v = (*env)->GetPrimitiveArrayCritical(env, vecs, NULL);
    jdouble* e = (*env)->GetPrimitiveArrayCritical(env, evals, NULL);
    int32_t nz = 0;
    int rc = (v && e) ? vpca_compute_pca(ctx, k, v, e, &nz) : VPCA_ERR_NOMEM;
    if (e) (*env)->ReleasePrimitiveArrayCritical(env, evals, e, 0);
    if (v) (*env)->ReleasePrimitiveArrayCritical(env, vecs, v, 0);
    if (rc != VPCA_OK) throw_last(env, ctx);
    return nz;
}
