/*
 * vpca_jni.c -- JNI shim between com.google.cloud.genomics.spark.examples.{NativePca, NativePcaPool} and libvpca.so.
 *
 * Build on a JVM host:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude vpca_jni.c -L. -lvpca -o libvpca_jni.so
 * This image has no JDK: the file is type-checked and EXECUTED against tests/stubs/jni.h + the mock JNIEnv of
 * tests/jni_harness.c (`-m gpu`: real GPU work behind every call; `-m "not gpu"`: argument validation).
 *
 * JVM rules this shim keeps:
 *   - no Get/ReleasePrimitiveArrayCritical: every vpca_* call may lock, enqueue copies and block on a stream, and a
 *     JNI critical region must not block (on HotSpot it holds off every GC for its whole duration).  Java arrays are
 *     copied with Get<Type>ArrayRegion into a native buffer first; tasks that care about the copy pack their rows
 *     straight into PINNED direct ByteBuffers (allocPinned + accumulateCallsDirect): zero copies, no array pinning.
 *   - every count taken from the caller (nv, stride, offsets[nv]) is checked against GetArrayLength /
 *     GetDirectBufferCapacity before a pointer derived from it is handed to the library.
 *   - no JVM reference is retained across calls; errors surface as RuntimeException(vpca_last_error) (an index outside
 *     [0, N) as IndexOutOfBoundsException, what Breeze throws at VariantsPca.scala:188).
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vpca.h"

#define PCA(name) Java_com_google_cloud_genomics_spark_examples_NativePca_00024_##name
#define POOL(name) Java_com_google_cloud_genomics_spark_examples_NativePcaPool_00024_##name

/* one handle type for both classes: a single context (NativePca) or a pool of them (NativePcaPool) */
typedef struct target {
    vpca_ctx* ctx;
    vpca_pool* pool;
} target;

static target from_ctx(jlong h) { target t = {(vpca_ctx*)(intptr_t)h, NULL}; return t; }
static target from_pool(jlong h) { target t = {NULL, (vpca_pool*)(intptr_t)h}; return t; }

static const char* last_error(target t) { return t.pool ? vpca_pool_last_error(t.pool) : vpca_last_error(t.ctx); }

static void throw_msg(JNIEnv* env, const char* cls, const char* msg) {
    jclass ex = (*env)->FindClass(env, cls);
    if (ex != NULL) (*env)->ThrowNew(env, ex, msg);
}

static void throw_rc(JNIEnv* env, target t, int rc) {
    throw_msg(env, rc == VPCA_ERR_INDEX_OUT_OF_RANGE ? "java/lang/IndexOutOfBoundsException"
                   : rc == VPCA_ERR_NOMEM            ? "java/lang/OutOfMemoryError"
                   : rc == VPCA_ERR_BAD_ARG          ? "java/lang/IllegalArgumentException"
                                                     : "java/lang/RuntimeException",
              last_error(t));
}

static void throw_arg(JNIEnv* env, const char* fmt, long long a, long long b) {
    char buf[256];
    snprintf(buf, sizeof(buf), fmt, a, b);
    throw_msg(env, "java/lang/IllegalArgumentException", buf);
}

static void* xmalloc(JNIEnv* env, size_t bytes) {
    void* p = malloc(bytes > 0 ? bytes : 1);
    if (p == NULL) throw_msg(env, "java/lang/OutOfMemoryError", "vpca_jni: native staging buffer");
    return p;
}

/* ------------------------------------------------------------------------------------------- dispatch */
static int t_reset(target t) { return t.pool ? vpca_pool_reset(t.pool) : vpca_reset(t.ctx); }
static int t_commit(target t, jlong pid) { return t.pool ? vpca_pool_commit(t.pool, pid) : vpca_commit(t.ctx, pid); }
static int t_abort(target t, jlong pid) { return t.pool ? vpca_pool_abort(t.pool, pid) : vpca_abort(t.ctx, pid); }
static int t_calls(target t, jlong pid, const int64_t* off, const void* idx, int idx_bytes, jlong nv) {
    if (idx_bytes == 2)
        return t.pool ? vpca_pool_accumulate_calls_u16(t.pool, pid, off, (const uint16_t*)idx, nv)
                      : vpca_accumulate_calls_u16(t.ctx, pid, off, (const uint16_t*)idx, nv);
    return t.pool ? vpca_pool_accumulate_calls(t.pool, pid, off, (const int32_t*)idx, nv)
                  : vpca_accumulate_calls(t.ctx, pid, off, (const int32_t*)idx, nv);
}
static int t_packed(target t, jlong pid, const uint8_t* rows, jlong nv, jlong stride, int mode) {
    if (mode == 0) return t.pool ? vpca_pool_accumulate_bits(t.pool, pid, rows, nv, stride) : vpca_accumulate_bits(t.ctx, pid, rows, nv, stride);
    return t.pool ? vpca_pool_accumulate_bed(t.pool, pid, rows, nv, stride, mode) : vpca_accumulate_bed(t.ctx, pid, rows, nv, stride, mode);
}
static int t_get_gram(target t, int32_t* out) { return t.pool ? vpca_pool_get_gram(t.pool, out) : vpca_get_gram(t.ctx, out); }
static int t_compute_pca(target t, int k, double* vecs, double* evals, int32_t* nz) {
    return t.pool ? vpca_pool_compute_pca(t.pool, k, vecs, evals, nz) : vpca_compute_pca(t.ctx, k, vecs, evals, nz);
}

/* ------------------------------------------------------------------------------------------- shared bodies */
/* offsets: nv + 1 entries; idx: the concatenated rows of one batch of RDD[Seq[Int]] (VariantsPca.scala:153-168) */
static void calls_from_arrays(JNIEnv* env, target t, jlong pid, jlongArray offsets, jarray idx, int idx_bytes, jlong nv) {
    if (offsets == NULL || idx == NULL) { throw_arg(env, "offsets / sampleIdx is null%lld%lld", 0, 0); return; }
    const jlong off_len = (*env)->GetArrayLength(env, offsets), idx_len = (*env)->GetArrayLength(env, idx);
    if (nv < 0 || nv + 1 > off_len) { throw_arg(env, "nv = %lld needs nv + 1 offsets, the array holds %lld", nv, off_len); return; }
    int64_t* off = (int64_t*)xmalloc(env, (size_t)(nv + 1) * sizeof(int64_t));
    if (off == NULL) return;
    (*env)->GetLongArrayRegion(env, offsets, 0, (jsize)(nv + 1), (jlong*)off);
    /* rows are offsets[0] .. offsets[nv] of the index array: everything the library will read must exist */
    if (off[0] < 0 || off[nv] < off[0] || off[nv] > idx_len) {
        throw_arg(env, "offsets[nv] = %lld lies outside the sampleIdx array of %lld entries", off[nv], idx_len);
        free(off);
        return;
    }
    const size_t count = (size_t)(off[nv] - off[0]);
    void* ix = xmalloc(env, count * (size_t)idx_bytes);
    if (ix == NULL) { free(off); return; }
    if (count > 0) {
        if (idx_bytes == 2) (*env)->GetShortArrayRegion(env, idx, (jsize)off[0], (jsize)count, (jshort*)ix);
        else (*env)->GetIntArrayRegion(env, idx, (jsize)off[0], (jsize)count, (jint*)ix);
    }
    const int64_t base = off[0];          /* the native copy starts at the first row: rebase */
    for (jlong v = 0; v <= nv; ++v) off[v] -= base;
    const int rc = t_calls(t, pid, off, ix, idx_bytes, nv);
    free(ix);
    free(off);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
}

/* the same rows in PINNED direct buffers (allocPinned): nothing is copied on the host */
static void calls_from_direct(JNIEnv* env, target t, jlong pid, jobject offsets, jobject idx, jlong nv, jint idx_bytes) {
    if (idx_bytes != 2 && idx_bytes != 4) { throw_arg(env, "idxBytes must be 2 or 4, not %lld%lld", idx_bytes, 0); return; }
    const int64_t* off = offsets ? (const int64_t*)(*env)->GetDirectBufferAddress(env, offsets) : NULL;
    const char* ix = idx ? (const char*)(*env)->GetDirectBufferAddress(env, idx) : NULL;
    if (off == NULL || ix == NULL) { throw_arg(env, "offsets / sampleIdx must be direct ByteBuffers%lld%lld", 0, 0); return; }
    const jlong off_cap = (*env)->GetDirectBufferCapacity(env, offsets), idx_cap = (*env)->GetDirectBufferCapacity(env, idx);
    if (nv < 0 || (nv + 1) * 8 > off_cap) { throw_arg(env, "nv = %lld needs (nv + 1) * 8 bytes of offsets, the buffer holds %lld", nv, off_cap); return; }
    if (off[0] < 0 || off[nv] < off[0] || off[nv] * idx_bytes > idx_cap) {
        throw_arg(env, "offsets[nv] = %lld entries do not fit the sampleIdx buffer of %lld bytes", off[nv], idx_cap);
        return;
    }
    const int rc = t_calls(t, pid, off, ix, idx_bytes, nv);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
}

/* packed rows: mode 0 = bitmaps (vpca_accumulate_bits), 1 / 2 = PLINK .bed rows counting A1 / A2 */
static void packed_from_array(JNIEnv* env, target t, jlong pid, jbyteArray rows, jlong nv, jlong stride, int mode) {
    if (rows == NULL) { throw_arg(env, "rows is null%lld%lld", 0, 0); return; }
    const jlong len = (*env)->GetArrayLength(env, rows);
    if (nv < 0 || stride <= 0 || nv > len / stride) { throw_arg(env, "nv x strideBytes = %lld bytes, the array holds %lld", nv * stride, len); return; }
    uint8_t* p = (uint8_t*)xmalloc(env, (size_t)(nv * stride));
    if (p == NULL) return;
    if (nv > 0) (*env)->GetByteArrayRegion(env, rows, 0, (jsize)(nv * stride), (jbyte*)p);
    const int rc = t_packed(t, pid, p, nv, stride, mode);
    free(p);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
}

static void packed_from_direct(JNIEnv* env, target t, jlong pid, jobject rows, jlong nv, jlong stride, int mode) {
    const uint8_t* p = rows ? (const uint8_t*)(*env)->GetDirectBufferAddress(env, rows) : NULL;
    if (p == NULL) { throw_arg(env, "rows must be a direct ByteBuffer%lld%lld", 0, 0); return; }
    const jlong cap = (*env)->GetDirectBufferCapacity(env, rows);
    if (nv < 0 || stride <= 0 || nv > cap / stride) { throw_arg(env, "nv x strideBytes = %lld bytes, the buffer holds %lld", nv * stride, cap); return; }
    const int rc = t_packed(t, pid, p, nv, stride, mode);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
}

static void get_gram(JNIEnv* env, target t, jint n, jintArray out) {
    if (out == NULL || (jlong)(*env)->GetArrayLength(env, out) < (jlong)n * n) { throw_arg(env, "out must hold n * n = %lld ints%lld", (jlong)n * n, 0); return; }
    int32_t* p = (int32_t*)xmalloc(env, (size_t)n * n * sizeof(int32_t));
    if (p == NULL) return;
    const int rc = t_get_gram(t, p);
    if (rc == VPCA_OK) (*env)->SetIntArrayRegion(env, out, 0, (jsize)((jlong)n * n), (const jint*)p);
    free(p);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
}

static jint compute_pca(JNIEnv* env, target t, jint n, jint k, jdoubleArray vecs, jdoubleArray evals) {
    if (vecs == NULL || k < 1 || (jlong)(*env)->GetArrayLength(env, vecs) < (jlong)n * k ||
        (evals != NULL && (*env)->GetArrayLength(env, evals) < k)) {
        throw_arg(env, "vecs must hold n * k = %lld doubles and evals k = %lld", (jlong)n * k, k);
        return 0;
    }
    double* v = (double*)xmalloc(env, ((size_t)n * k + (size_t)k) * sizeof(double));
    if (v == NULL) return 0;
    int32_t nz = 0;
    const int rc = t_compute_pca(t, k, v, v + (size_t)n * k, &nz);
    if (rc == VPCA_OK) {
        (*env)->SetDoubleArrayRegion(env, vecs, 0, (jsize)((jlong)n * k), v);          /* layout of pca.toArray (:227) */
        if (evals != NULL) (*env)->SetDoubleArrayRegion(env, evals, 0, k, v + (size_t)n * k);
    }
    free(v);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
    return nz;
}

static void fill_config(vpca_config* cfg, jint n, jint device, jint dtype, jint numPc, jint maxMult, jint inFlight, jint lanes) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->n_samples = n;
    cfg->device = device;
    cfg->dtype = dtype;
    cfg->num_pc = numPc;
    cfg->max_multiplicity = maxMult;
    cfg->partitions_in_flight = inFlight;
    cfg->staging_lanes = lanes;
}

/* =========================================================================================== NativePca (one GPU) */
JNIEXPORT jlong JNICALL PCA(create)(JNIEnv* env, jobject self, jint n, jint device, jint dtype, jint numPc, jint maxMult,
                                    jint inFlight) {
    (void)self;
    vpca_config cfg;
    fill_config(&cfg, n, device, dtype, numPc, maxMult, inFlight, 0);
    vpca_ctx* ctx = NULL;
    const int rc = vpca_create(&cfg, &ctx);
    if (rc != VPCA_OK) {
        throw_rc(env, from_ctx(0), rc);
        return 0;
    }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL PCA(destroy)(JNIEnv* env, jobject self, jlong h) { (void)env; (void)self; vpca_destroy((vpca_ctx*)(intptr_t)h); }

JNIEXPORT void JNICALL PCA(reset)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    const int rc = t_reset(from_ctx(h));
    if (rc != VPCA_OK) throw_rc(env, from_ctx(h), rc);
}

JNIEXPORT void JNICALL PCA(accumulateCalls)(JNIEnv* env, jobject self, jlong h, jlong pid, jlongArray offsets, jintArray idx, jlong nv) {
    (void)self;
    calls_from_arrays(env, from_ctx(h), pid, offsets, idx, 4, nv);
}

JNIEXPORT void JNICALL PCA(accumulateCallsU16)(JNIEnv* env, jobject self, jlong h, jlong pid, jlongArray offsets, jshortArray idx, jlong nv) {
    (void)self;
    calls_from_arrays(env, from_ctx(h), pid, offsets, idx, 2, nv);
}

JNIEXPORT void JNICALL PCA(accumulateCallsDirect)(JNIEnv* env, jobject self, jlong h, jlong pid, jobject offsets, jobject idx, jlong nv, jint idxBytes) {
    (void)self;
    calls_from_direct(env, from_ctx(h), pid, offsets, idx, nv, idxBytes);
}

JNIEXPORT void JNICALL PCA(accumulateBits)(JNIEnv* env, jobject self, jlong h, jlong pid, jbyteArray bits, jlong nv, jlong stride) {
    (void)self;
    packed_from_array(env, from_ctx(h), pid, bits, nv, stride, 0);
}

JNIEXPORT void JNICALL PCA(accumulateBed)(JNIEnv* env, jobject self, jlong h, jlong pid, jbyteArray rows, jlong nv, jlong stride, jint counted) {
    (void)self;
    if (counted != 1 && counted != 2) { throw_arg(env, "countedAllele must be 1 (A1) or 2 (A2), not %lld%lld", counted, 0); return; }
    packed_from_array(env, from_ctx(h), pid, rows, nv, stride, counted);
}

JNIEXPORT void JNICALL PCA(commit)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    (void)self;
    const int rc = t_commit(from_ctx(h), pid);
    if (rc != VPCA_OK) throw_rc(env, from_ctx(h), rc);
}

JNIEXPORT void JNICALL PCA(abort)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    (void)self;
    const int rc = t_abort(from_ctx(h), pid);
    if (rc != VPCA_OK) throw_rc(env, from_ctx(h), rc);
}

JNIEXPORT void JNICALL PCA(finalizeGram)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    const int rc = vpca_finalize_gram((vpca_ctx*)(intptr_t)h);
    if (rc != VPCA_OK) throw_rc(env, from_ctx(h), rc);
}

JNIEXPORT void JNICALL PCA(getGram)(JNIEnv* env, jobject self, jlong h, jint n, jintArray out) { (void)self; get_gram(env, from_ctx(h), n, out); }

JNIEXPORT void JNICALL PCA(setGram)(JNIEnv* env, jobject self, jlong h, jint n, jintArray gram) {
    (void)self;
    if (gram == NULL || (jlong)(*env)->GetArrayLength(env, gram) < (jlong)n * n) { throw_arg(env, "gram must hold n * n = %lld ints%lld", (jlong)n * n, 0); return; }
    int32_t* p = (int32_t*)xmalloc(env, (size_t)n * n * sizeof(int32_t));
    if (p == NULL) return;
    (*env)->GetIntArrayRegion(env, gram, 0, (jsize)((jlong)n * n), (jint*)p);
    const int rc = vpca_set_gram((vpca_ctx*)(intptr_t)h, p);
    free(p);
    if (rc != VPCA_OK) throw_rc(env, from_ctx(h), rc);
}

JNIEXPORT jlong JNICALL PCA(gramDevicePtr)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    void* p = NULL;
    const int rc = vpca_gram_device_ptr((vpca_ctx*)(intptr_t)h, &p);
    if (rc != VPCA_OK) throw_rc(env, from_ctx(h), rc);
    return (jlong)(intptr_t)p;
}

JNIEXPORT jint JNICALL PCA(computePca)(JNIEnv* env, jobject self, jlong h, jint n, jint k, jdoubleArray vecs, jdoubleArray evals) {
    (void)self;
    return compute_pca(env, from_ctx(h), n, k, vecs, evals);
}

/* pinned host memory as direct ByteBuffers (shared by both classes) */
/* joinDatasets / mergeDatasets (VariantsPca.scala:115-148) on the device: key bytes + rows of all datasets in, the joined
 * rows stay in the context (accumulateJoined).  Returns the number of joined rows. */
JNIEXPORT jlong JNICALL PCA(joinRows)(JNIEnv* env, jobject self, jlong h, jint mode, jint variantSetCount, jlong nLeft,
                                      jbyteArray keyBytes, jlongArray keyOffsets, jlongArray offsets, jintArray idx, jlong nrows) {
    (void)self;
    target t = from_ctx(h);
    if (keyBytes == NULL || keyOffsets == NULL || offsets == NULL || idx == NULL) { throw_arg(env, "joinRows: null array%lld%lld", 0, 0); return 0; }
    const jlong kb_len = (*env)->GetArrayLength(env, keyBytes), ko_len = (*env)->GetArrayLength(env, keyOffsets);
    const jlong off_len = (*env)->GetArrayLength(env, offsets), idx_len = (*env)->GetArrayLength(env, idx);
    if (nrows < 0 || nrows + 1 > ko_len || nrows + 1 > off_len) {
        throw_arg(env, "joinRows: nrows = %lld needs nrows + 1 key offsets and row offsets (%lld present)", nrows, ko_len < off_len ? ko_len : off_len);
        return 0;
    }
    int64_t* ko = (int64_t*)xmalloc(env, (size_t)(nrows + 1) * 8);
    int64_t* off = ko ? (int64_t*)xmalloc(env, (size_t)(nrows + 1) * 8) : NULL;
    if (ko == NULL || off == NULL) { free(ko); return 0; }
    (*env)->GetLongArrayRegion(env, keyOffsets, 0, (jsize)(nrows + 1), (jlong*)ko);
    (*env)->GetLongArrayRegion(env, offsets, 0, (jsize)(nrows + 1), (jlong*)off);
    jlong rows = 0;
    if (ko[0] != 0 || off[0] != 0 || ko[nrows] < 0 || ko[nrows] > kb_len || off[nrows] < 0 || off[nrows] > idx_len) {
        throw_arg(env, "joinRows: offsets must start at 0 and end inside the arrays (%lld key bytes, %lld calls)", kb_len, idx_len);
    } else {
        uint8_t* kb = (uint8_t*)xmalloc(env, (size_t)ko[nrows]);
        int32_t* ix = kb ? (int32_t*)xmalloc(env, (size_t)off[nrows] * 4) : NULL;
        if (kb != NULL && ix != NULL) {
            if (ko[nrows] > 0) (*env)->GetByteArrayRegion(env, keyBytes, 0, (jsize)ko[nrows], (jbyte*)kb);
            if (off[nrows] > 0) (*env)->GetIntArrayRegion(env, idx, 0, (jsize)off[nrows], (jint*)ix);
            int64_t out_rows = 0, out_nnz = 0;
            const int rc = vpca_join_rows(t.ctx, mode, variantSetCount, nLeft, kb, ko, off, ix, nrows, &out_rows, &out_nnz);
            if (rc != VPCA_OK) throw_rc(env, t, rc);
            rows = out_rows;
        }
        free(ix);
        free(kb);
    }
    free(off);
    free(ko);
    return rows;
}

/* the joined rows of the last joinRows as Java arrays (JoinedCallsRDD.compute): offsets needs rows + 1 entries, idx
 * enough room for the calls (its length bounds what is written) */
JNIEXPORT void JNICALL PCA(joinFetch)(JNIEnv* env, jobject self, jlong h, jlongArray outOffsets, jintArray outIdx) {
    (void)self;
    target t = from_ctx(h);
    if (outOffsets == NULL || outIdx == NULL) { throw_arg(env, "joinFetch: null array%lld%lld", 0, 0); return; }
    const jlong no = (*env)->GetArrayLength(env, outOffsets), ni = (*env)->GetArrayLength(env, outIdx);
    if (no < 1) { throw_arg(env, "joinFetch: offsets must hold rows + 1 entries%lld%lld", 0, 0); return; }
    int64_t* off = (int64_t*)xmalloc(env, (size_t)no * 8);
    int32_t* ix = off ? (int32_t*)xmalloc(env, (size_t)(ni > 0 ? ni : 1) * 4) : NULL;
    if (off == NULL || ix == NULL) { free(off); return; }
    /* sizes of the retained result: a join of zero input rows reports them without touching anything */
    int64_t rows = 0, nnz = 0;
    int rc = vpca_join_size(t.ctx, &rows, &nnz);
    if (rc == VPCA_OK && (rows + 1 > no || nnz > ni)) {
        throw_arg(env, "joinFetch: the joined CSR has %lld rows and %lld calls; the arrays are smaller", rows, nnz);
    } else {
        if (rc == VPCA_OK) rc = vpca_join_fetch(t.ctx, off, ix);
        if (rc != VPCA_OK) throw_rc(env, t, rc);
        else {
            (*env)->SetLongArrayRegion(env, outOffsets, 0, (jsize)(rows + 1), (const jlong*)off);
            if (nnz > 0) (*env)->SetIntArrayRegion(env, outIdx, 0, (jsize)nnz, (const jint*)ix);
        }
    }
    free(ix);
    free(off);
}

JNIEXPORT jlong JNICALL PCA(joinRowCount)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    target t = from_ctx(h);
    int64_t rows = 0, nnz = 0;
    const int rc = vpca_join_size(t.ctx, &rows, &nnz);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
    return rows;
}

JNIEXPORT jlong JNICALL PCA(joinCallCount)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    target t = from_ctx(h);
    int64_t rows = 0, nnz = 0;
    const int rc = vpca_join_size(t.ctx, &rows, &nnz);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
    return nnz;
}

JNIEXPORT void JNICALL PCA(accumulateJoined)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    (void)self;
    target t = from_ctx(h);
    const int rc = vpca_accumulate_joined(t.ctx, pid);
    if (rc != VPCA_OK) throw_rc(env, t, rc);
}

JNIEXPORT jobject JNICALL PCA(allocPinned)(JNIEnv* env, jobject self, jlong bytes) {
    (void)self;
    void* p = NULL;
    if (bytes <= 0) { throw_arg(env, "allocPinned(%lld)%lld", bytes, 0); return NULL; }
    const int rc = vpca_host_alloc((size_t)bytes, &p);
    if (rc != VPCA_OK) {
        throw_rc(env, from_ctx(0), rc);
        return NULL;
    }
    jobject buf = (*env)->NewDirectByteBuffer(env, p, bytes);
    if (buf == NULL) vpca_host_free(p);
    return buf;
}

JNIEXPORT void JNICALL PCA(freePinned)(JNIEnv* env, jobject self, jobject buf) {
    (void)self;
    void* p = buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL;
    if (p != NULL) vpca_host_free(p);
}

/* =========================================================================================== NativePcaPool (all GPUs) */
JNIEXPORT jlong JNICALL POOL(create)(JNIEnv* env, jobject self, jint n, jint nGpus, jint dtype, jint numPc, jint maxMult,
                                     jint inFlight, jint lanes) {
    (void)self;
    vpca_config cfg;
    fill_config(&cfg, n, 0, dtype, numPc, maxMult, inFlight, lanes);
    vpca_pool* pool = NULL;
    const int rc = vpca_pool_create(&cfg, nGpus, NULL, &pool);
    if (rc != VPCA_OK) {
        target t = {NULL, NULL};
        throw_msg(env, rc == VPCA_ERR_BAD_ARG ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException", vpca_pool_last_error(NULL));
        (void)t;
        return 0;
    }
    return (jlong)(intptr_t)pool;
}

JNIEXPORT void JNICALL POOL(destroy)(JNIEnv* env, jobject self, jlong h) { (void)env; (void)self; vpca_pool_destroy((vpca_pool*)(intptr_t)h); }

JNIEXPORT jint JNICALL POOL(size)(JNIEnv* env, jobject self, jlong h) { (void)env; (void)self; return vpca_pool_size((vpca_pool*)(intptr_t)h); }

JNIEXPORT jlong JNICALL POOL(ctx)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    (void)env; (void)self;
    return (jlong)(intptr_t)vpca_pool_ctx((vpca_pool*)(intptr_t)h, pid);
}

JNIEXPORT void JNICALL POOL(reset)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    const int rc = t_reset(from_pool(h));
    if (rc != VPCA_OK) throw_rc(env, from_pool(h), rc);
}

JNIEXPORT void JNICALL POOL(accumulateCalls)(JNIEnv* env, jobject self, jlong h, jlong pid, jlongArray offsets, jintArray idx, jlong nv) {
    (void)self;
    calls_from_arrays(env, from_pool(h), pid, offsets, idx, 4, nv);
}

JNIEXPORT void JNICALL POOL(accumulateCallsU16)(JNIEnv* env, jobject self, jlong h, jlong pid, jlongArray offsets, jshortArray idx, jlong nv) {
    (void)self;
    calls_from_arrays(env, from_pool(h), pid, offsets, idx, 2, nv);
}

JNIEXPORT void JNICALL POOL(accumulateCallsDirect)(JNIEnv* env, jobject self, jlong h, jlong pid, jobject offsets, jobject idx, jlong nv, jint idxBytes) {
    (void)self;
    calls_from_direct(env, from_pool(h), pid, offsets, idx, nv, idxBytes);
}

JNIEXPORT void JNICALL POOL(accumulateBits)(JNIEnv* env, jobject self, jlong h, jlong pid, jbyteArray bits, jlong nv, jlong stride) {
    (void)self;
    packed_from_array(env, from_pool(h), pid, bits, nv, stride, 0);
}

JNIEXPORT void JNICALL POOL(accumulateBitsDirect)(JNIEnv* env, jobject self, jlong h, jlong pid, jobject bits, jlong nv, jlong stride) {
    (void)self;
    packed_from_direct(env, from_pool(h), pid, bits, nv, stride, 0);
}

JNIEXPORT void JNICALL POOL(accumulateBed)(JNIEnv* env, jobject self, jlong h, jlong pid, jbyteArray rows, jlong nv, jlong stride, jint counted) {
    (void)self;
    if (counted != 1 && counted != 2) { throw_arg(env, "countedAllele must be 1 (A1) or 2 (A2), not %lld%lld", counted, 0); return; }
    packed_from_array(env, from_pool(h), pid, rows, nv, stride, counted);
}

JNIEXPORT void JNICALL POOL(commit)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    (void)self;
    const int rc = t_commit(from_pool(h), pid);
    if (rc != VPCA_OK) throw_rc(env, from_pool(h), rc);
}

JNIEXPORT void JNICALL POOL(abort)(JNIEnv* env, jobject self, jlong h, jlong pid) {
    (void)self;
    const int rc = t_abort(from_pool(h), pid);
    if (rc != VPCA_OK) throw_rc(env, from_pool(h), rc);
}

/* reduceByKey(_ + _) (VariantsPca.scala:190) over the GPUs of the box + symmetrize */
JNIEXPORT void JNICALL POOL(reduceAndFinalize)(JNIEnv* env, jobject self, jlong h) {
    (void)self;
    const int rc = vpca_pool_reduce_and_finalize((vpca_pool*)(intptr_t)h);
    if (rc != VPCA_OK) throw_rc(env, from_pool(h), rc);
}

JNIEXPORT void JNICALL POOL(getGram)(JNIEnv* env, jobject self, jlong h, jint n, jintArray out) { (void)self; get_gram(env, from_pool(h), n, out); }

/* rows [row0, row0 + rows) of the reduced matrix: what one partition of GramRDD.compute iterates */
JNIEXPORT void JNICALL POOL(getGramRows)(JNIEnv* env, jobject self, jlong h, jint n, jint row0, jint rows, jintArray out) {
    (void)self;
    vpca_pool* pool = (vpca_pool*)(intptr_t)h;
    if (out == NULL || rows < 0 || (jlong)(*env)->GetArrayLength(env, out) < (jlong)rows * n) { throw_arg(env, "out must hold rows * n = %lld ints%lld", (jlong)rows * n, 0); return; }
    int32_t* p = (int32_t*)xmalloc(env, (size_t)rows * n * sizeof(int32_t));
    if (p == NULL) return;
    const int rc = vpca_get_gram_band(vpca_pool_ctx(pool, 0), row0, rows, p);
    if (rc == VPCA_OK) (*env)->SetIntArrayRegion(env, out, 0, (jsize)((jlong)rows * n), (const jint*)p);
    free(p);
    if (rc != VPCA_OK) {
        target t = {vpca_pool_ctx(pool, 0), NULL};
        throw_rc(env, t, rc);
    }
}

JNIEXPORT jint JNICALL POOL(computePca)(JNIEnv* env, jobject self, jlong h, jint n, jint k, jdoubleArray vecs, jdoubleArray evals) {
    (void)self;
    return compute_pca(env, from_pool(h), n, k, vecs, evals);
}
