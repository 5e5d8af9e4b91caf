/*
 * jni_harness.c -- compiles spark_examples_b200/jvm/vpca_jni.c against tests/stubs/jni.h and RUNS its entry points
 * against a mock JNIEnv (Java arrays and direct buffers are small C structs), so the JNI half of the drop-in boundary is
 * exercised in an image without a JDK.
 *
 *   jni_harness validate            no GPU needed: every length / null check must throw before the library is touched
 *   jni_harness gpu <n> <nv> <gpus> NativePcaPool end to end (int / uint16 arrays, pinned direct buffers, bitmaps,
 *                                   abort + retry) -> Gram bit-exact vs the oracle, computePca sane
 * Test infrastructure (links the oracle as the checker).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../spark_examples_b200/jvm/vpca_jni.c"

/* oracle/vpca_oracle.c */
int64_t vo_synth_calls(uint64_t seed, int32_t n, int64_t v0, int64_t nv, int64_t* off, int32_t* idx, int64_t* nv_out);
int vo_similarity(int32_t n, int64_t nv, const int64_t* off, const int32_t* idx, int32_t n_partitions, int32_t* S);

/* ------------------------------------------------------------------------------------------------- mock JVM */
enum { K_BYTE = 1, K_SHORT = 2, K_INT = 4, K_LONG = 8, K_DOUBLE = 9, K_DIRECT = 100, K_CLASS = 200 };
struct _jobject {
    int kind;
    jlong len; /* elements (arrays) or bytes (direct buffers) */
    void* data;
    const char* name; /* K_CLASS */
};

static char pending_class[128];
static char pending_msg[512];
static int pending = 0;

static jclass m_FindClass(JNIEnv* env, const char* name) {
    (void)env;
    jobject c = (jobject)calloc(1, sizeof(struct _jobject));
    c->kind = K_CLASS;
    c->name = name;
    return c;
}
static jint m_ThrowNew(JNIEnv* env, jclass clazz, const char* msg) {
    (void)env;
    pending = 1;
    snprintf(pending_class, sizeof(pending_class), "%s", clazz->name);
    snprintf(pending_msg, sizeof(pending_msg), "%s", msg ? msg : "");
    return 0;
}
static jboolean m_ExceptionCheck(JNIEnv* env) { (void)env; return (jboolean)pending; }
static jsize m_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return (jsize)a->len; }
#define REGION_GET(NAME, T, KIND)                                                              \
    static void NAME(JNIEnv* env, jarray a, jsize start, jsize len, T* buf) {                  \
        (void)env;                                                                             \
        if (a->kind != KIND || start < 0 || len < 0 || (jlong)start + len > a->len) {          \
            fprintf(stderr, "mock JVM: ArrayIndexOutOfBoundsException in " #NAME "\n");        \
            abort();                                                                           \
        }                                                                                      \
        memcpy(buf, (const T*)a->data + start, (size_t)len * sizeof(T));                       \
    }
REGION_GET(m_GetByteArrayRegion, jbyte, K_BYTE)
REGION_GET(m_GetShortArrayRegion, jshort, K_SHORT)
REGION_GET(m_GetIntArrayRegion, jint, K_INT)
REGION_GET(m_GetLongArrayRegion, jlong, K_LONG)
#define REGION_SET(NAME, T, KIND)                                                              \
    static void NAME(JNIEnv* env, jarray a, jsize start, jsize len, const T* buf) {            \
        (void)env;                                                                             \
        if (a->kind != KIND || start < 0 || len < 0 || (jlong)start + len > a->len) {          \
            fprintf(stderr, "mock JVM: ArrayIndexOutOfBoundsException in " #NAME "\n");        \
            abort();                                                                           \
        }                                                                                      \
        memcpy((T*)a->data + start, buf, (size_t)len * sizeof(T));                             \
    }
REGION_SET(m_SetIntArrayRegion, jint, K_INT)
REGION_SET(m_SetLongArrayRegion, jlong, K_LONG)
REGION_SET(m_SetDoubleArrayRegion, jdouble, K_DOUBLE)
static jobject m_NewDirectByteBuffer(JNIEnv* env, void* address, jlong capacity) {
    (void)env;
    jobject b = (jobject)calloc(1, sizeof(struct _jobject));
    b->kind = K_DIRECT;
    b->len = capacity;
    b->data = address;
    return b;
}
static void* m_GetDirectBufferAddress(JNIEnv* env, jobject b) { (void)env; return b->kind == K_DIRECT ? b->data : NULL; }
static jlong m_GetDirectBufferCapacity(JNIEnv* env, jobject b) { (void)env; return b->kind == K_DIRECT ? b->len : -1; }

static const struct JNINativeInterface_ mock_table = {
    m_FindClass,          m_ThrowNew,          m_ExceptionCheck,      m_GetArrayLength,        m_GetByteArrayRegion,
    m_GetShortArrayRegion, m_GetIntArrayRegion, m_GetLongArrayRegion,  m_SetIntArrayRegion,     m_SetLongArrayRegion,    m_SetDoubleArrayRegion,
    m_NewDirectByteBuffer, m_GetDirectBufferAddress, m_GetDirectBufferCapacity};
static JNIEnv mock_env = &mock_table;

static jarray new_array(int kind, jlong len) {
    const size_t esz = kind == K_BYTE ? 1 : kind == K_SHORT ? 2 : kind == K_INT ? 4 : 8;
    jarray a = (jarray)calloc(1, sizeof(struct _jobject));
    a->kind = kind;
    a->len = len;
    a->data = calloc((size_t)(len > 0 ? len : 1), esz);
    return a;
}
static void free_array(jarray a) {
    free(a->data);
    free(a);
}

static int expect_throw(const char* what, const char* cls_part) {
    if (!pending || strstr(pending_class, cls_part) == NULL) {
        fprintf(stderr, "FAIL %s: expected %s, got %s '%s'\n", what, cls_part, pending ? pending_class : "no exception", pending_msg);
        return 1;
    }
    pending = 0;
    return 0;
}
static int expect_clean(const char* what) {
    if (pending) {
        fprintf(stderr, "FAIL %s: %s: %s\n", what, pending_class, pending_msg);
        pending = 0;
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------- validate */
static int run_validate(void) {
    JNIEnv* env = &mock_env;
    int bad = 0;
    jarray off = new_array(K_LONG, 4), idx = new_array(K_INT, 10), idx16 = new_array(K_SHORT, 10), rows = new_array(K_BYTE, 64);
    jlong* o = (jlong*)off->data;
    o[0] = 0; o[1] = 3; o[2] = 6; o[3] = 9;
    /* handle 0 is never dereferenced: every call below must be rejected by the shim's own checks */
    POOL(accumulateCalls)(env, NULL, 0, 1, off, idx, 4);              /* nv + 1 = 5 offsets needed, 4 present */
    bad += expect_throw("offsets shorter than nv + 1", "IllegalArgument");
    o[3] = 11;                                                        /* rows run past the 10 index entries */
    POOL(accumulateCalls)(env, NULL, 0, 1, off, idx, 3);
    bad += expect_throw("offsets[nv] beyond sampleIdx", "IllegalArgument");
    POOL(accumulateCallsU16)(env, NULL, 0, 1, off, idx16, 3);
    bad += expect_throw("offsets[nv] beyond sampleIdx (u16)", "IllegalArgument");
    o[3] = 9;
    POOL(accumulateCalls)(env, NULL, 0, 1, NULL, idx, 3);
    bad += expect_throw("null offsets", "IllegalArgument");
    POOL(accumulateCalls)(env, NULL, 0, 1, off, idx, -1);
    bad += expect_throw("negative nv", "IllegalArgument");
    POOL(accumulateBits)(env, NULL, 0, 1, rows, 9, 8);                /* 72 bytes needed, 64 present */
    bad += expect_throw("bitmap rows beyond the array", "IllegalArgument");
    POOL(accumulateBed)(env, NULL, 0, 1, rows, 2, 8, 3);
    bad += expect_throw("countedAllele 3", "IllegalArgument");
    jobject direct_off = m_NewDirectByteBuffer(env, o, 4 * 8), direct_idx = m_NewDirectByteBuffer(env, idx->data, 10 * 4);
    POOL(accumulateCallsDirect)(env, NULL, 0, 1, direct_off, direct_idx, 4, 4);
    bad += expect_throw("direct offsets shorter than nv + 1", "IllegalArgument");
    POOL(accumulateCallsDirect)(env, NULL, 0, 1, direct_off, direct_idx, 3, 3);
    bad += expect_throw("idxBytes 3", "IllegalArgument");
    POOL(accumulateCallsDirect)(env, NULL, 0, 1, off, idx, 3, 4);     /* heap arrays are not direct buffers */
    bad += expect_throw("non-direct buffers", "IllegalArgument");
    jarray small = new_array(K_INT, 5);
    POOL(getGram)(env, NULL, 0, 3, small);
    bad += expect_throw("getGram output too small", "IllegalArgument");
    jarray vecs = new_array(K_DOUBLE, 5);
    POOL(computePca)(env, NULL, 0, 3, 2, vecs, NULL);
    bad += expect_throw("computePca output too small", "IllegalArgument");
    PCA(joinRows)(env, NULL, 0, 0, 2, 1, NULL, off, off, idx, 3);
    bad += expect_throw("joinRows with null key bytes", "IllegalArgument");
    PCA(joinRows)(env, NULL, 0, 0, 2, 1, rows, off, off, idx, 4);      /* nrows + 1 = 5 offsets needed, 4 present */
    bad += expect_throw("joinRows offsets shorter than nrows + 1", "IllegalArgument");
    PCA(joinFetch)(env, NULL, 0, NULL, idx);
    bad += expect_throw("joinFetch with null offsets", "IllegalArgument");
    /* bad configuration: rejected by the library, surfaced as an exception, no handle returned */
    const jlong h = POOL(create)(env, NULL, 1, 1, 0, 2, 2, 4, 2);     /* n_samples = 1 */
    if (h != 0) { fprintf(stderr, "FAIL create(n = 1) returned a handle\n"); ++bad; }
    bad += expect_throw("create with n_samples = 1", "Exception");
    free(direct_off);
    free(direct_idx);
    free_array(off); free_array(idx); free_array(idx16); free_array(rows); free_array(small); free_array(vecs);
    printf("{\"mode\": \"validate\", \"failures\": %d}\n", bad);
    return bad ? 1 : 0;
}

/* NativePca.joinRows / joinRowCount / joinCallCount / joinFetch / accumulateJoined on a hand-made pair of datasets
 * (VariantsPca.scala:115-128): left = {kA: [0,1], kB: [2], kA: [3]}, right = {kA: [4], kC: [0], kA: [5,5]} over 6 samples.
 * Inner join, left row then right row: [0,1,4] [0,1,5,5] [3,4] [3,5,5]. */
static int run_join(JNIEnv* env) {
    int bad = 0;
    const int n = 70;
    const jlong h = PCA(create)(env, NULL, n, 0, 0, 2, 4, 4);
    if (expect_clean("NativePca.create") || h == 0) return 1;
    const char* keys[6] = {"kA", "kB", "kA", "kA", "kC", "kA"};
    const int lens[6] = {2, 1, 1, 1, 1, 2};
    const jint calls[8] = {0, 1, 2, 3, 4, 0, 5, 5};
    jarray kb = new_array(K_BYTE, 12), ko = new_array(K_LONG, 7), ro = new_array(K_LONG, 7), ix = new_array(K_INT, 8);
    for (int i = 0; i < 6; ++i) {
        memcpy((char*)kb->data + 2 * i, keys[i], 2);
        ((jlong*)ko->data)[i + 1] = 2 * (i + 1);
        ((jlong*)ro->data)[i + 1] = ((jlong*)ro->data)[i] + lens[i];
    }
    memcpy(ix->data, calls, sizeof(calls));
    const jlong rows = PCA(joinRows)(env, NULL, h, 0, 2, 3, kb, ko, ro, ix, 6);
    bad += expect_clean("joinRows");
    const jlong ncalls = PCA(joinCallCount)(env, NULL, h);
    if (rows != 4 || PCA(joinRowCount)(env, NULL, h) != 4 || ncalls != 12) { fprintf(stderr, "FAIL join size %lld rows %lld calls\n", (long long)rows, (long long)ncalls); ++bad; }
    jarray fo = new_array(K_LONG, 5), fi = new_array(K_INT, 12);
    PCA(joinFetch)(env, NULL, h, fo, fi);
    bad += expect_clean("joinFetch");
    const jlong want_off[5] = {0, 3, 7, 9, 12};
    const jint want_idx[12] = {0, 1, 4, 0, 1, 5, 5, 3, 4, 3, 5, 5};
    if (memcmp(fo->data, want_off, sizeof(want_off)) != 0 || memcmp(fi->data, want_idx, sizeof(want_idx)) != 0) { fprintf(stderr, "FAIL joined rows differ\n"); ++bad; }
    PCA(accumulateJoined)(env, NULL, h, 11);
    bad += expect_clean("accumulateJoined");
    PCA(commit)(env, NULL, h, 11);
    PCA(finalizeGram)(env, NULL, h);
    bad += expect_clean("commit / finalizeGram");
    jarray gram = new_array(K_INT, (jlong)n * n);
    PCA(getGram)(env, NULL, h, n, gram);
    bad += expect_clean("getGram");
    int32_t* want = (int32_t*)calloc((size_t)n * n, sizeof(int32_t));
    for (int r = 0; r < 4; ++r)                              /* for (c1 <- row; c2 <- row) m(c1, c2) += 1 (:186-188) */
        for (jlong a = want_off[r]; a < want_off[r + 1]; ++a)
            for (jlong b = want_off[r]; b < want_off[r + 1]; ++b) want[(size_t)want_idx[a] * n + want_idx[b]] += 1;
    if (memcmp(gram->data, want, (size_t)n * n * sizeof(int32_t)) != 0) { fprintf(stderr, "FAIL Gram of the joined rows differs\n"); ++bad; }
    free(want);
    PCA(destroy)(env, NULL, h);
    free_array(kb); free_array(ko); free_array(ro); free_array(ix); free_array(fo); free_array(fi); free_array(gram);
    return bad;
}

/* ------------------------------------------------------------------------------------------------- gpu */
static int run_gpu(int n, int64_t nv_req, int gpus) {
    JNIEnv* env = &mock_env;
    int bad = 0;
    const uint64_t seed = 20240901ull;
    int64_t* off = (int64_t*)malloc((size_t)(nv_req + 1) * sizeof(int64_t));
    int64_t nv = 0;
    const int64_t nnz = vo_synth_calls(seed, n, 0, nv_req, off, NULL, &nv);
    int32_t* idx = (int32_t*)malloc((size_t)nnz * sizeof(int32_t));
    vo_synth_calls(seed, n, 0, nv_req, off, idx, &nv);
    const size_t nn = (size_t)n * n;
    int32_t* want = (int32_t*)malloc(nn * sizeof(int32_t));
    vo_similarity(n, nv, off, idx, 4, want);

    const jlong pool = POOL(create)(env, NULL, n, gpus, 0, 2, 1, 8, 2);
    if (expect_clean("create") || pool == 0) return 1;
    if (POOL(size)(env, NULL, pool) != gpus) { fprintf(stderr, "FAIL size\n"); ++bad; }
    /* four partitions, one per wire format: int arrays, short arrays, pinned direct buffers, bitmaps.  Partition 1 is
       first staged and aborted (a failed task), then retried. */
    const int64_t q[5] = {0, nv / 4, nv / 2, 3 * nv / 4, nv};
    for (int p = 0; p < 4; ++p) {
        const int64_t r0 = q[p], rows = q[p + 1] - q[p], e0 = off[r0], cnt = off[q[p + 1]] - e0;
        for (int attempt = 0; attempt < (p == 1 ? 2 : 1); ++attempt) {
            if (p == 0 || p == 1) {
                /* offsets deliberately NOT rebased and with a leading unused row: the shim must honour offsets[0] */
                jarray jo = new_array(K_LONG, rows + 1);
                for (int64_t v = 0; v <= rows; ++v) ((jlong*)jo->data)[v] = off[r0 + v] - e0 + 5;
                jarray ji = new_array(p == 0 ? K_INT : K_SHORT, cnt + 5);
                for (int64_t e = 0; e < cnt; ++e) {
                    if (p == 0) ((jint*)ji->data)[e + 5] = idx[e0 + e];
                    else ((jshort*)ji->data)[e + 5] = (jshort)(uint16_t)idx[e0 + e];
                }
                if (p == 0) POOL(accumulateCalls)(env, NULL, pool, p, jo, ji, rows);
                else POOL(accumulateCallsU16)(env, NULL, pool, p, jo, ji, rows);
                free_array(jo);
                free_array(ji);
            } else if (p == 2) {
                jobject bo = PCA(allocPinned)(env, NULL, (rows + 1) * 8), bi = PCA(allocPinned)(env, NULL, (cnt > 0 ? cnt : 1) * 4);
                if (expect_clean("allocPinned") || bo == NULL || bi == NULL) return 1;
                for (int64_t v = 0; v <= rows; ++v) ((int64_t*)bo->data)[v] = off[r0 + v] - e0;
                memcpy(bi->data, idx + e0, (size_t)cnt * 4);
                POOL(accumulateCallsDirect)(env, NULL, pool, p, bo, bi, rows, 4);
                PCA(freePinned)(env, NULL, bo);
                PCA(freePinned)(env, NULL, bi);
                free(bo);
                free(bi);
            } else {
                const int64_t stride = (n + 7) / 8;
                jarray jb = new_array(K_BYTE, rows * stride);
                for (int64_t v = 0; v < rows; ++v)
                    for (int64_t e = off[r0 + v]; e < off[r0 + v + 1]; ++e)
                        ((uint8_t*)jb->data)[v * stride + idx[e] / 8] |= (uint8_t)(1u << (idx[e] % 8));
                POOL(accumulateBits)(env, NULL, pool, p, jb, rows, stride);
                free_array(jb);
            }
            bad += expect_clean("accumulate");
            if (p == 1 && attempt == 0) {
                POOL(abort)(env, NULL, pool, p);
                bad += expect_clean("abort");
            }
        }
        POOL(commit)(env, NULL, pool, p);
        bad += expect_clean("commit");
    }
    /* a corrupt row must surface as IndexOutOfBoundsException (Breeze at VariantsPca.scala:188) and poison only its partition */
    {
        jarray jo = new_array(K_LONG, 2), ji = new_array(K_INT, 2);
        ((jlong*)jo->data)[1] = 2;
        ((jint*)ji->data)[0] = 1;
        ((jint*)ji->data)[1] = n + 9;
        POOL(accumulateCalls)(env, NULL, pool, 77, jo, ji, 1);
        bad += expect_throw("sample index out of range", "IndexOutOfBounds");
        free_array(jo);
        free_array(ji);
    }
    POOL(reduceAndFinalize)(env, NULL, pool);
    bad += expect_clean("reduceAndFinalize");
    jarray gram = new_array(K_INT, (jlong)nn);
    POOL(getGram)(env, NULL, pool, n, gram);
    bad += expect_clean("getGram");
    size_t diff = 0;
    for (size_t i = 0; i < nn; ++i) diff += ((jint*)gram->data)[i] != want[i];
    if (diff) { fprintf(stderr, "FAIL %zu Gram entries differ from the oracle\n", diff); ++bad; }
    jarray band = new_array(K_INT, (jlong)7 * n);
    POOL(getGramRows)(env, NULL, pool, n, 5, 7, band);
    bad += expect_clean("getGramRows");
    if (memcmp(band->data, want + (size_t)5 * n, (size_t)7 * n * 4) != 0) { fprintf(stderr, "FAIL getGramRows differs\n"); ++bad; }
    jarray vecs = new_array(K_DOUBLE, (jlong)2 * n), evals = new_array(K_DOUBLE, 2);
    const jint nz = POOL(computePca)(env, NULL, pool, n, 2, vecs, evals);
    bad += expect_clean("computePca");
    double norm = 0.0;
    for (int i = 0; i < n; ++i) norm += ((double*)vecs->data)[i] * ((double*)vecs->data)[i];
    if (norm < 0.999999 || norm > 1.000001 || ((double*)evals->data)[0] < ((double*)evals->data)[1]) { fprintf(stderr, "FAIL PCs: norm %g\n", norm); ++bad; }
    POOL(destroy)(env, NULL, pool);
    bad += run_join(env);
    printf("{\"mode\": \"gpu\", \"n_samples\": %d, \"rows\": %lld, \"gpus\": %d, \"gram_entries_differing\": %zu, \"non_zero_rows\": %d, "
           "\"eval0\": %.17g, \"eval1\": %.17g, \"failures\": %d}\n",
           n, (long long)nv, gpus, diff, (int)nz, ((double*)evals->data)[0], ((double*)evals->data)[1], bad);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "validate") == 0) return run_validate();
    if (argc >= 5 && strcmp(argv[1], "gpu") == 0) return run_gpu(atoi(argv[2]), atoll(argv[3]), atoi(argv[4]));
    fprintf(stderr, "usage: %s validate | gpu <n_samples> <variants> <gpus>\n", argv[0]);
    return 2;
}
