// libvpca C ABI (include/vpca.h): context management, host<->device staging and the call sequence
// encode -> Gram -> (host-driven all-reduce) -> symmetrize -> centering -> eigensolve.
// Mirrors the method set of the reference's VariantsPcaDriver
// (src/main/scala/com/google/cloud/genomics/spark/examples/VariantsPca.scala:81-286); see vpca.h for the mapping.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "vpca_internal.h"

using namespace vpca;

namespace {
thread_local std::string tls_error;
}

struct vpca_ctx {
    vpca_config cfg{};
    int n = 0;
    int elem_bits = 8;   // 8 = int8, 16 = bf16, 4 = packed e2m1
    int max_mult = 2;
    int num_pc = 2;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    int32_t* d_S = nullptr;
    bool own_S = false;
    bool finalized = false;
    bool pca_done = false;
    GramPlan plan;
    EigWork eig;
    bool eig_ready = false;

    struct Slot {
        int64_t pid = -1;
        int32_t* d_S = nullptr;
        bool used = false;
        int64_t nv = 0;
    };
    std::vector<Slot> slots;

    // CSR / dense staging, double buffered
    int64_t chunk_variants = 0, chunk_nnz = 0;
    int64_t panel = 8192;   // cells per panel row of the internal dense staging tiles (VPCA_PANEL)
    int64_t* d_off[2] = {nullptr, nullptr};
    int32_t* d_idx[2] = {nullptr, nullptr};
    void* d_x[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    int* d_flags = nullptr;
    int* h_flags = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_e0 = nullptr, ev_e1 = nullptr;
    bool gram_timed = false, eig_timed = false;

    int64_t total_variants = 0;   // committed + direct
    vpca_stats st{};
    std::mutex mu;
    std::string err;
};

namespace {

int fail(vpca_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    tls_error = buf;
    return code;
}

#define CUDA_OK(ctx, call)                                                                                      \
    do {                                                                                                        \
        cudaError_t _e = (call);                                                                                \
        if (_e != cudaSuccess)                                                                                  \
            return fail(ctx, VPCA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__,   \
                        __LINE__);                                                                              \
    } while (0)

int check_overflow(vpca_ctx* ctx, int64_t extra_variants) {
    // every similarity count is at most (#variants) * max_mult^2 and must stay a Java Int (VariantsPca.scala:185)
    const long double worst = (long double)(ctx->total_variants + extra_variants) * ctx->max_mult * ctx->max_mult;
    if (worst > 2147483647.0L)
        return fail(ctx, VPCA_ERR_OVERFLOW, "%lld variants x multiplicity %d^2 could overflow an int32 similarity count",
                    (long long)(ctx->total_variants + extra_variants), ctx->max_mult);
    return VPCA_OK;
}

int ensure_staging(vpca_ctx* ctx) {
    if (ctx->d_x[0] != nullptr) return VPCA_OK;
    const int n = ctx->n, bits = ctx->elem_bits;
    int64_t cv = ctx->cfg.chunk_variants;
    if (cv <= 0) {
        cv = (256ll << 20) * 8 / ((int64_t)n * bits);
        cv = std::max<int64_t>(1024, std::min<int64_t>(cv, 1 << 20));
    }
    if (const char* pe = getenv("VPCA_PANEL")) {
        const int64_t pv = atoll(pe);
        if (pv >= 128 && pv % 128 == 0) ctx->panel = pv;
    }
    cv = std::max<int64_t>(ctx->panel, (cv / ctx->panel) * ctx->panel);   // whole panels
    int64_t cz = ctx->cfg.chunk_nnz;
    if (cz <= 0) cz = 64ll << 20;
    cz = std::max<int64_t>(cz, 1024);
    ctx->chunk_variants = cv;
    ctx->chunk_nnz = cz;
    for (int b = 0; b < 2; ++b) {
        CUDA_OK(ctx, cudaMalloc(&ctx->d_off[b], (size_t)(cv + 1) * sizeof(int64_t)));
        CUDA_OK(ctx, cudaMalloc(&ctx->d_idx[b], (size_t)cz * sizeof(int32_t)));
        CUDA_OK(ctx, cudaMalloc(&ctx->d_x[b], (size_t)n * (size_t)cv * bits / 8));
        CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_copy[b], cudaEventDisableTiming));
        CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_done[b], cudaEventDisableTiming));
    }
    CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    return VPCA_OK;
}

int launch_gram(vpca_ctx* ctx, const void* d_x, int64_t nv, int64_t ld, int64_t panel, int32_t* d_target) {
    // fp32 TMEM accumulation (bf16 / e2m1) is exact only below 2^24: bound the variants one launch may fold
    int64_t limit = nv;
    if (ctx->elem_bits != 8) {
        limit = (int64_t)(16777216ll / ((int64_t)ctx->max_mult * ctx->max_mult));
        const int64_t q = panel > 0 ? panel : 128;
        limit = std::max<int64_t>(q, (limit / q) * q);
    }
    for (int64_t v0 = 0; v0 < nv; v0 += limit) {
        const int64_t cnt = std::min<int64_t>(limit, nv - v0);
        std::string msg;
        cudaEventRecord(ctx->ev_t0, ctx->stream);
        // sub-launches start on a panel boundary (panel layout) or at column v0 (row-major)
        const size_t byte_off = panel > 0 ? (size_t)(v0 / panel) * (size_t)ctx->n * (size_t)panel * ctx->elem_bits / 8
                                          : (size_t)v0 * ctx->elem_bits / 8;
        cudaError_t e = gram_accumulate(ctx->plan, static_cast<const char*>(d_x) + byte_off, ctx->elem_bits, ctx->n, cnt, ld,
                                        panel, d_target, ctx->stream, &msg);
        cudaEventRecord(ctx->ev_t1, ctx->stream);
        if (e != cudaSuccess)
            return fail(ctx, VPCA_ERR_CUDA, "Gram launch failed: %s %s", cudaGetErrorString(e), msg.c_str());
        ctx->gram_timed = true;
        ctx->st.gram_launches += 1;
        ctx->st.kernel_launches += 1;
    }
    ctx->st.gram_cta_group = ctx->plan.cta_group;
    ctx->st.gram_resident = ctx->plan.last_resident;
    return VPCA_OK;
}

vpca_ctx::Slot* find_slot(vpca_ctx* ctx, int64_t pid, bool create, int* rc) {
    *rc = VPCA_OK;
    for (auto& s : ctx->slots)
        if (s.used && s.pid == pid) return &s;
    if (!create) return nullptr;
    for (auto& s : ctx->slots)
        if (!s.used) {
            if (s.d_S == nullptr) {
                cudaError_t e = cudaMalloc(&s.d_S, (size_t)ctx->n * ctx->n * sizeof(int32_t));
                if (e != cudaSuccess) {
                    *rc = fail(ctx, VPCA_ERR_NOMEM, "cudaMalloc of a partition Gram failed: %s", cudaGetErrorString(e));
                    return nullptr;
                }
            }
            cudaError_t e = cudaMemsetAsync(s.d_S, 0, (size_t)ctx->n * ctx->n * sizeof(int32_t), ctx->stream);
            if (e != cudaSuccess) {
                *rc = fail(ctx, VPCA_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
                return nullptr;
            }
            s.used = true;
            s.pid = pid;
            s.nv = 0;
            return &s;
        }
    *rc = fail(ctx, VPCA_ERR_STATE, "more than %d partitions in flight; commit or abort one first",
               (int)ctx->slots.size());
    return nullptr;
}

// CSR rows -> encode -> (optionally) Gram.  out_tile != nullptr: copy the encoded tile back instead of the Gram.
int process_calls(vpca_ctx* ctx, const int64_t* offsets, const void* sample_idx, int idx_bytes, int64_t nv,
                  int32_t* d_target, void* out_tile, int64_t out_ld) {
    int rc = ensure_staging(ctx);
    if (rc != VPCA_OK) return rc;
    const int bits = ctx->elem_bits;
    if (offsets[0] < 0) return fail(ctx, VPCA_ERR_BAD_ARG, "offsets[0] must be >= 0");
    *ctx->h_flags = 0;
    CUDA_OK(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int64_t v = 0;
    int chunk = 0;
    while (v < nv) {
        // largest run of rows that fits both the variant and the index budget
        int64_t vend = std::min(nv, v + ctx->chunk_variants);
        if (offsets[vend] - offsets[v] > ctx->chunk_nnz) {
            const int64_t* hi = std::upper_bound(offsets + v, offsets + vend + 1, offsets[v] + ctx->chunk_nnz);
            vend = (hi - offsets) - 1;
            if (bits == 4 && vend - v >= 128) vend = v + ((vend - v) / 128) * 128;   // keep packed rows byte aligned
            if (vend <= v)
                return fail(ctx, VPCA_ERR_BAD_ARG, "row %lld has %lld entries, more than chunk_nnz=%lld", (long long)v,
                            (long long)(offsets[v + 1] - offsets[v]), (long long)ctx->chunk_nnz);
        }
        for (int64_t q = v; q < vend; ++q)
            if (offsets[q + 1] < offsets[q]) return fail(ctx, VPCA_ERR_BAD_ARG, "offsets must be non-decreasing");
        const int64_t nvc = vend - v, nnz = offsets[vend] - offsets[v];
        const int b = chunk & 1;
        // the copy stream may overwrite buffer b only after the kernels that read it have run
        CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_off[b], offsets + v, (size_t)(nvc + 1) * sizeof(int64_t),
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
        if (nnz > 0)
            CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_idx[b], static_cast<const char*>(sample_idx) + (size_t)offsets[v] * idx_bytes,
                                         (size_t)nnz * idx_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
        ctx->st.h2d_bytes += (nvc + 1) * 8 + nnz * idx_bytes;
        CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
        const int64_t P = ctx->panel;
        CUDA_OK(ctx, encode_calls(ctx->d_off[b], offsets[v], ctx->d_idx[b], idx_bytes, nvc, ctx->n, bits, ctx->max_mult, ctx->d_x[b], P,
                                  P, ctx->d_flags, ctx->stream));
        ctx->st.kernel_launches += 2;
        if (out_tile != nullptr) {
            // panel layout -> the caller's row-major tile, one 2-D copy per panel (chunk boundaries are multiples of
            // 128 variants, so 4-bit rows split on byte boundaries)
            for (int64_t pv = 0; pv < nvc; pv += P) {
                const int64_t wv = std::min(P, nvc - pv);
                CUDA_OK(ctx, cudaMemcpy2DAsync(static_cast<char*>(out_tile) + (size_t)(v + pv) * bits / 8,
                                               (size_t)out_ld * bits / 8,
                                               static_cast<const char*>(ctx->d_x[b]) + (size_t)(pv / P) * ctx->n * P * bits / 8,
                                               (size_t)P * bits / 8, (size_t)(wv * bits + 7) / 8, (size_t)ctx->n,
                                               cudaMemcpyDeviceToHost, ctx->stream));
            }
            ctx->st.d2h_bytes += (nvc * bits + 7) / 8 * (int64_t)ctx->n;
        } else {
            rc = launch_gram(ctx, ctx->d_x[b], nvc, P, P, d_target);
            if (rc != VPCA_OK) return rc;
        }
        CUDA_OK(ctx, cudaEventRecord(ctx->ev_done[b], ctx->stream));
        v = vend;
        ++chunk;
    }
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    // the caller's buffers are read asynchronously: do not return before every copy has completed
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flags & 1)
        return fail(ctx, VPCA_ERR_INDEX_OUT_OF_RANGE, "sample index outside [0, %d) (the reference throws at "
                    "VariantsPca.scala:59/:188)", ctx->n);
    if (*ctx->h_flags & 2)
        return fail(ctx, VPCA_ERR_OVERFLOW, "a sample is listed more than max_multiplicity=%d times in one row",
                    ctx->max_mult);
    return VPCA_OK;
}

}  // namespace

extern "C" {

int vpca_version(void) { return VPCA_VERSION_MAJOR * 1000 + VPCA_VERSION_MINOR; }

const char* vpca_last_error(const vpca_ctx* ctx) {
    if (ctx != nullptr) return ctx->err.c_str();
    return tls_error.c_str();
}

int vpca_create(const vpca_config* cfg, vpca_ctx** out) {
    if (out == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (cfg == nullptr || cfg->struct_size != sizeof(vpca_config))
        return fail(nullptr, VPCA_ERR_BAD_ARG, "cfg is NULL or struct_size != sizeof(vpca_config) (%zu)",
                    sizeof(vpca_config));
    if (cfg->n_samples < 2) return fail(nullptr, VPCA_ERR_BAD_ARG, "n_samples must be >= 2");
    if (cfg->dtype != VPCA_DTYPE_I8 && cfg->dtype != VPCA_DTYPE_BF16 && cfg->dtype != VPCA_DTYPE_E2M1)
        return fail(nullptr, VPCA_ERR_BAD_ARG, "unknown dtype %d", cfg->dtype);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VPCA_ERR_CUDA, "no CUDA device: %s (libvpca has no CPU fallback)", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, VPCA_ERR_BAD_ARG, "device %d of %d", cfg->device, ndev);
    CUDA_OK(nullptr, cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_OK(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(nullptr, VPCA_ERR_UNSUPPORTED, "device %d is sm_%d%d; libvpca is built for sm_100a (B200) only",
                    cfg->device, prop.major, prop.minor);
    vpca_ctx* ctx = new (std::nothrow) vpca_ctx();
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_NOMEM, "out of host memory");
    ctx->cfg = *cfg;
    ctx->n = cfg->n_samples;
    ctx->elem_bits = cfg->dtype == VPCA_DTYPE_I8 ? 8 : (cfg->dtype == VPCA_DTYPE_BF16 ? 16 : 4);
    ctx->max_mult = cfg->max_multiplicity > 0 ? cfg->max_multiplicity : 2;
    ctx->num_pc = cfg->num_pc > 0 ? cfg->num_pc : 2;
    if (ctx->elem_bits == 4 && ctx->max_mult > 2) {
        delete ctx;
        return fail(nullptr, VPCA_ERR_BAD_ARG, "VPCA_DTYPE_E2M1 represents multiplicities 0, 1, 2 only (max_multiplicity <= 2)");
    }
    if (cfg->stream != nullptr) {
        ctx->stream = static_cast<cudaStream_t>(cfg->stream);
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        ctx->own_stream = true;
    }
    const size_t gram_bytes = (size_t)ctx->n * ctx->n * sizeof(int32_t);
    if (e == cudaSuccess) {
        if (cfg->d_gram != nullptr) {
            ctx->d_S = static_cast<int32_t*>(cfg->d_gram);
        } else {
            e = cudaMalloc(&ctx->d_S, gram_bytes + 64 * sizeof(int32_t));   // + barrier flags of the peer-reduce mode
            ctx->own_S = true;
            if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_S + (size_t)ctx->n * ctx->n, 0, 64 * sizeof(int32_t), ctx->stream);
        }
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_S, 0, gram_bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_flags, sizeof(int));
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_flags, sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_t1);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_e0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_e1);
    if (e != cudaSuccess) {
        const int rc = fail(nullptr, VPCA_ERR_CUDA, "vpca_create: %s", cudaGetErrorString(e));
        vpca_destroy(ctx);
        return rc;
    }
    const int nslots = cfg->partitions_in_flight > 0 ? cfg->partitions_in_flight : 4;
    ctx->slots.resize(nslots);
    *out = ctx;
    return VPCA_OK;
}

int vpca_destroy(vpca_ctx* ctx) {
    if (ctx == nullptr) return VPCA_OK;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->d_off[b]);
        cudaFree(ctx->d_idx[b]);
        cudaFree(ctx->d_x[b]);
        if (ctx->ev_copy[b]) cudaEventDestroy(ctx->ev_copy[b]);
        if (ctx->ev_done[b]) cudaEventDestroy(ctx->ev_done[b]);
    }
    for (auto& s : ctx->slots) cudaFree(s.d_S);
    for (int d = 0; d < ctx->plan.num_peers; ++d)
        if (d != ctx->plan.peer_rank && ctx->plan.peer_S[d] != nullptr) cudaIpcCloseMemHandle(ctx->plan.peer_S[d]);
    if (ctx->own_S) cudaFree(ctx->d_S);
    cudaFree(ctx->d_flags);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->eig_ready) eig_free(ctx->eig);
    gram_plan_free(ctx->plan);
    for (cudaEvent_t ev : {ctx->ev_t0, ctx->ev_t1, ctx->ev_e0, ctx->ev_e1})
        if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VPCA_OK;
}

int vpca_reset(vpca_ctx* ctx) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, cudaMemsetAsync(ctx->d_S, 0, (size_t)ctx->n * ctx->n * sizeof(int32_t), ctx->stream));
    for (auto& s : ctx->slots) s.used = false;
    ctx->finalized = false;
    ctx->pca_done = false;
    ctx->total_variants = 0;
    ctx->st.variants_accumulated = 0;
    return VPCA_OK;
}

int vpca_encode_calls(vpca_ctx* ctx, const int64_t* offsets, const int32_t* sample_idx, int64_t nv, void* out,
                      int64_t ld) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (offsets == nullptr || out == nullptr || nv < 0 || ld < nv || (nv > 0 && sample_idx == nullptr && offsets[nv] > offsets[0]))
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_encode_calls: bad argument");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    if (nv == 0) return VPCA_OK;
    return process_calls(ctx, offsets, sample_idx, 4, nv, nullptr, out, ld);
}

static int accumulate_calls_impl(vpca_ctx* ctx, int64_t partition_id, const int64_t* offsets, const void* sample_idx,
                                 int idx_bytes, int64_t nv);

int vpca_accumulate_calls(vpca_ctx* ctx, int64_t partition_id, const int64_t* offsets, const int32_t* sample_idx,
                          int64_t nv) {
    return accumulate_calls_impl(ctx, partition_id, offsets, sample_idx, 4, nv);
}

int vpca_accumulate_calls_u16(vpca_ctx* ctx, int64_t partition_id, const int64_t* offsets, const uint16_t* sample_idx,
                              int64_t nv) {
    if (ctx != nullptr && ctx->n > 65536) return fail(ctx, VPCA_ERR_BAD_ARG, "16-bit sample indices need n_samples <= 65536");
    return accumulate_calls_impl(ctx, partition_id, offsets, sample_idx, 2, nv);
}

static int accumulate_calls_impl(vpca_ctx* ctx, int64_t partition_id, const int64_t* offsets, const void* sample_idx,
                                 int idx_bytes, int64_t nv) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (offsets == nullptr || nv < 0 || (nv > 0 && sample_idx == nullptr && offsets[nv] > offsets[0]))
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_accumulate_calls: bad argument");
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized; call vpca_reset first");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    if (nv == 0) return VPCA_OK;
    int rc = check_overflow(ctx, nv);
    if (rc != VPCA_OK) return rc;
    int32_t* target = ctx->d_S;
    vpca_ctx::Slot* slot = nullptr;
    if (partition_id >= 0) {
        slot = find_slot(ctx, partition_id, true, &rc);
        if (slot == nullptr) return rc;
        target = slot->d_S;
    }
    rc = process_calls(ctx, offsets, sample_idx, idx_bytes, nv, target, nullptr, 0);
    if (rc != VPCA_OK) {
        // a failed batch poisons the partition (its staging Gram may be partially updated): drop it
        if (slot) slot->used = false;
        else if (rc == VPCA_ERR_INDEX_OUT_OF_RANGE || rc == VPCA_ERR_OVERFLOW)
            ctx->err += " [direct accumulation: the Gram may hold a partial batch, call vpca_reset]";
        return rc;
    }
    if (slot) slot->nv += nv;
    else ctx->total_variants += nv;
    ctx->st.variants_accumulated += nv;
    return VPCA_OK;
}

// code 0: bitmap rows; 1 / 2: PLINK .bed rows counting A1 / A2 (see encode.cu)
static int accumulate_packed(vpca_ctx* ctx, int64_t partition_id, const uint8_t* bits, int64_t nv, int64_t stride_bytes,
                             int code) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int64_t min_stride = code == 0 ? (ctx->n + 7) / 8 : (ctx->n + 3) / 4;
    if (nv < 0 || (nv > 0 && bits == nullptr) || stride_bytes < min_stride)
        return fail(ctx, VPCA_ERR_BAD_ARG, "packed rows: stride_bytes must be >= ceil(n_samples / %d)", code == 0 ? 8 : 4);
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized; call vpca_reset first");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    if (nv == 0) return VPCA_OK;
    int rc = check_overflow(ctx, nv);
    if (rc != VPCA_OK) return rc;
    rc = ensure_staging(ctx);
    if (rc != VPCA_OK) return rc;
    int32_t* target = ctx->d_S;
    vpca_ctx::Slot* slot = nullptr;
    if (partition_id >= 0) {
        slot = find_slot(ctx, partition_id, true, &rc);
        if (slot == nullptr) return rc;
        target = slot->d_S;
    }
    // bits beyond sample n-1 in the last byte of a row would be read as carriers of non-existent samples: the kernel
    // masks them (smp >= n), nothing to validate on the host.
    const int64_t P = ctx->panel;
    const int64_t cap_rows = std::min<int64_t>(ctx->chunk_variants, (ctx->chunk_nnz * (int64_t)sizeof(int32_t)) / stride_bytes);
    if (cap_rows < 32) {
        if (slot) slot->used = false;
        return fail(ctx, VPCA_ERR_BAD_ARG, "stride_bytes too large for the staging buffer");
    }
    const int64_t step = std::max<int64_t>(P, (cap_rows / P) * P) <= cap_rows ? std::max<int64_t>(P, (cap_rows / P) * P)
                                                                              : (cap_rows / 32) * 32;
    int chunk = 0;
    for (int64_t v = 0; v < nv; v += step, ++chunk) {
        const int64_t nvc = std::min(step, nv - v);
        const int b = chunk & 1;
        CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_idx[b], bits + (size_t)v * stride_bytes, (size_t)nvc * stride_bytes,
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
        ctx->st.h2d_bytes += nvc * stride_bytes;
        CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
        CUDA_OK(ctx, encode_bits(reinterpret_cast<const uint8_t*>(ctx->d_idx[b]), stride_bytes, nvc, ctx->n, ctx->elem_bits,
                                 ctx->d_x[b], P, P, code, ctx->stream));
        ctx->st.kernel_launches += 1;
        rc = launch_gram(ctx, ctx->d_x[b], nvc, P, P, target);
        if (rc != VPCA_OK) {
            if (slot) slot->used = false;
            return rc;
        }
        CUDA_OK(ctx, cudaEventRecord(ctx->ev_done[b], ctx->stream));
    }
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // the caller's buffer is free to reuse on return
    if (slot) slot->nv += nv;
    else ctx->total_variants += nv;
    ctx->st.variants_accumulated += nv;
    return VPCA_OK;
}

int vpca_accumulate_bits(vpca_ctx* ctx, int64_t partition_id, const uint8_t* bits, int64_t nv, int64_t stride_bytes) {
    return accumulate_packed(ctx, partition_id, bits, nv, stride_bytes, 0);
}

int vpca_accumulate_bed(vpca_ctx* ctx, int64_t partition_id, const uint8_t* rows, int64_t nv, int64_t stride_bytes,
                        int32_t counted_allele) {
    if (counted_allele != 1 && counted_allele != 2)
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_accumulate_bed: counted_allele must be 1 (A1) or 2 (A2)");
    return accumulate_packed(ctx, partition_id, rows, nv, stride_bytes, counted_allele);
}

int vpca_commit(vpca_ctx* ctx, int64_t partition_id) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized");
    int rc;
    vpca_ctx::Slot* s = find_slot(ctx, partition_id, false, &rc);
    if (s == nullptr) return VPCA_OK;   // an empty partition never staged anything
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    rc = check_overflow(ctx, 0);
    if (rc != VPCA_OK) return rc;
    if (ctx->plan.num_peers > 1 && ctx->plan.peer_mode == 1)
        CUDA_OK(ctx, gram_add_owners(ctx->plan, s->d_S, ctx->n, ctx->stream));
    else if (ctx->plan.num_peers > 1)
        CUDA_OK(ctx, gram_add_peers(ctx->plan, s->d_S, (int64_t)ctx->n * ctx->n, ctx->stream));
    else
        CUDA_OK(ctx, gram_add(ctx->d_S, s->d_S, (int64_t)ctx->n * ctx->n, ctx->stream));
    ctx->st.kernel_launches += 1;
    ctx->total_variants += s->nv;
    s->used = false;
    return VPCA_OK;
}

int vpca_abort(vpca_ctx* ctx, int64_t partition_id) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc;
    vpca_ctx::Slot* s = find_slot(ctx, partition_id, false, &rc);
    if (s != nullptr) {
        ctx->st.variants_accumulated -= s->nv;
        s->used = false;
    }
    return VPCA_OK;
}

int vpca_accumulate_dense(vpca_ctx* ctx, const void* x, int64_t nv, int64_t ld, int on_device) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (x == nullptr || nv < 0 || ld < nv) return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_accumulate_dense: bad argument");
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized; call vpca_reset first");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    if (nv == 0) return VPCA_OK;
    int rc = check_overflow(ctx, nv);
    if (rc != VPCA_OK) return rc;
    const int bits = ctx->elem_bits;
    if (bits == 4 && (ld % 128) != 0)
        return fail(ctx, VPCA_ERR_BAD_ARG, "packed e2m1 tiles need ld %% 128 == 0 (and zero padding up to a multiple of 128 variants)");
    if (on_device) {
        const int align = bits == 4 ? 31 : 15;
        if ((reinterpret_cast<uintptr_t>(x) & align) != 0 || ((ld * bits / 8) & align) != 0)
            return fail(ctx, VPCA_ERR_BAD_ARG, "device tile must be %d-byte aligned with a %d-byte multiple row pitch",
                        align + 1, align + 1);
        rc = launch_gram(ctx, x, nv, ld, 0, ctx->d_S);
        if (rc != VPCA_OK) return rc;
    } else {
        rc = ensure_staging(ctx);
        if (rc != VPCA_OK) return rc;
        int chunk = 0;
        for (int64_t v = 0; v < nv; v += ctx->chunk_variants, ++chunk) {
            const int64_t nvc = std::min(ctx->chunk_variants, nv - v);
            const int b = chunk & 1;
            CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
            // the caller's row-major tile -> panel layout, one 2-D copy per panel; a partial last panel is zeroed first
            const int64_t P = ctx->panel;
            if ((nvc % P) != 0)
                CUDA_OK(ctx, cudaMemsetAsync(static_cast<char*>(ctx->d_x[b]) + (size_t)(nvc / P) * ctx->n * P * bits / 8, 0,
                                             (size_t)ctx->n * P * bits / 8, ctx->copy_stream));
            for (int64_t pv = 0; pv < nvc; pv += P) {
                const int64_t wv = std::min(P, nvc - pv);
                CUDA_OK(ctx, cudaMemcpy2DAsync(static_cast<char*>(ctx->d_x[b]) + (size_t)(pv / P) * ctx->n * P * bits / 8,
                                               (size_t)P * bits / 8, static_cast<const char*>(x) + (size_t)(v + pv) * bits / 8,
                                               (size_t)ld * bits / 8, (size_t)(wv * bits + 7) / 8, (size_t)ctx->n,
                                               cudaMemcpyHostToDevice, ctx->copy_stream));
            }
            CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
            ctx->st.h2d_bytes += (nvc * bits + 7) / 8 * (int64_t)ctx->n;
            CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
            rc = launch_gram(ctx, ctx->d_x[b], nvc, P, P, ctx->d_S);
            if (rc != VPCA_OK) return rc;
            CUDA_OK(ctx, cudaEventRecord(ctx->ev_done[b], ctx->stream));
        }
        CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // caller's buffer is free to reuse on return
    }
    ctx->total_variants += nv;
    ctx->st.variants_accumulated += nv;
    return VPCA_OK;
}

int vpca_accumulate_panels(vpca_ctx* ctx, const void* d_x, int64_t nv, int64_t panel_variants) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (d_x == nullptr || nv < 0 || panel_variants < 128 || (panel_variants % 128) != 0)
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_accumulate_panels: panel_variants must be a positive multiple of 128");
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized; call vpca_reset first");
    if ((reinterpret_cast<uintptr_t>(d_x) & 31) != 0) return fail(ctx, VPCA_ERR_BAD_ARG, "panels must be 32-byte aligned");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    if (nv == 0) return VPCA_OK;
    int rc = check_overflow(ctx, nv);
    if (rc != VPCA_OK) return rc;
    rc = launch_gram(ctx, d_x, nv, panel_variants, panel_variants, ctx->d_S);
    if (rc != VPCA_OK) return rc;
    ctx->total_variants += nv;
    ctx->st.variants_accumulated += nv;
    return VPCA_OK;
}

int vpca_synth_panels_device(vpca_ctx* ctx, uint64_t seed, int64_t v0, int64_t nv, int mode, void* d_x,
                             int64_t panel_variants) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (d_x == nullptr || nv < 0 || v0 < 0 || (mode != 0 && mode != 1) || panel_variants < 128 || (panel_variants % 128) != 0)
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_synth_panels_device: bad argument");
    if (mode == 1 && ctx->max_mult < 2) return fail(ctx, VPCA_ERR_BAD_ARG, "dosage mode needs max_multiplicity >= 2");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    cudaError_t e = synth_dense(seed, ctx->n, v0, nv, mode, ctx->elem_bits, d_x, panel_variants, panel_variants, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, VPCA_ERR_CUDA, "synthetic generator: %s", cudaGetErrorString(e));
    ctx->st.kernel_launches += 2 * ((nv + (1 << 22) - 1) >> 22);
    return VPCA_OK;
}

int vpca_gram_device_ptr(vpca_ctx* ctx, void** d_gram) {
    if (ctx == nullptr || d_gram == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    *d_gram = ctx->d_S;
    return VPCA_OK;
}

int vpca_finalize_gram(vpca_ctx* ctx) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->finalized) return VPCA_OK;
    for (auto& s : ctx->slots)
        if (s.used)
            return fail(ctx, VPCA_ERR_STATE, "partition %lld is neither committed nor aborted", (long long)s.pid);
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, gram_symmetrize(ctx->d_S, ctx->n, ctx->stream));
    ctx->st.kernel_launches += 1;
    ctx->finalized = true;
    ctx->pca_done = false;
    return VPCA_OK;
}

int vpca_get_gram(vpca_ctx* ctx, int32_t* out) {
    if (ctx == nullptr || out == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "call vpca_finalize_gram first");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t bytes = (size_t)ctx->n * ctx->n * sizeof(int32_t);
    CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->d_S, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += (int64_t)bytes;
    return VPCA_OK;
}

int vpca_get_partial_gram(vpca_ctx* ctx, int32_t* out) {
    if (ctx == nullptr || out == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized: use vpca_get_gram");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t bytes = (size_t)ctx->n * ctx->n * sizeof(int32_t);
    CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->d_S, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += (int64_t)bytes;
    return VPCA_OK;
}

int vpca_load_partial_gram(vpca_ctx* ctx, const int32_t* gram) {
    if (ctx == nullptr || gram == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized; call vpca_reset first");
    for (auto& s : ctx->slots)
        if (s.used) return fail(ctx, VPCA_ERR_STATE, "partition %lld is in flight", (long long)s.pid);
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t bytes = (size_t)ctx->n * ctx->n * sizeof(int32_t);
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_S, gram, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->st.h2d_bytes += (int64_t)bytes;
    return VPCA_OK;
}

int vpca_set_gram(vpca_ctx* ctx, const int32_t* gram) {
    if (ctx == nullptr || gram == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t bytes = (size_t)ctx->n * ctx->n * sizeof(int32_t);
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_S, gram, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->st.h2d_bytes += (int64_t)bytes;
    for (auto& s : ctx->slots) s.used = false;
    ctx->finalized = true;
    ctx->pca_done = false;
    return VPCA_OK;
}

static int run_center(vpca_ctx* ctx) {
    if (!ctx->eig_ready) {
        cudaError_t e = eig_alloc(ctx->eig, ctx->n, std::max(ctx->num_pc, 16));
        if (e != cudaSuccess) {
            eig_free(ctx->eig);
            return fail(ctx, VPCA_ERR_NOMEM, "eigensolver workspace: %s", cudaGetErrorString(e));
        }
        ctx->eig_ready = true;
    }
    CUDA_OK(ctx, center_gram(ctx->eig, ctx->d_S, ctx->stream));
    ctx->st.kernel_launches += 3;
    return VPCA_OK;
}

int vpca_compute_pca(vpca_ctx* ctx, int32_t k, double* vecs, double* evals, int32_t* non_zero_rows) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (vecs == nullptr || k < 1 || k > ctx->n || k > std::max(ctx->num_pc, 16))
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_compute_pca: k=%d out of range", k);
    if (!ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "call vpca_finalize_gram first");
    if (ctx->n > 65535)
        return fail(ctx, VPCA_ERR_UNSUPPORTED, "computePca is limited to 65535 samples, like the reference (MLlib RowMatrix "
                    "behind VariantsPca.scala:226 refuses more columns); the Gram itself has no such limit");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_e0, ctx->stream));
    int rc = run_center(ctx);
    if (rc != VPCA_OK) return rc;
    {   // VPCA_EIG=direct|lanczos|auto (default auto: Lanczos from 512 samples up, direct reduction as its fallback)
        const char* em = getenv("VPCA_EIG");
        ctx->eig.mode = (em != nullptr && strcmp(em, "direct") == 0) ? 1 : (em != nullptr && strcmp(em, "lanczos") == 0) ? 2 : 0;
    }
    CUDA_OK(ctx, eig_topk(ctx->eig, k, ctx->stream, &ctx->st.kernel_launches));
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_e1, ctx->stream));
    ctx->st.eig_method = ctx->eig.last_method;
    ctx->st.eig_iterations = ctx->eig.last_iters;
    ctx->eig_timed = true;
    const size_t nb = (size_t)ctx->n * k * sizeof(double);
    CUDA_OK(ctx, cudaMemcpyAsync(vecs, ctx->eig.d_evecs, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (evals) CUDA_OK(ctx, cudaMemcpyAsync(evals, ctx->eig.d_evals, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    int nz = 0;
    CUDA_OK(ctx, cudaMemcpyAsync(&nz, ctx->eig.d_nz, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (non_zero_rows) *non_zero_rows = nz;
    ctx->st.d2h_bytes += (int64_t)nb + (evals ? k * 8 : 0) + 4;
    ctx->pca_done = true;
    return VPCA_OK;
}

int vpca_get_centered(vpca_ctx* ctx, double* out) {
    if (ctx == nullptr || out == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "call vpca_finalize_gram first");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    int rc = run_center(ctx);   // the eigensolve overwrites C, so recompute it
    if (rc != VPCA_OK) return rc;
    const size_t bytes = (size_t)ctx->n * ctx->n * sizeof(double);
    CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->eig.d_C, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += (int64_t)bytes;
    ctx->pca_done = false;
    return VPCA_OK;
}

int vpca_get_tridiagonal(vpca_ctx* ctx, double* diag, double* offdiag) {
    if (ctx == nullptr || diag == nullptr || offdiag == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->pca_done) return fail(ctx, VPCA_ERR_STATE, "call vpca_compute_pca first");
    if (ctx->eig.last_method == 2)
        return fail(ctx, VPCA_ERR_STATE, "the last solve used Lanczos and did not tridiagonalise C (set VPCA_EIG=direct)");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, cudaMemcpyAsync(diag, ctx->eig.d_diag, (size_t)ctx->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(offdiag, ctx->eig.d_off, (size_t)(ctx->n - 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return VPCA_OK;
}

int vpca_synth_dense_device(vpca_ctx* ctx, uint64_t seed, int64_t v0, int64_t nv, int mode, void* d_x, int64_t ld) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (d_x == nullptr || nv < 0 || ld < nv || v0 < 0 || (mode != 0 && mode != 1))
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_synth_dense_device: bad argument");
    if (mode == 1 && ctx->max_mult < 2)
        return fail(ctx, VPCA_ERR_BAD_ARG, "dosage mode needs max_multiplicity >= 2");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    cudaError_t e = synth_dense(seed, ctx->n, v0, nv, mode, ctx->elem_bits, d_x, ld, 0, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, VPCA_ERR_CUDA, "synthetic generator: %s", cudaGetErrorString(e));
    ctx->st.kernel_launches += 2 * ((nv + (1 << 22) - 1) >> 22);
    return VPCA_OK;
}

int vpca_get_stats(vpca_ctx* ctx, vpca_stats* out) {
    if (ctx == nullptr || out == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    if (ctx->gram_timed && cudaEventSynchronize(ctx->ev_t1) == cudaSuccess) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1) == cudaSuccess) ctx->st.last_gram_ms = ms;
    }
    if (ctx->eig_timed && cudaEventSynchronize(ctx->ev_e1) == cudaSuccess) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_e0, ctx->ev_e1) == cudaSuccess) ctx->st.last_eig_ms = ms;
    }
    *out = ctx->st;
    return VPCA_OK;
}

int vpca_gram_export_ipc(vpca_ctx* ctx, void* handle64) {
    if (ctx == nullptr || handle64 == nullptr) return fail(ctx, VPCA_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->own_S) return fail(ctx, VPCA_ERR_STATE, "the peer-reduce mode needs a library-owned Gram (vpca_config.d_gram == NULL)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    cudaIpcMemHandle_t h;
    CUDA_OK(ctx, cudaIpcGetMemHandle(&h, ctx->d_S));
    memcpy(handle64, &h, sizeof(h));
    return VPCA_OK;
}

int vpca_gram_set_peers(vpca_ctx* ctx, const void* handles, int32_t world, int32_t rank) {
    if (ctx == nullptr || handles == nullptr || world < 1 || world > 16 || rank < 0 || rank >= world)
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_gram_set_peers: bad argument (world <= 16)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->own_S) return fail(ctx, VPCA_ERR_STATE, "the peer-reduce mode needs a library-owned Gram");
    if (ctx->plan.num_peers != 0) return fail(ctx, VPCA_ERR_STATE, "peers already set");
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t nn = (size_t)ctx->n * ctx->n;
    for (int d = 0; d < world; ++d) {
        int32_t* base = ctx->d_S;
        if (d != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const char*>(handles) + (size_t)d * sizeof(h), sizeof(h));
            void* ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int q = 0; q < d; ++q)
                    if (q != rank) cudaIpcCloseMemHandle(ctx->plan.peer_S[q]);
                return fail(ctx, VPCA_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", d, cudaGetErrorString(e));
            }
            base = static_cast<int32_t*>(ptr);
        }
        ctx->plan.peer_S[d] = base;
        ctx->plan.peer_flags[d] = base + nn;
    }
    ctx->plan.peer_rank = rank;
    ctx->plan.num_peers = world;
    return VPCA_OK;
}

int vpca_gram_set_peer_mode(vpca_ctx* ctx, int32_t mode) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (mode != VPCA_PEER_REPLICATE && mode != VPCA_PEER_OWNER_ROWS)
        return fail(ctx, VPCA_ERR_BAD_ARG, "vpca_gram_set_peer_mode: unknown mode %d", mode);
    if (ctx->plan.num_peers < 1) return fail(ctx, VPCA_ERR_STATE, "call vpca_gram_set_peers first");
    const int world = ctx->plan.num_peers, n = ctx->n;
    if (mode == VPCA_PEER_OWNER_ROWS) {
        if (n < 64 * world)
            return fail(ctx, VPCA_ERR_UNSUPPORTED, "owner-rows mode needs n_samples >= 64 x world (%d < %d)", n, 64 * world);
        // equal shares of the lower triangle: rows [0, R) hold R^2 / 2 cells -> R_q = n sqrt(q / world), on multiples of 32
        int prev = 0;
        for (int q = 0; q < world; ++q) {
            int end = (q + 1 == world) ? n : (int)(std::sqrt((double)(q + 1) / world) * n / 32.0 + 0.5) * 32;
            end = std::max(end, prev + 32);
            if (q + 1 < world) end = std::min(end, n - 32 * (world - 1 - q));
            ctx->plan.own_end[q] = end;
            prev = end;
        }
        for (int q = world; q < 16; ++q) ctx->plan.own_end[q] = n;
    }
    ctx->plan.peer_mode = mode;
    return VPCA_OK;
}

int vpca_gram_gather(vpca_ctx* ctx) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->finalized) return fail(ctx, VPCA_ERR_STATE, "Gram already finalized");
    if (ctx->plan.num_peers < 2) return VPCA_OK;
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, gram_peer_barrier(ctx->plan, ctx->stream));          // every rank's contributions have landed
    ctx->st.kernel_launches += 1;
    if (ctx->plan.peer_mode == 1) {
        CUDA_OK(ctx, gram_gather_rows(ctx->plan, ctx->d_S, ctx->n, ctx->stream));
        CUDA_OK(ctx, gram_peer_barrier(ctx->plan, ctx->stream));      // nobody resets a Gram a peer is still reading
        ctx->st.kernel_launches += 2;
    }
    return VPCA_OK;
}

int vpca_peer_barrier(vpca_ctx* ctx) {
    if (ctx == nullptr) return fail(nullptr, VPCA_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->plan.num_peers < 2) return VPCA_OK;
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, gram_peer_barrier(ctx->plan, ctx->stream));
    ctx->st.kernel_launches += 1;
    return VPCA_OK;
}

int vpca_debug_gram_profile(vpca_ctx* ctx, int64_t* out, int32_t max_ctas) {
    if (ctx == nullptr || out == nullptr || max_ctas <= 0) return fail(ctx, VPCA_ERR_BAD_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_OK(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return gram_read_profile(ctx->plan, reinterpret_cast<long long*>(out), max_ctas);
}

}  // extern "C"
