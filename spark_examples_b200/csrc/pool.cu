// vpca_pool: one process, all GPUs of the box (include/vpca.h "one process, all GPUs").
//
// The reference's process model is one driver JVM whose `mapPartitions` tasks run concurrently and whose
// `reduceByKey(_ + _)` merges the per-partition matrices (VariantsPca.scala:38-50, :184-190).  The pool is that model
// on a multi-GPU host: one vpca_ctx per GPU, partition p served by GPU p % G, the contexts wired together with
// vpca_gram_set_peers_local so that every commit is added straight into the owners of its Gram rows over NVLink
// (fused reduce-scatter); vpca_pool_reduce_and_finalize is then only the closing barrier, the all-gather of the
// row bands and the symmetrize.  Built on the public C ABI only.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vpca.h"

namespace {
thread_local std::string tls_pool_error;
}

struct vpca_pool {
    std::vector<vpca_ctx*> ctx;
    int n = 0;
    int max_mult = 2;
    bool peers = false;            // false: no peer path between the devices -> host-staged reduce at the end
    bool reduced = false;
    std::mutex mu;
    std::unordered_map<int64_t, int64_t> staged;   // partition id -> variants staged (not yet committed)
    int64_t committed = 0, reserved = 0;           // variants across ALL GPUs: the int32 bound is on the summed Gram
};

namespace {

int pfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    tls_pool_error = buf;
    return code;
}

int from_ctx(vpca_ctx* c, int rc) {
    if (rc != VPCA_OK) tls_pool_error = vpca_last_error(c);
    return rc;
}

vpca_ctx* route(vpca_pool* pool, int64_t pid) {
    const int64_t g = (int64_t)pool->ctx.size();
    return pool->ctx[pid < 0 ? 0 : (size_t)(pid % g)];
}

// Every similarity count of the SUMMED matrix must stay a Java Int (VariantsPca.scala:185): reserve the variants of a
// batch against the bound over all GPUs before it is staged anywhere.
int reserve(vpca_pool* pool, int64_t pid, int64_t nv) {
    std::lock_guard<std::mutex> lk(pool->mu);
    if (pool->reduced) return pfail(VPCA_ERR_STATE, "Gram already reduced and finalized; call vpca_pool_reset first");
    const long double worst = (long double)(pool->committed + pool->reserved + nv) * pool->max_mult * pool->max_mult;
    if (worst > 2147483647.0L)
        return pfail(VPCA_ERR_OVERFLOW, "%lld variants x multiplicity %d^2 over all GPUs could overflow an int32 similarity count",
                     (long long)(pool->committed + pool->reserved + nv), pool->max_mult);
    pool->reserved += nv;
    if (pid >= 0) pool->staged[pid] += nv;
    return VPCA_OK;
}

// A failed batch poisons its partition (the context dropped the staging Gram): release everything it had reserved.
void settle(vpca_pool* pool, int64_t pid, int64_t nv, int rc) {
    std::lock_guard<std::mutex> lk(pool->mu);
    if (pid < 0) {
        pool->reserved -= nv;
        if (rc == VPCA_OK) pool->committed += nv;
        return;
    }
    if (rc != VPCA_OK) {
        auto it = pool->staged.find(pid);
        if (it != pool->staged.end()) {
            pool->reserved -= it->second;
            pool->staged.erase(it);
        }
    }
}

template <typename F>
int accumulate(vpca_pool* pool, int64_t pid, int64_t nv, F&& call) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    if (nv < 0) return pfail(VPCA_ERR_BAD_ARG, "nv < 0");
    int rc = reserve(pool, pid, nv);
    if (rc != VPCA_OK) return rc;
    vpca_ctx* c = route(pool, pid);
    rc = from_ctx(c, call(c));
    settle(pool, pid, nv, rc);
    return rc;
}

}  // namespace

extern "C" {

int vpca_pool_create(const vpca_config* cfg, int32_t n_gpus, const int32_t* devices, vpca_pool** out) {
    if (out == nullptr) return pfail(VPCA_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (cfg == nullptr || cfg->struct_size != sizeof(vpca_config))
        return pfail(VPCA_ERR_BAD_ARG, "cfg is NULL or struct_size != sizeof(vpca_config)");
    if (n_gpus < 1 || n_gpus > 16) return pfail(VPCA_ERR_BAD_ARG, "n_gpus must be in [1, 16]");
    vpca_pool* pool = new (std::nothrow) vpca_pool();
    if (pool == nullptr) return pfail(VPCA_ERR_NOMEM, "out of host memory");
    pool->n = cfg->n_samples;
    pool->max_mult = cfg->max_multiplicity > 0 ? cfg->max_multiplicity : 2;
    for (int g = 0; g < n_gpus; ++g) {
        vpca_config c = *cfg;
        c.device = devices != nullptr ? devices[g] : g;
        c.stream = nullptr;    // one private stream per context
        c.d_gram = nullptr;    // the peer wiring needs library-owned Grams
        c.gram_band_row0 = c.gram_band_rows = 0;
        vpca_ctx* ctx = nullptr;
        const int rc = vpca_create(&c, &ctx);
        if (rc != VPCA_OK) {
            tls_pool_error = vpca_last_error(nullptr);
            vpca_pool_destroy(pool);
            return rc;
        }
        pool->ctx.push_back(ctx);
    }
    if (n_gpus > 1) {
        int rc = vpca_gram_set_peers_local(pool->ctx.data(), n_gpus);
        if (rc == VPCA_OK) {
            const int mode = cfg->n_samples >= 64 * n_gpus ? VPCA_PEER_OWNER_ROWS : VPCA_PEER_REPLICATE;
            for (vpca_ctx* c : pool->ctx)
                if (rc == VPCA_OK) rc = from_ctx(c, vpca_gram_set_peer_mode(c, mode));
            pool->peers = rc == VPCA_OK;
        } else if (rc == VPCA_ERR_NCCL) {
            pool->peers = false;   // no peer path: keep per-GPU Grams and sum them through the host at the end
            rc = VPCA_OK;
        } else {
            tls_pool_error = vpca_last_error(nullptr);
        }
        if (rc != VPCA_OK) {
            vpca_pool_destroy(pool);
            return rc;
        }
    }
    for (vpca_ctx* c : pool->ctx) {   // the zeroed Grams are visible before any peer adds into them
        const int rc = from_ctx(c, vpca_synchronize(c));
        if (rc != VPCA_OK) {
            vpca_pool_destroy(pool);
            return rc;
        }
    }
    *out = pool;
    return VPCA_OK;
}

int vpca_pool_destroy(vpca_pool* pool) {
    if (pool == nullptr) return VPCA_OK;
    for (vpca_ctx* c : pool->ctx) vpca_synchronize(c);   // nobody frees a Gram a peer kernel may still write
    for (vpca_ctx* c : pool->ctx) vpca_destroy(c);
    delete pool;
    return VPCA_OK;
}

int32_t vpca_pool_size(const vpca_pool* pool) { return pool == nullptr ? 0 : (int32_t)pool->ctx.size(); }

vpca_ctx* vpca_pool_ctx(vpca_pool* pool, int64_t partition_id) { return pool == nullptr ? nullptr : route(pool, partition_id); }

const char* vpca_pool_last_error(const vpca_pool*) { return tls_pool_error.c_str(); }

int vpca_pool_reset(vpca_pool* pool) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    for (vpca_ctx* c : pool->ctx) {
        const int rc = from_ctx(c, vpca_reset(c));
        if (rc != VPCA_OK) return rc;
    }
    for (vpca_ctx* c : pool->ctx) {   // every Gram is zero before any peer may add into it
        const int rc = from_ctx(c, vpca_synchronize(c));
        if (rc != VPCA_OK) return rc;
    }
    std::lock_guard<std::mutex> lk(pool->mu);
    pool->staged.clear();
    pool->committed = pool->reserved = 0;
    pool->reduced = false;
    return VPCA_OK;
}

int vpca_pool_accumulate_calls(vpca_pool* pool, int64_t pid, const int64_t* offsets, const int32_t* sample_idx, int64_t nv) {
    return accumulate(pool, pid, nv, [&](vpca_ctx* c) { return vpca_accumulate_calls(c, pid, offsets, sample_idx, nv); });
}

int vpca_pool_accumulate_calls_u16(vpca_pool* pool, int64_t pid, const int64_t* offsets, const uint16_t* sample_idx,
                                   int64_t nv) {
    return accumulate(pool, pid, nv, [&](vpca_ctx* c) { return vpca_accumulate_calls_u16(c, pid, offsets, sample_idx, nv); });
}

int vpca_pool_accumulate_bits(vpca_pool* pool, int64_t pid, const uint8_t* bits, int64_t nv, int64_t stride_bytes) {
    return accumulate(pool, pid, nv, [&](vpca_ctx* c) { return vpca_accumulate_bits(c, pid, bits, nv, stride_bytes); });
}

int vpca_pool_accumulate_bed(vpca_pool* pool, int64_t pid, const uint8_t* rows, int64_t nv, int64_t stride_bytes,
                             int32_t counted_allele) {
    return accumulate(pool, pid, nv,
                      [&](vpca_ctx* c) { return vpca_accumulate_bed(c, pid, rows, nv, stride_bytes, counted_allele); });
}

int vpca_pool_commit(vpca_pool* pool, int64_t pid) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    vpca_ctx* c = route(pool, pid);
    const int rc = from_ctx(c, vpca_commit(c, pid));
    if (rc == VPCA_OK) {
        std::lock_guard<std::mutex> lk(pool->mu);
        auto it = pool->staged.find(pid);
        if (it != pool->staged.end()) {
            pool->reserved -= it->second;
            pool->committed += it->second;
            pool->staged.erase(it);
        }
    }
    return rc;
}

int vpca_pool_abort(vpca_pool* pool, int64_t pid) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    vpca_ctx* c = route(pool, pid);
    const int rc = from_ctx(c, vpca_abort(c, pid));
    if (rc == VPCA_OK) {
        std::lock_guard<std::mutex> lk(pool->mu);
        auto it = pool->staged.find(pid);
        if (it != pool->staged.end()) {
            pool->reserved -= it->second;
            pool->staged.erase(it);
        }
    }
    return rc;
}

int vpca_pool_reduce_and_finalize(vpca_pool* pool) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        if (pool->reduced) return VPCA_OK;
        if (!pool->staged.empty())
            return pfail(VPCA_ERR_STATE, "partition %lld is neither committed nor aborted", (long long)pool->staged.begin()->first);
    }
    const size_t G = pool->ctx.size();
    if (G > 1 && !pool->peers) {
        // no peer path between the devices: sum the per-GPU lower triangles through the host into GPU 0
        const size_t nn = (size_t)pool->n * pool->n;
        std::vector<int32_t> sum(nn), part(nn);
        int64_t variants = 0, v = 0;
        int rc = from_ctx(pool->ctx[0], vpca_get_partial_gram(pool->ctx[0], sum.data(), &variants));
        for (size_t g = 1; g < G && rc == VPCA_OK; ++g) {
            rc = from_ctx(pool->ctx[g], vpca_get_partial_gram(pool->ctx[g], part.data(), &v));
            for (size_t i = 0; i < nn; ++i) sum[i] += part[i];
            variants += v;
        }
        if (rc == VPCA_OK) rc = from_ctx(pool->ctx[0], vpca_load_partial_gram(pool->ctx[0], sum.data(), variants));
        if (rc == VPCA_OK) rc = from_ctx(pool->ctx[0], vpca_finalize_gram(pool->ctx[0]));
        if (rc != VPCA_OK) return rc;
    } else {
        // the commits are already in the owners' row bands: closing barrier + all-gather of the bands (enqueued on every
        // context's stream without blocking the host in between -- the barrier kernels wait for each other on the GPUs)
        for (vpca_ctx* c : pool->ctx) {
            const int rc = from_ctx(c, vpca_gram_gather(c));
            if (rc != VPCA_OK) return rc;
        }
        for (vpca_ctx* c : pool->ctx) {
            const int rc = from_ctx(c, vpca_finalize_gram(c));
            if (rc != VPCA_OK) return rc;
        }
        for (vpca_ctx* c : pool->ctx) {
            const int rc = from_ctx(c, vpca_synchronize(c));
            if (rc != VPCA_OK) return rc;
        }
    }
    std::lock_guard<std::mutex> lk(pool->mu);
    pool->reduced = true;
    return VPCA_OK;
}

int vpca_pool_get_gram(vpca_pool* pool, int32_t* out) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    return from_ctx(pool->ctx[0], vpca_get_gram(pool->ctx[0], out));
}

int vpca_pool_compute_pca(vpca_pool* pool, int32_t k, double* vecs, double* evals, int32_t* non_zero_rows) {
    if (pool == nullptr) return pfail(VPCA_ERR_BAD_ARG, "pool is NULL");
    return from_ctx(pool->ctx[0], vpca_compute_pca(pool->ctx[0], k, vecs, evals, non_zero_rows));
}

int vpca_pool_get_stats(vpca_pool* pool, vpca_stats* out) {
    if (pool == nullptr || out == nullptr) return pfail(VPCA_ERR_BAD_ARG, "NULL argument");
    vpca_stats sum{};
    for (size_t g = 0; g < pool->ctx.size(); ++g) {
        vpca_stats st{};
        const int rc = from_ctx(pool->ctx[g], vpca_get_stats(pool->ctx[g], &st));
        if (rc != VPCA_OK) return rc;
        sum.variants_accumulated += st.variants_accumulated;
        sum.gram_launches += st.gram_launches;
        sum.kernel_launches += st.kernel_launches;
        sum.h2d_bytes += st.h2d_bytes;
        sum.d2h_bytes += st.d2h_bytes;
        sum.last_gram_ms = st.last_gram_ms > sum.last_gram_ms ? st.last_gram_ms : sum.last_gram_ms;
        sum.last_eig_ms = st.last_eig_ms > sum.last_eig_ms ? st.last_eig_ms : sum.last_eig_ms;
        if (g == 0) {
            sum.gram_cta_group = st.gram_cta_group;
            sum.gram_resident = st.gram_resident;
            sum.eig_method = st.eig_method;
            sum.eig_iterations = st.eig_iterations;
        }
    }
    *out = sum;
    return VPCA_OK;
}

}  // extern "C"
