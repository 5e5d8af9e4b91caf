/*
 * abi_threads.c -- ONE process, G GPU contexts, T task threads, driven through the C ABI only (include/vpca.h).
 *
 * What it proves (SURVEY.md 8b "threading" / "process model"; reference: the bodies of `mapPartitions` run
 * concurrently, one task thread per core, and a failed task is retried -- VariantsPca.scala:184-190):
 *   - vpca_pool_accumulate_* / commit / abort are safe from many threads at once, on contexts that share a process;
 *   - a retried partition is counted exactly once (abort discards the staged rows; a batch with a bad sample index
 *     fails with VPCA_ERR_INDEX_OUT_OF_RANGE, poisons only its own partition, and the retry succeeds);
 *   - the fused same-process reduce (vpca_gram_set_peers_local, owner-rows) gives the oracle's matrix bit for bit,
 *     with G contexts on one device (1-GPU box) or on G devices.
 *
 *   abi_threads <contexts> <threads> <n_samples> <partitions> <variants_per_partition> [spread_devices=1] [seed]
 *
 * Built by __graft_entry__.build() into tests/_build/abi_threads; run by tests/test_abi_threads_gpu.py.
 * Test infrastructure: links the oracle (checker) next to libvpca.so (product).
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/vpca.h"

/* oracle/vpca_oracle.c */
int64_t vo_synth_calls(uint64_t seed, int32_t n, int64_t v0, int64_t nv, int64_t* off, int32_t* idx, int64_t* nv_out);
int vo_similarity(int32_t n, int64_t nv, const int64_t* off, const int32_t* idx, int32_t n_partitions, int32_t* S);
void vo_set_threads(int n);

typedef struct {
    int64_t nv;      /* rows (variants with at least one carrier) */
    int64_t* off;    /* nv + 1 */
    int32_t* idx;
} partition_t;

typedef struct {
    vpca_pool* pool;
    partition_t* parts;
    int nparts, n;
    atomic_int next;
    atomic_int failures, retries_abort, retries_badidx, wire16;
    unsigned seed;
} job_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* one attempt at a partition: rows go in as `nbatch` batches; mode 1 aborts half way (simulated task failure),
 * mode 2 corrupts one index of the middle batch (the reference would throw at VariantsPca.scala:59 / :188) */
static int attempt(job_t* job, int pid, int mode, int wire16) {
    partition_t* p = &job->parts[pid];
    const int nbatch = 3;
    for (int b = 0; b < nbatch; ++b) {
        const int64_t r0 = p->nv * b / nbatch, r1 = p->nv * (b + 1) / nbatch;
        const int64_t rows = r1 - r0;
        if (rows == 0) continue;
        if (mode == 1 && b == 1) {
            if (vpca_pool_abort(job->pool, pid) != VPCA_OK) return -100;
            return 1; /* retry */
        }
        /* batch-local CSR: offsets rebased to the slice */
        int64_t* off = (int64_t*)malloc((size_t)(rows + 1) * sizeof(int64_t));
        for (int64_t r = 0; r <= rows; ++r) off[r] = p->off[r0 + r] - p->off[r0];
        const int64_t nnz = off[rows];
        int rc;
        if (wire16) {
            uint16_t* ix = (uint16_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(uint16_t));
            for (int64_t e = 0; e < nnz; ++e) ix[e] = (uint16_t)p->idx[p->off[r0] + e];
            if (mode == 2 && b == 1 && nnz > 0) ix[nnz / 2] = (uint16_t)(job->n + 3);
            rc = vpca_pool_accumulate_calls_u16(job->pool, pid, off, ix, rows);
            free(ix);
        } else {
            int32_t* ix = (int32_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
            memcpy(ix, p->idx + p->off[r0], (size_t)nnz * sizeof(int32_t));
            if (mode == 2 && b == 1 && nnz > 0) ix[nnz / 2] = job->n + 3;
            rc = vpca_pool_accumulate_calls(job->pool, pid, off, ix, rows);
            free(ix);
        }
        free(off);
        if (mode == 2 && b == 1 && nnz > 0) {
            if (rc != VPCA_ERR_INDEX_OUT_OF_RANGE) {
                fprintf(stderr, "partition %d: corrupt batch returned %d (%s), expected INDEX_OUT_OF_RANGE\n", pid, rc,
                        vpca_pool_last_error(job->pool));
                return -101;
            }
            if (vpca_pool_abort(job->pool, pid) != VPCA_OK) return -102; /* already dropped by the library: a no-op */
            return 2;                                                     /* retry */
        }
        if (rc != VPCA_OK) {
            fprintf(stderr, "partition %d batch %d: %d %s\n", pid, b, rc, vpca_pool_last_error(job->pool));
            return -103;
        }
    }
    const int rc = vpca_pool_commit(job->pool, pid);
    if (rc != VPCA_OK) {
        fprintf(stderr, "partition %d commit: %d %s\n", pid, rc, vpca_pool_last_error(job->pool));
        return -104;
    }
    return 0;
}

static void* worker(void* arg) {
    job_t* job = (job_t*)arg;
    for (;;) {
        const int pid = atomic_fetch_add(&job->next, 1);
        if (pid >= job->nparts) break;
        /* deterministic per partition: every 3rd partition fails once by abort, every 5th once by a corrupt index */
        int mode = (pid % 3 == 1) ? 1 : ((pid % 5 == 2) ? 2 : 0);
        const int wire16 = (pid & 1) && job->n <= 65536;
        if (wire16) atomic_fetch_add(&job->wire16, 1);
        for (int tries = 0; tries < 3; ++tries) {
            const int r = attempt(job, pid, mode, wire16);
            if (r == 0) break;
            if (r < 0) {
                atomic_fetch_add(&job->failures, 1);
                break;
            }
            atomic_fetch_add(r == 1 ? &job->retries_abort : &job->retries_badidx, 1);
            mode = 0; /* the retry is clean */
        }
    }
    return NULL;
}

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s contexts threads n_samples partitions variants_per_partition [spread_devices=1] [seed]\n", argv[0]);
        return 2;
    }
    const int G = atoi(argv[1]), T = atoi(argv[2]), n = atoi(argv[3]), P = atoi(argv[4]);
    const int64_t vpp = atoll(argv[5]);
    const int spread = argc > 6 ? atoi(argv[6]) : 1;
    const uint64_t seed = argc > 7 ? strtoull(argv[7], NULL, 10) : 20240901ull;
    const int ndev = argc > 8 ? atoi(argv[8]) : 1;   /* devices to spread over (the caller knows the box) */

    /* inputs + expected matrix from the oracle (checker side) */
    vo_set_threads(8);
    partition_t* parts = (partition_t*)calloc((size_t)P, sizeof(partition_t));
    int64_t total_rows = 0, total_nnz = 0;
    for (int p = 0; p < P; ++p) {
        int64_t* off = (int64_t*)malloc((size_t)(vpp + 1) * sizeof(int64_t));
        int64_t kept = 0;
        const int64_t nnz = vo_synth_calls(seed, n, (int64_t)p * vpp, vpp, off, NULL, &kept);
        int32_t* idx = (int32_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
        vo_synth_calls(seed, n, (int64_t)p * vpp, vpp, off, idx, &kept);
        parts[p].nv = kept;
        parts[p].off = off;
        parts[p].idx = idx;
        total_rows += kept;
        total_nnz += nnz;
    }
    int64_t* all_off = (int64_t*)malloc((size_t)(total_rows + 1) * sizeof(int64_t));
    int32_t* all_idx = (int32_t*)malloc((size_t)(total_nnz > 0 ? total_nnz : 1) * sizeof(int32_t));
    int64_t r = 0, e = 0;
    all_off[0] = 0;
    for (int p = 0; p < P; ++p) {
        for (int64_t v = 0; v < parts[p].nv; ++v) all_off[++r] = e + parts[p].off[v + 1];
        memcpy(all_idx + e, parts[p].idx, (size_t)parts[p].off[parts[p].nv] * sizeof(int32_t));
        e += parts[p].off[parts[p].nv];
    }
    const size_t nn = (size_t)n * (size_t)n;
    int32_t* want = (int32_t*)malloc(nn * sizeof(int32_t));
    if (vo_similarity(n, total_rows, all_off, all_idx, 4, want) != 0) {
        fprintf(stderr, "oracle rejected its own input\n");
        return 3;
    }

    vpca_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.n_samples = n;
    cfg.dtype = VPCA_DTYPE_I8;
    cfg.num_pc = 2;
    cfg.max_multiplicity = 1;
    cfg.partitions_in_flight = T + 2;   /* every task thread may hold one partition open per GPU */
    cfg.staging_lanes = 2;
    cfg.chunk_variants = 8192;          /* small staging so that G x lanes fit beside each other on one device */
    cfg.chunk_nnz = 1 << 22;
    int32_t devices[16];
    for (int g = 0; g < G && g < 16; ++g) devices[g] = spread ? g % ndev : 0;
    vpca_pool* pool = NULL;
    int rc = vpca_pool_create(&cfg, G, devices, &pool);
    if (rc != VPCA_OK) {
        fprintf(stderr, "vpca_pool_create: %d %s\n", rc, vpca_pool_last_error(NULL));
        return 4;
    }

    int bad = 0;
    double secs = 0.0;
    int retries_abort = 0, retries_badidx = 0, wire16 = 0;
    for (int pass = 0; pass < 2 && !bad; ++pass) {   /* second pass: vpca_pool_reset and the same analysis again */
        if (pass > 0 && (rc = vpca_pool_reset(pool)) != VPCA_OK) {
            fprintf(stderr, "vpca_pool_reset: %d %s\n", rc, vpca_pool_last_error(pool));
            return 5;
        }
        job_t job;
        memset(&job, 0, sizeof(job));
        job.pool = pool;
        job.parts = parts;
        job.nparts = P;
        job.n = n;
        atomic_init(&job.next, 0);
        pthread_t* th = (pthread_t*)malloc((size_t)T * sizeof(pthread_t));
        const double t0 = now_s();
        for (int t = 0; t < T; ++t) pthread_create(&th[t], NULL, worker, &job);
        for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
        free(th);
        if (atomic_load(&job.failures) != 0) {
            fprintf(stderr, "pass %d: %d partitions failed\n", pass, atomic_load(&job.failures));
            return 6;
        }
        rc = vpca_pool_reduce_and_finalize(pool);
        if (rc != VPCA_OK) {
            fprintf(stderr, "vpca_pool_reduce_and_finalize: %d %s\n", rc, vpca_pool_last_error(pool));
            return 7;
        }
        secs = now_s() - t0;
        retries_abort = atomic_load(&job.retries_abort);
        retries_badidx = atomic_load(&job.retries_badidx);
        wire16 = atomic_load(&job.wire16);
        /* every context holds the whole matrix after the gather: check all of them, not only GPU 0 */
        int32_t* got = (int32_t*)malloc(nn * sizeof(int32_t));
        for (int g = 0; g < G; ++g) {
            rc = vpca_get_gram(vpca_pool_ctx(pool, g), got);
            if (rc != VPCA_OK) {
                fprintf(stderr, "vpca_get_gram(ctx %d): %d %s\n", g, rc, vpca_last_error(vpca_pool_ctx(pool, g)));
                return 8;
            }
            size_t diff = 0;
            for (size_t i = 0; i < nn; ++i) diff += got[i] != want[i];
            if (diff != 0) {
                fprintf(stderr, "pass %d ctx %d: %zu of %zu Gram entries differ from the oracle\n", pass, g, diff, nn);
                bad = 1;
            }
        }
        free(got);
    }
    double vecs[2 * 65536];
    double evals[2] = {0, 0};
    int32_t nz = 0;
    if (!bad && n <= 65535) {
        rc = vpca_pool_compute_pca(pool, 2, (double*)vecs, evals, &nz);
        if (rc != VPCA_OK) {
            fprintf(stderr, "vpca_pool_compute_pca: %d %s\n", rc, vpca_pool_last_error(pool));
            bad = 1;
        }
    }
    vpca_stats st;
    memset(&st, 0, sizeof(st));
    vpca_pool_get_stats(pool, &st);
    vpca_pool_destroy(pool);
    printf("{\"contexts\": %d, \"devices\": %d, \"threads\": %d, \"n_samples\": %d, \"partitions\": %d, \"rows\": %lld, "
           "\"retries_after_abort\": %d, \"retries_after_bad_index\": %d, \"partitions_on_uint16_wire\": %d, "
           "\"gram_bit_exact_vs_oracle\": %s, \"non_zero_rows\": %d, \"eval0\": %.17g, \"eval1\": %.17g, "
           "\"gram_launches\": %lld, \"h2d_bytes\": %lld, \"seconds_last_pass\": %.4f}\n",
           G, spread ? ndev : 1, T, n, P, (long long)total_rows, retries_abort, retries_badidx, wire16, bad ? "false" : "true",
           (int)nz, evals[0], evals[1], (long long)st.gram_launches, (long long)st.h2d_bytes, secs);
    return bad ? 1 : 0;
}
