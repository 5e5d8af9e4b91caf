// Synthetic cohort generator on the device (specification in DESIGN.md "Synthetic generator").
//
// Stands in for the retired Genomics API ingestion (reference: rdd/VariantsRDD.scala:187-236) and writes the
// encoded genotype matrix directly in the layout the Gram kernel consumes: dense, sample-major, int8 or bf16.
// Counter-based (Philox4x32-10 keyed by the seed, counter = (variant, sample pair)), so any tile is reproducible
// and the CPU oracle regenerates the same cells bit for bit.  All floating point uses the round-to-nearest
// intrinsics (__dmul_rn, __dadd_rn, __dsqrt_rn, __ddiv_rn) so that no FMA contraction can change a threshold.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "vpca_internal.h"

namespace vpca {
namespace {

constexpr uint32_t kTagVariant = 0xA11E1E00u;
constexpr uint32_t kTagCell = 0xC0FFEE00u;
constexpr int kNPop = 5;

struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

struct PopBounds {
    int b[kNPop];
};

// thresholds T_k = floor(p_k * 2^32) for variants [v0, v0 + nv)
__global__ void thresholds_kernel(uint64_t seed, int64_t v0, int64_t nv, uint32_t* __restrict__ thr) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nv) return;
    const uint64_t v = (uint64_t)(v0 + j);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const Philox4 a = philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), 0u, kTagVariant, k0, k1);
    const Philox4 b = philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), 1u, kTagVariant, k0, k1);
    const double u = __dmul_rn(__dadd_rn((double)a.x, 0.5), 1.0 / 4294967296.0);
    const double p = __dadd_rn(0.02, __dmul_rn(0.48, u));
    const double pq = __dmul_rn(p, __dadd_rn(1.0, -p));
    const double F[kNPop] = {0.15, 0.10, 0.07, 0.05, 0.03};
    const uint32_t w[kNPop] = {a.y, a.z, a.w, b.x, b.y};
#pragma unroll
    for (int k = 0; k < kNPop; ++k) {
        const int sum4 = (int)(w[k] & 255u) + (int)((w[k] >> 8) & 255u) + (int)((w[k] >> 16) & 255u) + (int)(w[k] >> 24);
        const double z = __ddiv_rn((double)(sum4 - 510), 147.80054127);
        double pk = __dadd_rn(p, __dmul_rn(__dsqrt_rn(__dmul_rn(F[k], pq)), z));
        if (pk < 0.001) pk = 0.001;
        if (pk > 0.999) pk = 0.999;
        thr[j * kNPop + k] = (uint32_t)__dmul_rn(pk, 4294967296.0);   // truncating cast = floor for pk > 0
    }
}

__device__ __forceinline__ int pop_of(int s, const PopBounds& pb) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kNPop - 1; ++i) k += (s >= pb.b[i]);
    return k;
}

// One thread = one sample pair (rows 2p, 2p+1) x 16 consecutive variants.  Lanes run along the variant axis so
// a warp writes 512 contiguous bytes (int8) of each of its two rows.
// Tile column jt = jbase + j of the cell; row-major (panel == 0): x[s * ld + jt]; panel layout: see vpca_internal.h.
__device__ __forceinline__ int64_t tile_index(int s, int64_t jt, int64_t ld, int64_t panel, int n) {
    if (panel == 0) return (int64_t)s * ld + jt;
    const int64_t pnl = jt / panel;
    return pnl * (int64_t)n * panel + (int64_t)s * panel + (jt - pnl * panel);
}

template <typename T>
__global__ void synth_dense_kernel(uint64_t seed, int n, int64_t v0, int64_t nv, int mode, const PopBounds pb,
                                   const uint32_t* __restrict__ thr, T* __restrict__ x, int64_t ld, int64_t panel,
                                   int64_t jbase) {
    const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 16-variant chunk
    const int pair = blockIdx.y * blockDim.y + threadIdx.y;
    const int64_t j0 = chunk * 16;
    const int s0 = pair * 2;
    if (j0 >= nv || s0 >= n) return;
    const bool has1 = (s0 + 1) < n;
    const int pop0 = pop_of(s0, pb), pop1 = pop_of(has1 ? s0 + 1 : s0, pb);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    T v0row[16], v1row[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int64_t j = j0 + i;
        int g0 = 0, g1 = 0;
        if (j < nv) {
            const uint64_t v = (uint64_t)(v0 + j);
            const Philox4 r = philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)pair, kTagCell, k0, k1);
            const uint32_t t0 = thr[j * kNPop + pop0], t1 = thr[j * kNPop + pop1];
            g0 = (int)(r.x < t0) + (int)(r.y < t0);
            g1 = (int)(r.z < t1) + (int)(r.w < t1);
            if (mode == 0) {
                g0 = g0 > 0;
                g1 = g1 > 0;
            }
        }
        if constexpr (sizeof(T) == 1) {
            v0row[i] = (T)g0;
            v1row[i] = (T)g1;
        } else {
            v0row[i] = __float2bfloat16_rn((float)g0);
            v1row[i] = __float2bfloat16_rn((float)g1);
        }
    }
    T* r0 = x + tile_index(s0, jbase + j0, ld, panel, n);
    T* r1 = x + tile_index(has1 ? s0 + 1 : s0, jbase + j0, ld, panel, n);
    if (j0 + 16 <= nv) {
        constexpr int nvec = 16 * (int)sizeof(T) / 16;
        const uint4* p0 = reinterpret_cast<const uint4*>(v0row);
        const uint4* p1 = reinterpret_cast<const uint4*>(v1row);
#pragma unroll
        for (int q = 0; q < nvec; ++q) {
            reinterpret_cast<uint4*>(r0)[q] = p0[q];
            if (has1) reinterpret_cast<uint4*>(r1)[q] = p1[q];
        }
    } else {
        for (int i = 0; i < 16 && j0 + i < nv; ++i) {
            r0[i] = v0row[i];
            if (has1) r1[i] = v1row[i];
        }
    }
}

// packed e2m1 variant: one thread = one sample pair x 16 consecutive variants = 8 bytes per row (cell j of a row in
// nibble j & 1 of byte j / 2; dosage m is the code 2 m).  Cells in [nv, round_up(nv, 128)) are written as zero.
__global__ void synth_e2m1_kernel(uint64_t seed, int n, int64_t v0, int64_t nv, int mode, const PopBounds pb,
                                  const uint32_t* __restrict__ thr, uint8_t* __restrict__ x, int64_t ld, int64_t panel,
                                  int64_t jbase) {
    const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = blockIdx.y * blockDim.y + threadIdx.y;
    const int64_t j0 = chunk * 16;
    const int s0 = pair * 2;
    const int64_t nv_pad = ((nv + 127) / 128) * 128;
    if (j0 >= nv_pad || s0 >= n) return;
    const bool has1 = (s0 + 1) < n;
    const int pop0 = pop_of(s0, pb), pop1 = pop_of(has1 ? s0 + 1 : s0, pb);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t w0[2] = {0u, 0u}, w1[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int64_t j = j0 + i;
        if (j < nv) {
            const uint64_t v = (uint64_t)(v0 + j);
            const Philox4 r = philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)pair, kTagCell, k0, k1);
            const uint32_t t0 = thr[j * kNPop + pop0], t1 = thr[j * kNPop + pop1];
            uint32_t g0 = (uint32_t)(r.x < t0) + (uint32_t)(r.y < t0);
            uint32_t g1 = (uint32_t)(r.z < t1) + (uint32_t)(r.w < t1);
            if (mode == 0) {
                g0 = g0 > 0;
                g1 = g1 > 0;
            }
            w0[i >> 3] |= (2u * g0) << (4 * (i & 7));
            w1[i >> 3] |= (2u * g1) << (4 * (i & 7));
        }
    }
    *reinterpret_cast<uint2*>(x + tile_index(s0, jbase + j0, ld, panel, n) / 2) = make_uint2(w0[0], w0[1]);
    if (has1) *reinterpret_cast<uint2*>(x + tile_index(s0 + 1, jbase + j0, ld, panel, n) / 2) = make_uint2(w1[0], w1[1]);
}

}  // namespace

cudaError_t synth_dense(uint64_t seed, int n, int64_t v0, int64_t nv, int mode, int elem_bits, void* d_x, int64_t ld,
                        int64_t panel, cudaStream_t stream) {
    if (nv <= 0 || n <= 0) return cudaSuccess;
    const int elem_bytes = elem_bits / 8;
    if (panel > 0) {
        if ((panel % 128) != 0) return cudaErrorInvalidValue;
        ld = panel;
        // zero the cells after nv in the last panel (the Gram kernel reads whole k-blocks of it)
        const int64_t npanels = (nv + panel - 1) / panel, tail = npanels * panel - nv;
        if (tail > 0 && elem_bits != 4) {
            const size_t bytes = (size_t)elem_bytes;
            cudaError_t em = cudaMemset2DAsync(static_cast<char*>(d_x) + ((size_t)(npanels - 1) * n * panel + (size_t)(panel - tail)) * bytes,
                                               (size_t)panel * bytes, 0, (size_t)tail * bytes, (size_t)n, stream);
            if (em != cudaSuccess) return em;
        } else if (tail > 0) {
            cudaError_t em = cudaMemsetAsync(static_cast<char*>(d_x) + (size_t)(npanels - 1) * n * panel / 2, 0,
                                             (size_t)n * panel / 2, stream);   // e2m1: clear the whole last panel first
            if (em != cudaSuccess) return em;
        }
    }
    if (elem_bits == 4) {
        if ((reinterpret_cast<uintptr_t>(d_x) & 31) != 0 || (ld % 128) != 0 || (panel == 0 && ld < ((nv + 127) / 128) * 128))
            return cudaErrorInvalidValue;
    } else if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0 || ((ld * elem_bytes) & 15) != 0) {
        return cudaErrorInvalidValue;
    }
    static const int cum[kNPop] = {26, 40, 60, 80, 100};
    PopBounds pb;
    for (int k = 0; k < kNPop; ++k) pb.b[k] = (int)(((int64_t)cum[k] * n) / 100);
    // thresholds are regenerated per slab of variants to bound the scratch buffer
    const int64_t slab = 1 << 22;
    uint32_t* d_thr = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&d_thr), (size_t)(nv < slab ? nv : slab) * kNPop * 4, stream);
    if (e != cudaSuccess) return e;
    for (int64_t b = 0; b < nv && e == cudaSuccess; b += slab) {
        const int64_t cnt = (nv - b) < slab ? (nv - b) : slab;
        thresholds_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(seed, v0 + b, cnt, d_thr);
        const int64_t chunks = (cnt + 15) / 16;
        const dim3 block(32, 8);
        const dim3 grid((unsigned)((chunks + 31) / 32), (unsigned)(((n + 1) / 2 + 7) / 8));
        if (elem_bits == 4) {
            const int64_t chunks4 = (((cnt + 127) / 128) * 128) / 16;
            const dim3 grid4((unsigned)((chunks4 + 31) / 32), grid.y);
            synth_e2m1_kernel<<<grid4, block, 0, stream>>>(seed, n, v0 + b, cnt, mode, pb, d_thr,
                                                           reinterpret_cast<uint8_t*>(d_x), ld, panel, b);
        } else if (elem_bytes == 1)
            synth_dense_kernel<int8_t><<<grid, block, 0, stream>>>(seed, n, v0 + b, cnt, mode, pb, d_thr,
                                                                   reinterpret_cast<int8_t*>(d_x), ld, panel, b);
        else
            synth_dense_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(seed, n, v0 + b, cnt, mode, pb, d_thr,
                                                                          reinterpret_cast<__nv_bfloat16*>(d_x), ld, panel, b);
        e = cudaGetLastError();
    }
    cudaFreeAsync(d_thr, stream);
    return e;
}

}  // namespace vpca
