// Shared device helpers for the thermompnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "weights.h"

namespace tmpnn {


// exact-erf GELU, as torch.nn.GELU() (protein_mpnn_utils.py:813,856,888)
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16_ca(void* smem_dst, const void* gmem_src) {   // keeps the line in L1
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------
// Edge-tile geometry.  A tile is NPT whole residues = TM = 48*NPT edge rows by
// 128 output columns, computed by NPT*4 warps.  Warp w: wm = w/4 owns the 48
// rows of one residue, wn = w%4 owns 32 columns.  Lane: ty = lane/4 owns 6
// rows, tx = lane%4 owns columns {4tx..4tx+3} and {16+4tx..16+4tx+3} of the
// warp's 32.  So the neighbour sum of a residue is a shuffle over ty.
// ---------------------------------------------------------------------------
template <int NPT>
struct Tile {
    static constexpr int TM = KNB * NPT;
    static constexpr int THREADS = 128 * NPT;
};

struct ThreadPos {
    int wm, wn, ty, tx, m0, n0;
    __device__ __forceinline__ ThreadPos() {
        int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        wm = w >> 2; wn = w & 3; ty = lane >> 2; tx = lane & 3;
        m0 = wm * KNB + ty * 6;
        n0 = wn * 32 + tx * 4;
    }
    // column of accumulator index c (0..7)
    __device__ __forceinline__ int col(int c) const { return n0 + (c & 3) + ((c >> 2) << 4); }
};

// ---------------------------------------------------------------------------
// Weight stream: a kernel's matrices are stored back to back in HBM as
// [k][128] rows, so the whole sequence is one run of 32-row (16 KB) chunks
// that is pulled through a 3-slot shared-memory ring with cp.async, two chunks
// ahead of the FMA loop, across GEMM stage boundaries.
// ---------------------------------------------------------------------------
constexpr int WCHUNK_K = 32;
constexpr int WCHUNK_FLOATS = WCHUNK_K * HID;   // 4096 floats = 16 KB
constexpr int WRING = 3;

template <int THREADS>
struct WeightStream {
    const float* base;
    float* ring;
    int total;
    __device__ __forceinline__ void issue(int g) {
        if (g < total) {
            const float* src = base + (size_t)g * WCHUNK_FLOATS;
            float* dst = ring + (g % WRING) * WCHUNK_FLOATS;
            for (int i = threadIdx.x; i < WCHUNK_FLOATS / 4; i += THREADS) cp_async16(dst + i * 4, src + i * 4);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void start(const float* b, float* r, int t) {
        base = b; ring = r; total = t;
        issue(0);
        issue(1);
    }
    // Block-wide: returns chunk g (must be called with g = 0,1,2,... in order).
    // The barrier inside also publishes every shared-memory write made before the call.
    __device__ __forceinline__ const float* get(int g) {
        cp_async_wait<1>();
        __syncthreads();
        issue(g + 2);
        return ring + (g % WRING) * WCHUNK_FLOATS;
    }
};

// acc[6][8] += A[k][m0..m0+5] * W[k][cols], k over one 32-row chunk.  A is [k][TM] in shared memory.
template <int TM>
__device__ __forceinline__ void fma_chunk(float (&acc)[6][8], const float* __restrict__ A, const float* __restrict__ W,
                                          const ThreadPos& tp) {
    const float* ap = A + tp.m0;
    const float* wp = W + tp.n0;
#pragma unroll 4
    for (int k = 0; k < WCHUNK_K; ++k) {
        float2 a01 = *reinterpret_cast<const float2*>(ap + k * TM);
        float2 a23 = *reinterpret_cast<const float2*>(ap + k * TM + 2);
        float2 a45 = *reinterpret_cast<const float2*>(ap + k * TM + 4);
        float4 b0 = *reinterpret_cast<const float4*>(wp + k * HID);
        float4 b1 = *reinterpret_cast<const float4*>(wp + k * HID + 16);
        float a[6] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
        float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
}

// One full 128-deep GEMM stage: acc = A_s[128][TM] x (next 4 chunks of the stream).
template <int NPT>
__device__ __forceinline__ void gemm_stage128(float (&acc)[6][8], const float* A_s, WeightStream<Tile<NPT>::THREADS>& ws,
                                              int& g, const ThreadPos& tp) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    for (int c = 0; c < HID / WCHUNK_K; ++c) {
        const float* W = ws.get(g++);
        fma_chunk<Tile<NPT>::TM>(acc, A_s + c * WCHUNK_K * Tile<NPT>::TM, W, tp);
    }
}

// Store the thread's 6x8 values as the next stage's A operand: A_s[n][m].
template <int TM>
__device__ __forceinline__ void store_as_A(float* A_s, const float (&v)[6][8], const ThreadPos& tp) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float* p = A_s + tp.col(c) * TM + tp.m0;
        *reinterpret_cast<float2*>(p) = make_float2(v[0][c], v[1][c]);
        *reinterpret_cast<float2*>(p + 2) = make_float2(v[2][c], v[3][c]);
        *reinterpret_cast<float2*>(p + 4) = make_float2(v[4][c], v[5][c]);
    }
}

// Row-wise LayerNorm over the 128 columns of a tile held as 6x8 per thread.
// red: shared scratch [4][TM].  Two-pass (mean, then centred second moment).
// Contains block barriers; all threads must call it.
template <int TM>
__device__ __forceinline__ void tile_layernorm(float (&v)[6][8], const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float* red, const ThreadPos& tp) {
    float mean[6], rstd[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) s += v[r][c];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (tp.tx == 0) red[tp.wn * TM + tp.m0 + r] = s;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        int m = tp.m0 + r;
        mean[r] = (red[m] + red[TM + m] + red[2 * TM + m] + red[3 * TM + m]) * (1.0f / HID);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float d = v[r][c] - mean[r]; s = fmaf(d, d, s); }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (tp.tx == 0) red[tp.wn * TM + tp.m0 + r] = s;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        int m = tp.m0 + r;
        float var = (red[m] + red[TM + m] + red[2 * TM + m] + red[3 * TM + m]) * (1.0f / HID);
        rstd[r] = 1.0f / sqrtf(var + LN_EPS);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float g = __ldg(gamma + tp.col(c)), b = __ldg(beta + tp.col(c));
#pragma unroll
        for (int r = 0; r < 6; ++r) v[r][c] = (v[r][c] - mean[r]) * rstd[r] * g + b;
    }
}

}  // namespace tmpnn
