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

// Weight chunk of the per-residue FP32 / warp-level kernels: [32 k][128 n] fp32 = 16 KB (engine.cu: put_chunk).
constexpr int WCHUNK_K = 32;
constexpr int WCHUNK_FLOATS = WCHUNK_K * HID;   // 4096 floats = 16 KB

}  // namespace tmpnn
