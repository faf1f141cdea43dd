// Cost model of tcgen05.mma kind::tf32 (M=128, A from TMEM, B from smem): cycles per instruction for
// back-to-back issue into one accumulator / alternating accumulators, N = 128 vs 256, and SS mode.
#include <cstdio>
#include <vector>
#include "../tc_common.cuh"
using namespace tmpnn::tc;

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(128) timing_kernel(long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 131072 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_s);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_s;
    {   // defined A operand
        uint32_t z[32];
        for (int j = 0; j < 32; ++j) z[j] = __float_as_uint(0.5f) & 0xffffe000u;
        for (int c = 0; c < 8; ++c) tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c * 32, z);
        tmem_wait_st(); tc_fence_before();
    }
    __syncthreads();
    if (warp == 0 && lane == 0) {
        tc_fence_after();
        const uint32_t b = smem_u32(smem);
        uint32_t parity = 0;
        auto run = [&](int variant, int n_mma) -> long long {
            const uint32_t id128 = make_idesc_tf32(128, 128), id256 = make_idesc_tf32(128, 256);
            long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const uint32_t k = (i & 15) * 8;          // A column offset (K step)
                const uint64_t bd = make_desc_sw128(b + ((i & 3) * 32) + ((i >> 2) & 3) * 16384);
                switch (variant) {
                    case 0: mma_tf32_ts(tmem + 0, tmem + 256 + k, bd, id128, i != 0); break;                    // one accumulator, N=128
                    case 1: mma_tf32_ts(tmem + (i & 1) * 128, tmem + 256 + k, bd, id128, i > 1); break;         // two accumulators alternating
                    case 2: mma_tf32_ts(tmem + 0, tmem + 256 + k, bd, id256, i != 0); break;                    // one accumulator, N=256
                    case 3: mma_tf32_ss(tmem + 0, make_desc_sw128(b + 65536 + ((i & 3) * 32)), bd, id128, i != 0); break;  // SS, N=128
                    case 4: mma_tf32_ss(tmem + 0, make_desc_sw128(b + 65536 + ((i & 3) * 32)), bd, id256, i != 0); break;  // SS, N=256
                }
            }
            long long t1 = clock64();
            mma_commit(&bar);
            mbar_wait(&bar, parity); parity ^= 1;
            long long t2 = clock64();
            out[variant * 4 + 0] = t1 - t0; out[variant * 4 + 1] = t2 - t0; out[variant * 4 + 2] = n_mma;
            return t2 - t0;
        };
        for (int v = 0; v < 5; ++v) { run(v, 96); run(v, 96); }
        // commit/wait latency alone
        long long t0 = clock64();
        mma_tf32_ts(tmem, tmem + 256, make_desc_sw128(b), make_idesc_tf32(128, 128), 0);
        mma_commit(&bar); mbar_wait(&bar, parity); parity ^= 1;
        out[20] = clock64() - t0;
        // 8 batches of 6 with a commit+wait between batches (what the kernel's sub-chunking does, minus the epilogue)
        t0 = clock64();
        for (int bch = 0; bch < 8; ++bch) {
            for (int i = 0; i < 6; ++i) mma_tf32_ts(tmem, tmem + 256 + (i & 1) * 8 + bch * 16, make_desc_sw128(b + (i & 3) * 32), make_idesc_tf32(128, 128), (bch | i) != 0);
            tc_fence_before();
            asm volatile("" ::: "memory");
        }
        mma_commit(&bar); mbar_wait(&bar, parity); parity ^= 1;
        out[21] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
    long long* d; cudaMalloc(&d, 64 * sizeof(long long)); cudaMemset(d, 0, 64 * sizeof(long long));
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    timing_kernel<<<1, 128, 197 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[5] = {"TS N=128, one accumulator", "TS N=128, two accumulators alternating", "TS N=256, one accumulator", "SS N=128", "SS N=256"};
    for (int v = 0; v < 5; ++v)
        printf("%-42s issue %5lld cyc, to completion %6lld cyc for %lld MMAs -> %.1f cyc/MMA\n", nm[v], h[v * 4], h[v * 4 + 1], h[v * 4 + 2], (double)h[v * 4 + 1] / h[v * 4 + 2]);
    printf("single MMA + commit + wait: %lld cyc;  48 MMAs in 8 batches: %lld cyc\n", h[20], h[21]);
    return 0;
}
