// Hardware check of the tcgen05 building blocks in tc_common.cuh: one 128x128x128 GEMM, A from TMEM
// (written with tcgen05.st), B from pre-swizzled shared memory (bulk copy), 3xTF32 split, D read with tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_gemm_test tc_gemm_test.cu && ./tc_gemm_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tc_common.cuh"

using namespace tmpnn::tc;

constexpr int M = 128, N = 128, K = 128;
constexpr uint32_t D_COL = 0, AHI_COL = 128, ALO_COL = 256;

__global__ void __launch_bounds__(192) gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bimg /* hi | lo images */,
                                                   float* __restrict__ Dout, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* b_hi = smem;              // 64 KB
    uint8_t* b_lo = smem + 65536;      // 64 KB
    __shared__ uint64_t bar_b, bar_a, bar_d;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar_b, 1);
        mbar_init(&bar_a, 128);
        mbar_init(&bar_d, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 5 && lane == 0) {                       // producer: weights
        mbar_arrive_expect_tx(&bar_b, 131072);
        bulk_g2s(b_hi, Bimg, 65536, &bar_b);
        bulk_g2s(b_lo, Bimg + 16384, 65536, &bar_b);
    }
    if (warp < 4) {                                     // A rows -> TMEM (hi, lo)
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 4; ++c) {
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = A[row * K + c * 32 + j];
                hi[j] = tf32_hi(x);
                lo[j] = tf32_lo(x, hi[j]);
            }
            tmem_st32(lane_addr + AHI_COL + c * 32, hi);
            tmem_st32(lane_addr + ALO_COL + c * 32, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_a);
    }
    if (warp == 4 && lane == 0) {                       // MMA issuer
        mbar_wait(&bar_b, 0);
        mbar_wait(&bar_a, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(M, N);
        const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
        uint32_t acc = 0;
        for (int kb = 0; kb < 4; ++kb)
            for (int s = 0; s < 4; ++s) {
                const uint32_t koff = kb * 32 + s * 8;
                const uint64_t dh = make_desc_sw128(bh + kb * 16384 + s * 32), dl = make_desc_sw128(bl + kb * 16384 + s * 32);
                if (mode == 5) continue;   // corrections-first variant handled below
                mma_tf32_ts(tmem + D_COL, tmem + AHI_COL + koff, dh, idesc, acc);
                acc = 1;
                if (mode == 3 || mode == 4) {
                    mma_tf32_ts(tmem + D_COL, tmem + ALO_COL + koff, dh, idesc, 1);
                    mma_tf32_ts(tmem + D_COL, tmem + AHI_COL + koff, dl, idesc, 1);
                }
                if (mode == 4) mma_tf32_ts(tmem + D_COL, tmem + ALO_COL + koff, dl, idesc, 1);
            }
        if (mode == 5) {
            for (int pass = 0; pass < 2; ++pass)
                for (int kb = 0; kb < 4; ++kb)
                    for (int s = 0; s < 4; ++s) {
                        const uint32_t koff = kb * 32 + s * 8;
                        const uint64_t dh = make_desc_sw128(bh + kb * 16384 + s * 32), dl = make_desc_sw128(bl + kb * 16384 + s * 32);
                        if (pass == 0) {
                            mma_tf32_ts(tmem + D_COL, tmem + ALO_COL + koff, dh, idesc, acc);
                            acc = 1;
                            mma_tf32_ts(tmem + D_COL, tmem + AHI_COL + koff, dl, idesc, 1);
                        } else {
                            mma_tf32_ts(tmem + D_COL, tmem + AHI_COL + koff, dh, idesc, 1);
                        }
                    }
        }
        mma_commit(&bar_d);
    }
    if (warp < 4) {                                     // epilogue: D -> global
        mbar_wait(&bar_d, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(lane_addr + D_COL + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) Dout[row * N + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem);
}

__global__ void gelu_check_kernel(float* out, int n, float lo, float hi) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float x = lo + (hi - lo) * (float)i / (float)(n - 1); out[2 * i] = x; out[2 * i + 1] = gelu_fast(x); }
}
static int check_gelu() {
    const int n = 1 << 20;
    float* d; cudaMalloc(&d, 2 * n * sizeof(float));
    gelu_check_kernel<<<n / 256, 256>>>(d, n, -9.f, 9.f);
    std::vector<float> h(2 * n);
    cudaMemcpy(h.data(), d, 2 * n * sizeof(float), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxrel = 0;
    for (int i = 0; i < n; ++i) {
        double x = h[2 * i], ref = 0.5 * x * (1.0 + erf(x / sqrt(2.0)));
        double e = fabs(h[2 * i + 1] - ref);
        maxerr = fmax(maxerr, e);
        maxrel = fmax(maxrel, e / fmax(fabs(x), 1.0));
    }
    printf("gelu_fast on [-9,9]: max abs err %.3e, max err / max(1,|x|) %.3e\n", maxerr, maxrel);
    cudaFree(d);
    return (maxerr < 1.5e-6 && maxrel < 3e-7) ? 0 : 1;
}
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
int run(bool exact_inputs) {
    std::vector<float> A(M * K), W(N * K), D(M * N), img(2 * N * K);
    srand(1);
    for (auto& x : A) x = 4.f * ((float)rand() / RAND_MAX - 0.5f);
    for (auto& x : W) x = 0.5f * ((float)rand() / RAND_MAX - 0.5f);
    if (exact_inputs) { for (auto& x : A) x = trunc_tf32(x); for (auto& x : W) x = trunc_tf32(x); }
    printf("---- inputs %s\n", exact_inputs ? "exactly representable in tf32 (isolates accumulation error)" : "full fp32");
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            float x = W[n * K + k];
            uint32_t u; memcpy(&u, &x, 4);
            uint32_t h = u & 0xffffe000u;
            float hf; memcpy(&hf, &h, 4);
            float lf = x - hf;
            uint32_t off = sw128_offset(N, n, k) / 4;
            img[off] = hf;
            img[N * K + off] = lf;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, img.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
    int rc = 0;
    for (int mode : {1, 3, 4, 5}) {
        cudaMemset(dD, 0, D.size() * 4);
        gemm_kernel<<<1, 192, 131072, 0>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * (double)W[n * K + k];
                maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
                maxref = fmax(maxref, fabs(r));
            }
        // also the error of a plain fp32 FMA chain on the host, for scale
        double fmaerr = 0, rms = 0, bias = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0; float f = 0.f;
                for (int k = 0; k < K; ++k) { r += (double)A[m * K + k] * (double)W[n * K + k]; f = fmaf(A[m * K + k], W[n * K + k], f); }
                fmaerr = fmax(fmaerr, fabs(r - f));
                double e = D[m * N + n] - r; rms += e * e; bias += (r >= 0 ? e : -e);
            }
        const char* nm = mode == 1 ? "1xTF32" : mode == 3 ? "3xTF32" : mode == 4 ? "4xTF32 (+lo.lo)" : "3xTF32 corrections first";
        printf("mode %d %-26s max|D-ref| %.3e rel %.2e  rms %.2e  signed bias toward zero %.2e   [host fp32 fma chain max %.2e]\n", mode, nm, maxerr,
               maxerr / maxref, sqrt(rms / (M * N)), -bias / (M * N), fmaerr);
        if (!exact_inputs && mode == 3 && !(maxerr / maxref < 2e-6)) rc = 1;
        if (!exact_inputs && mode == 1 && !(maxerr / maxref < 5e-3)) rc = 1;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return rc;
}
int main() {
    int rc = run(false) | run(true) | check_gelu();
    printf(rc ? "FAIL\n" : "PASS\n");
    return rc;
}
