// The transfer head for large batches on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
//   x = [dec3 | dec2 | W_s[S]] (384)  ->  y = LightAttention centre tap (384 x 384) + b      transfer_model.py:82-83, 96-108, 140-155
//   z1 = Lin(384, 64)(relu y) ; z2 = Lin(64, 32)(relu z1) ; z3 = Lin(32, 21)(relu z2) ; T = w z3 + b    :67-73, 110-111
//
// One CTA works on tiles of 128 consecutive residues (thread = residue = TMEM lane).  The chain, in 128-wide K blocks:
//
//   for j = 0..2 (output block of y):   for k = 0..2:  A <- x_k ;  D0 (+)= x_k . LA[j,k]          9 blocks, N = 128
//                                       A <- relu(D0 + b_j) = relu(y_j) ;  D1 (+)= relu(y_j) . W1[:, j]   3 blocks, N = 64
//   A <- relu(D1 + b1) (64 wide) ; D0 = . W2 (N = 32) ;  A <- relu(D0 + b2) (32 wide) ; D1 = . W3 (N = 32) ;  T = w (D1 + b3) + b
//
// TMEM (512 columns): D0 | D1 | A_hi | A_lo as in the other chained-GEMM kernels; the 384-wide x tile does not fit
// beside the weight ring in shared memory, so its three blocks are (re)fetched block by block into ONE staging tile with
// coalesced loads (they are L2-resident: 3 x 64 KB per tile, read three times), each fetch issued while the previous
// block's GEMM runs.  Weights: 51 chunk images of 32 KB per tile through a 3-slot ring.
// Warp roles (320 threads): warps 0-7 epilogue (two sets of four; set s owns the 16-column sub-chunks of parity s),
// warp 8 MMA issuer (owns TMEM), warp 9 weight producer.  Same batch-invariance property as tc_node_kernels.cu.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "pdl.cuh"

namespace tmpnn {

using namespace tc;

namespace {

constexpr int TH_ROWS = 128;
constexpr int TH_THREADS = 320;
constexpr int TH_NW = 3;
constexpr uint32_t TH_HALF = 16384;
constexpr int TH_P = HID + 4;                 // row pitch of the staging tile
constexpr int TH_CHUNKS = 36 + 12 + 2 + 1;    // LA | both_out.1 | both_out.3 | both_out.5

struct __align__(1024) ThSmem {
    uint8_t w[TH_NW][2 * TH_HALF];            // 96 KB weight ring
    float xt[TH_ROWS * TH_P];                 // 66 KB staging tile: one 128-wide block of x; finally the [128][21] output
    float la_b[3 * HID], m1_b[64], m2_b[32], m3_b[32];
    uint64_t w_full[TH_NW], w_empty[TH_NW];
    uint64_t a_ready[8];
    uint64_t d_ready[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void ld16(float (&v)[16], const float* p) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}

}  // namespace

__global__ void __launch_bounds__(TH_THREADS, 1) tc_head_kernel(const __grid_constant__ HeadArgs a, int64_t tc_img, int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    ThSmem& S = *reinterpret_cast<ThSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N;
    const float* wb = a.weights;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TH_NW; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&S.a_ready[i], 128);
        for (int i = 0; i < 2; ++i) mbar_init(&S.d_ready[i], 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 3 * HID; i += TH_THREADS) S.la_b[i] = wb[a.la_b + i];
    if (threadIdx.x < 64) S.m1_b[threadIdx.x] = wb[a.m1_b + threadIdx.x];
    if (threadIdx.x < 32) { S.m2_b[threadIdx.x] = wb[a.m2_b + threadIdx.x]; S.m3_b[threadIdx.x] = threadIdx.x < 21 ? wb[a.m3_b + threadIdx.x] : 0.f; }
    if (warp == 8) tmem_alloc<512>(&S.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    pdl_trigger();
    if (warp != 9) pdl_wait();

    if (warp == 9) {
        if (lane == 0) {                                  // weight producer
            const uint8_t* img = reinterpret_cast<const uint8_t*>(wb + tc_img);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int c = 0; c < TH_CHUNKS; ++c, ++it) {
                    const uint32_t s = it % TH_NW, u = it / TH_NW;
                    if (u > 0) mbar_wait(&S.w_empty[s], (u - 1) & 1);
                    mbar_arrive_expect_tx(&S.w_full[s], 2 * TH_HALF);
                    bulk_g2s(S.w[s], img + (size_t)c * 2 * TH_HALF, 2 * TH_HALF, &S.w_full[s]);
                }
        }
    } else if (warp == 8) {
        // ---------------- MMA issuer ----------------
        const uint32_t tm = __shfl_sync(0xffffffffu, S.tmem_base, 0);
        uint32_t it = 0, apar = 0;                        // apar: bit sc = parity of a_ready[sc]'s next completion
        // D[dcol] (+)= A[:, 0 : 32 nchunks] . W(next nchunks chunks)^T, output width n_out (a multiple of 16)
        auto gemm = [&](uint32_t dcol, bool init, int nchunks, uint32_t idesc) {
            const uint32_t d = tm + dcol;
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c, ++it) {
                const uint32_t sw = it % TH_NW;
                const uint32_t bh = desc_sw128_lo(smem_u32(S.w[sw]));
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int sc = 2 * c + half;
                    mbar_wait(&S.a_ready[sc], (apar >> sc) & 1);
                    apar ^= 1u << sc;
                    if (half == 0) mbar_wait(&S.w_full[sw], (it / TH_NW) & 1);
                    tc_fence_after();
                    const uint32_t ahi = tm + COL_AHI + sc * 16, alo = tm + COL_ALO + sc * 16;
                    const bool first = init && c == 0 && half == 0;
                    if (elect_one()) {
                        if (first) mma_tf32_ts_lh_init(d, alo, bh, DESC_SW128_HI, idesc);
                        else mma_tf32_ts_lh(d, alo, bh + half * 4, DESC_SW128_HI, idesc);
                        mma_tf32_ts_lh(d, ahi, bh + 1024 + half * 4, DESC_SW128_HI, idesc);
                        mma_tf32_ts_lh(d, alo + 8, bh + half * 4 + 2, DESC_SW128_HI, idesc);
                        mma_tf32_ts_lh(d, ahi + 8, bh + 1024 + half * 4 + 2, DESC_SW128_HI, idesc);
                        mma_tf32_ts_lh(d, ahi, bh + half * 4, DESC_SW128_HI, idesc);
                        mma_tf32_ts_lh(d, ahi + 8, bh + half * 4 + 2, DESC_SW128_HI, idesc);
                    }
                    __syncwarp();
                }
                if (elect_one()) mma_commit(&S.w_empty[sw]);
            }
            if (elect_one()) mma_commit(&S.d_ready[dcol == COL_D0 ? 0 : 1]);
        };
        const uint32_t i128 = make_idesc_tf32(128, 128), i64 = make_idesc_tf32(128, 64), i32 = make_idesc_tf32(128, 32);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int j = 0; j < 3; ++j) {
                for (int k = 0; k < 3; ++k) gemm(COL_D0, k == 0, 4, i128);     // x_k . LA[j,k]
                gemm(COL_D1, j == 0, 4, i64);                                  // relu(y_j) . W1[:, j]
            }
            gemm(COL_D0, true, 2, i32);                                        // relu(z1) . W2      (K = 64)
            gemm(COL_D1, true, 1, i32);                                        // relu(z2) . W3      (K = 32)
        }
    } else {
        // ---------------- epilogue ----------------
        const int set = warp >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const int et = set * 128 + r;
        uint32_t nd[2] = {0, 0};
        auto wait_d = [&](int b) {
            mbar_wait(&S.d_ready[b], nd[b] & 1);
            ++nd[b];
            tc_fence_after();
        };
        float* xrow = S.xt + r * TH_P;
        const float sc_w = __ldg(wb + a.ddg), sc_b = __ldg(wb + a.ddg + 1);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const size_t row0 = (size_t)tile * TH_ROWS;
            const int left = N - tile * TH_ROWS, nrows = left < TH_ROWS ? left : TH_ROWS;
            // block k of x for this tile -> staging (coalesced; rows past the end: zeros)
            auto fetch = [&](int k) {
#pragma unroll 4
                for (int i = et; i < TH_ROWS * 32; i += 256) {
                    const int row = i >> 5, c4 = (i & 31) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < nrows) {
                        const float* src = k == 0 ? a.dec_last + (row0 + row) * HID : k == 1 ? a.dec_prev + (row0 + row) * HID
                                                  : wb + a.Ws + (size_t)a.S[row0 + row] * HID;
                        v = *reinterpret_cast<const float4*>(src + c4);
                    }
                    *reinterpret_cast<float4*>(S.xt + row * TH_P + c4) = v;
                }
            };
            auto publish_staged = [&]() {                  // staging rows -> A (all eight sub-chunks)
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int sc = set + 2 * c;
                    float v[16];
                    ld16(v, xrow + sc * 16);
                    publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                }
            };
            fetch(0);
            named_bar(3, 256);
            for (int j = 0; j < 3; ++j) {
                for (int k = 0; k < 3; ++k) {
                    // A is free here: the GEMM that read it last has been waited for (below), or this is the tile's first block
                    publish_staged();
                    named_bar(3, 256);                      // staging tile read by every row
                    if (!(j == 2 && k == 2)) fetch((k + 1) % 3);            // next block's fetch runs under this block's GEMM
                    named_bar(3, 256);
                    wait_d(0);                              // x_k . LA[j,k] done
                }
                {   // A <- relu(y_j) = relu(D0 + b_j)
                    uint32_t accb[2][16];
                    tmem_ld16(lane_addr + COL_D0 + set * 16, accb[0]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int sc = set + 2 * c;
                        uint32_t (&acc)[16] = accb[c & 1];
                        float bv[16], v[16];
                        load_vec16(bv, &S.la_b[j * HID + sc * 16]);
                        tmem_wait_ld();
                        if (c + 1 < 4) tmem_ld16(lane_addr + COL_D0 + (sc + 2) * 16, accb[(c + 1) & 1]);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(acc[i]) + bv[i], 0.f);
                        publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                    }
                }
                wait_d(1);                                  // relu(y_j) . W1[:, j] done: A is free for the next x block
            }
            // ---- z1 (64 wide): sub-chunks 0..3, two per set
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int sc = set + 2 * c;
                uint32_t acc[16];
                tmem_ld16(lane_addr + COL_D1 + sc * 16, acc);
                float bv[16], v[16];
                load_vec16(bv, &S.m1_b[sc * 16]);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(acc[i]) + bv[i], 0.f);
                publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
            }
            wait_d(0);
            {   // ---- z2 (32 wide): sub-chunk `set`
                uint32_t acc[16];
                tmem_ld16(lane_addr + COL_D0 + set * 16, acc);
                float bv[16], v[16];
                load_vec16(bv, &S.m2_b[set * 16]);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(acc[i]) + bv[i], 0.f);
                publish_A16(lane_addr, set, v, &S.a_ready[set]);
            }
            wait_d(1);
            {   // ---- T = w (z3 + b3) + b: 21 columns -> staging [128][21] -> coalesced store
                uint32_t acc[16];
                tmem_ld16(lane_addr + COL_D1 + set * 16, acc);
                float bv[16];
                load_vec16(bv, &S.m3_b[set * 16]);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int col = set * 16 + i;
                    if (col < 21) S.xt[r * 21 + col] = fmaf(__uint_as_float(acc[i]) + bv[i], sc_w, sc_b);   // ddg_out: Linear(1,1)
                }
                named_bar(3, 256);
                float* dst = a.table + row0 * 21;
                for (int i = et; i < nrows * 21; i += 256) dst[i] = S.xt[i];
                named_bar(3, 256);                          // the next tile's fetch overwrites the staging tile
            }
            tc_fence_before();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

cudaError_t tc_head_kernels_init() {
    return cudaFuncSetAttribute(tc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ThSmem) + 1024);
}

void launch_tc_head(const HeadArgs& a, int64_t tc_img, int sms, cudaStream_t st) {
    const int num_tiles = (a.N + TH_ROWS - 1) / TH_ROWS;
    const int grid = num_tiles < sms ? num_tiles : sms;
    launch_pdl(tc_head_kernel, dim3(grid), dim3(TH_THREADS), sizeof(ThSmem) + 1024, st, a, tc_img, num_tiles);
}

}  // namespace tmpnn
