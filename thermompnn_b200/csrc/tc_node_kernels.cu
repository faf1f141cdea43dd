// node_update for large batches on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
//   dh = (W3.Ssum + cnt*b3)/30 ; h1 = LN1(h_V + dh) ; h2 = mask * LN2(h1 + W_out gelu(W_in h1 + b_in) + b_out)
//   then up to four projections of h2 that the next edge kernels gather            protein_mpnn_utils.py:826-832, 871-880, 883-893
//
// One CTA works on tiles of 128 consecutive residues (thread = residue = TMEM lane, so both LayerNorms are per-thread
// reductions); 213 k MACs per residue = thirteen 128x128x128 GEMM blocks per tile, chained through TMEM:
//
//   A planes <- x = sum of the three 16-row partial neighbour sums            GEMM  D0  = x  . W3
//   A <- h1 = LN1(h_V + (D0 + cnt b3)/30)   (h1 also kept in shared memory)   GEMM  D1  = h1 . W_in[q]            q = 0..3:
//   A <- g_q = gelu(D1 + b_in[q])                                             GEMM  D0 (+)= g_q . W_out[:, q]
//   A <- h1 (from shared memory)                                              ... next q
//   A <- h2 = mask * LN2(h1 + D0 + b_out)   -> h_V                            GEMM  D0/D1 = h2 . W_proj[p]        p = 0..nproj-1
//
// TMEM (512 columns): D0 | D1 | A_hi | A_lo.  There is no room for a second A operand, so h1 and g_q alternate in the A
// planes; each rewrite is pipelined with the next GEMM at 16-column granularity (the MMA warp starts on a sub-chunk
// as soon as it is published), as in tc_edge_kernels.cu.  Weights stream through a 3-slot ring of 32 KB chunk images
// ([128 n][32 k] tf32-hi | tf32-lo, 128-byte swizzle), 52 chunks per tile with four projections.
//
// Warp roles (320 threads): warps 0-7 epilogue (two sets of four; set s owns the 16-column sub-chunks of parity s),
// warp 8 MMA issuer (owns TMEM), warp 9 weight producer.
// A residue's result does not depend on what else is in the batch: every (row, column) sees the same instruction
// sequence wherever its tile sits.  (The small-batch form, node_mma_kernels.cu, is a different arithmetic: the two agree
// to fp32 rounding, tests/test_gpu_parity.py.)
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "pdl.cuh"

namespace tmpnn {

using namespace tc;

namespace {

constexpr int TN_ROWS = 128;
constexpr int TN_THREADS = 320;
constexpr int TN_NW = 3;
constexpr uint32_t TN_HALF = 16384;          // one [128 n][32 k] tf32 plane
constexpr int TN_HP = HID + 4;               // row pitch of the h1 tile in shared memory (thread-per-row float4 access: conflict-free)

struct __align__(1024) TnSmem {
    uint8_t w[TN_NW][2 * TN_HALF];           // 96 KB weight ring
    float h1[TN_ROWS * TN_HP];               // 66 KB  LN1 output (FFN input and residual), fp32
    float b3[HID], n1g[HID], n1b[HID], bin[4 * HID], bout[HID], n2g[HID], n2b[HID], pb[4][HID];
    float2 lnx[2][TN_ROWS];                  // LayerNorm partial (mean, M2) exchange between the two epilogue sets
    uint64_t w_full[TN_NW], w_empty[TN_NW];
    uint64_t a_ready[8];                     // A sub-chunk sc written by one set's 128 threads
    uint64_t d_ready[2];                     // a GEMM into D0 / D1 has completed
    uint64_t d_free[2];                      // projections: the epilogue has finished reading D0 / D1
    uint32_t tmem_base;
};

__device__ __forceinline__ void load_row16(float (&v)[16], const float* p) {      // 16 consecutive floats, 16-byte aligned
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row16(float* p, const float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(p + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

}  // namespace

// dev tool (TMPNN_TC_TRACE=1 in a -DTMPNN_DEV build): clock64() stamps of block 0's first tile; epilogue set 0 uses slots
// 0..31 (warp 0) and the MMA warp slots 32..63
#define TN_STAMP(slot_) do { if (a.trace && blockIdx.x == 0 && lane == 0 && tile == 0) a.trace[slot_] = clock64(); } while (0)

__global__ void __launch_bounds__(TN_THREADS, 1) tc_node_update_kernel(const __grid_constant__ NodeUpdArgs a, int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    TnSmem& S = *reinterpret_cast<TnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, nproj = a.nproj;
    const float* wb = a.weights;
    const int chunks_per_tile = 4 + 32 + 4 * nproj;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TN_NW; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&S.a_ready[i], 128);
        for (int i = 0; i < 2; ++i) { mbar_init(&S.d_ready[i], 1); mbar_init(&S.d_free[i], 256); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < HID; i += TN_THREADS) {
        S.b3[i] = wb[a.w.b3 + i]; S.n1g[i] = wb[a.w.n1.g + i]; S.n1b[i] = wb[a.w.n1.b + i];
        S.bout[i] = wb[a.w.bout + i]; S.n2g[i] = wb[a.w.n2.g + i]; S.n2b[i] = wb[a.w.n2.b + i];
        for (int p = 0; p < 4; ++p) S.pb[p][i] = (p < nproj && a.proj[p].bias >= 0) ? wb[a.proj[p].bias + i] : 0.f;
    }
    for (int i = threadIdx.x; i < 4 * HID; i += TN_THREADS) S.bin[i] = wb[a.w.bin + i];
    if (warp == 8) tmem_alloc<512>(&S.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    pdl_trigger();                                        // everything above touched only constants, shared memory and TMEM
    if (warp != 9) pdl_wait();

    if (warp == 9) {
        // ---------------- weight producer: the tile's chunk images, in consumption order ----------------
        if (lane == 0) {
            const uint8_t* img = reinterpret_cast<const uint8_t*>(wb + a.tc_img);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int c = 0; c < chunks_per_tile; ++c, ++it) {
                    const uint32_t s = it % TN_NW, u = it / TN_NW;
                    if (u > 0) mbar_wait(&S.w_empty[s], (u - 1) & 1);
                    mbar_arrive_expect_tx(&S.w_full[s], 2 * TN_HALF);
                    bulk_g2s(S.w[s], img + (size_t)c * 2 * TN_HALF, 2 * TN_HALF, &S.w_full[s]);
                }
        }
    } else if (warp == 8) {
        // ---------------- MMA issuer (whole warp runs the loop with warp-uniform operands; one elected lane issues) ----------
        const uint32_t idesc = make_idesc_tf32(128, 128);
        const uint32_t tm = __shfl_sync(0xffffffffu, S.tmem_base, 0);
        uint32_t it = 0;            // weight chunk counter
        uint32_t agen = 0;          // A-operand generation (one per rewrite of the A planes)
        uint32_t tcount = 0;        // tiles done by this CTA (parity of d_free)
        // one 128x128x128 GEMM block: D[dcol] (+)= A . W(next 4 chunks); wait_a: the A planes were rewritten for it
        int tile = blockIdx.x, gcount = 0;
        auto gemm = [&](uint32_t dcol, bool init, bool wait_a) {
            const uint32_t d = tm + dcol;
            if (gcount < 15) TN_STAMP(32 + 2 * gcount);
#pragma unroll 1
            for (int c = 0; c < 4; ++c, ++it) {
                const uint32_t sw = it % TN_NW;
                const uint32_t bh = desc_sw128_lo(smem_u32(S.w[sw]));
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int sc = 2 * c + half;
                    if (wait_a) mbar_wait(&S.a_ready[sc], agen & 1);
                    if (half == 0) mbar_wait(&S.w_full[sw], (it / TN_NW) & 1);
                    tc_fence_after();
                    const uint32_t ahi = tm + COL_AHI + sc * 16, alo = tm + COL_ALO + sc * 16;
                    const bool first = init && c == 0 && half == 0;
                    if (elect_one()) {          // per sub-chunk: the two small correction products first, then hi.hi
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
            if (wait_a) ++agen;
            if (gcount < 15) TN_STAMP(33 + 2 * gcount);
            ++gcount;
        };
        for (; tile < num_tiles; tile += gridDim.x, ++tcount) {
            gemm(COL_D0, true, true);                                   // x . W3
            for (int q = 0; q < 4; ++q) {
                gemm(COL_D1, true, true);                               // h1 . W_in[q]
                gemm(COL_D0, q == 0, true);                             // g_q . W_out[:, q]   (accumulates over q)
            }
            for (int p = 0; p < nproj; ++p) {
                if (p >= 2) mbar_wait(&S.d_free[p & 1], tcount & 1);    // the epilogue has read projection p - 2
                gemm((p & 1) ? COL_D1 : COL_D0, true, p == 0);          // h2 . W_proj[p]: A is written once for all of them
            }
        }
    } else {
        // ---------------- epilogue: thread <-> (residue row, sub-chunk parity) ----------------
        const int set = warp >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;                    // 0..127 = TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        uint32_t nd[2] = {0, 0};                           // completed GEMMs into D0 / D1 seen so far (parity of d_ready)
        auto wait_d = [&](int b) {
            mbar_wait(&S.d_ready[b], nd[b] & 1);
            ++nd[b];
            tc_fence_after();
        };
        float* h1row = S.h1 + r * TN_HP;
        const int et = set * 128 + r;                      // 0..255 among the epilogue threads
        // Global traffic goes through the shared-memory tile with COALESCED 16-byte accesses (consecutive threads ->
        // consecutive 16 bytes of a row).  Thread-per-row global access made each load / store instruction touch 32
        // different lines: the x load alone took 27 k cycles, the whole tile 109 k (clock64 trace), 63 k with this.
        auto tile_rows = [&](int tile) { const int left = N - tile * TN_ROWS; return left < TN_ROWS ? left : TN_ROWS; };
        auto coop_load = [&](const float* src, int nrows) {        // [nrows][128] global -> S.h1 rows (rows past the end: zeros)
#pragma unroll 4
            for (int i = et; i < TN_ROWS * 32; i += 256) {
                const int row = i >> 5, c4 = (i & 31) * 4;
                const float4 v = row < nrows ? *reinterpret_cast<const float4*>(src + (size_t)row * HID + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(S.h1 + row * TN_HP + c4) = v;
            }
        };
        auto coop_store = [&](float* dst, int nrows) {             // S.h1 rows -> [nrows][128] global
#pragma unroll 4
            for (int i = et; i < TN_ROWS * 32; i += 256) {
                const int row = i >> 5, c4 = (i & 31) * 4;
                if (row < nrows) *reinterpret_cast<float4*>(dst + (size_t)row * HID + c4) = *reinterpret_cast<const float4*>(S.h1 + row * TN_HP + c4);
            }
        };
#define TN_ESTAMP(slot_) do { if (warp == 0) TN_STAMP(slot_); } while (0)
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int node = tile * TN_ROWS + r;
            const bool valid = node < N;
            const int nrows = tile_rows(tile);
            const size_t row0 = (size_t)tile * TN_ROWS;
            TN_ESTAMP(0);
            // ---- A <- x: the three partial neighbour sums of each residue, added in order
            {
                const float* sp = a.Spart + row0 * 3 * HID;
#pragma unroll 2
                for (int i = et; i < TN_ROWS * 32; i += 256) {
                    const int row = i >> 5, c4 = (i & 31) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < nrows) {
                        const float* q = sp + (size_t)row * 3 * HID + c4;
                        const float4 p0 = *reinterpret_cast<const float4*>(q), p1 = *reinterpret_cast<const float4*>(q + HID),
                                     p2 = *reinterpret_cast<const float4*>(q + 2 * HID);
                        v = make_float4((p0.x + p1.x) + p2.x, (p0.y + p1.y) + p2.y, (p0.z + p1.z) + p2.z, (p0.w + p1.w) + p2.w);
                    }
                    *reinterpret_cast<float4*>(S.h1 + row * TN_HP + c4) = v;
                }
                named_bar(3, 256);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int sc = set + 2 * c;
                    float v[16];
                    load_row16(v, h1row + sc * 16);
                    publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                }
                named_bar(3, 256);                          // every row has been read: the tile buffer is free
            }
            TN_ESTAMP(1);
            // ---- h1 = LN1(h_V + (D0 + cnt b3) / 30)      (the h_V tile is fetched while x . W3 runs)
            float mean, rstd;
            {
                const float cnt = valid ? a.cnt_pre[node] : 0.f;
                coop_load(a.hV + row0 * HID, nrows);
                named_bar(3, 256);
                wait_d(0);
                TN_ESTAMP(2);
                float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int sc = set + 2 * c;
                    uint32_t acc[16];
                    tmem_ld16(lane_addr + COL_D0 + sc * 16, acc);
                    float hvv[16], bv[16], v[16];
                    load_row16(hvv, h1row + sc * 16);
                    load_vec16(bv, &S.b3[sc * 16]);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = hvv[j] + fmaf(cnt, bv[j], __uint_as_float(acc[j])) / MSG_SCALE;
                        if (c == 0 && j == 0) shift = v[0];
                        const float dlt = v[j] - shift;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                    }
                    store_row16(h1row + sc * 16, v);
                }
                const float mean_a = shift + s1 * (1.0f / 64.0f);
                S.lnx[set][r] = make_float2(mean_a, fmaxf(s2 - s1 * s1 * (1.0f / 64.0f), 0.f));
                named_bar(3, 256);
                const float2 o0 = S.lnx[0][r], o1 = S.lnx[1][r];
                const float dm = o0.x - o1.x;
                mean = 0.5f * (o0.x + o1.x);
                rstd = 1.0f / sqrtf((o0.y + o1.y + 32.0f * dm * dm) * (1.0f / HID) + LN_EPS);   // Chan et al. combine
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {                  // normalise -> h1 (kept in shared memory) -> A
                const int sc = set + 2 * c;
                float v[16], gv[16], bv[16];
                load_row16(v, h1row + sc * 16);
                load_vec16(gv, &S.n1g[sc * 16]); load_vec16(bv, &S.n1b[sc * 16]);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = (v[j] - mean) * rstd * gv[j] + bv[j];
                store_row16(h1row + sc * 16, v);
                publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
            }
            TN_ESTAMP(3);
            // ---- FFN: four hidden blocks
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                wait_d(1);                                  // h1 . W_in[q] done: D1 holds the block, A is free
                TN_ESTAMP(4 + 4 * q);
                {
                    uint32_t accb[2][16];
                    tmem_ld16(lane_addr + COL_D1 + set * 16, accb[0]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int sc = set + 2 * c;
                        uint32_t (&acc)[16] = accb[c & 1];
                        float bv[16], v[16];
                        load_vec16(bv, &S.bin[q * HID + sc * 16]);
                        tmem_wait_ld();
                        if (c + 1 < 4) tmem_ld16(lane_addr + COL_D1 + (sc + 2) * 16, accb[(c + 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            upk2(gelu_fast2(add2(pk2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1])), pk2(bv[j], bv[j + 1]))), v[j], v[j + 1]);
                        publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                    }
                }
                TN_ESTAMP(5 + 4 * q);
                wait_d(0);                                  // g_q . W_out[:, q] done: A is free again
                TN_ESTAMP(6 + 4 * q);
                if (q < 3) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {           // A <- h1 for the next hidden block
                        const int sc = set + 2 * c;
                        float v[16];
                        load_row16(v, h1row + sc * 16);
                        publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                    }
                }
                TN_ESTAMP(7 + 4 * q);
            }
            // ---- h2 = mask * LN2(h1 + D0 + b_out)  -> h_V (and the decoder-state copy), A for the projections
            {
                float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int sc = set + 2 * c;
                    uint32_t acc[16];
                    tmem_ld16(lane_addr + COL_D0 + sc * 16, acc);
                    float hv[16], bv[16], v[16];
                    load_row16(hv, h1row + sc * 16);
                    load_vec16(bv, &S.bout[sc * 16]);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = hv[j] + (__uint_as_float(acc[j]) + bv[j]);
                        if (c == 0 && j == 0) shift = v[0];
                        const float dlt = v[j] - shift;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                    }
                    store_row16(h1row + sc * 16, v);
                }
                const float mean_a = shift + s1 * (1.0f / 64.0f);
                named_bar(3, 256);                          // (both sets are past their reads of lnx from LN1)
                S.lnx[set][r] = make_float2(mean_a, fmaxf(s2 - s1 * s1 * (1.0f / 64.0f), 0.f));
                named_bar(3, 256);
                const float2 o0 = S.lnx[0][r], o1 = S.lnx[1][r];
                const float dm = o0.x - o1.x;
                mean = 0.5f * (o0.x + o1.x);
                rstd = 1.0f / sqrtf((o0.y + o1.y + 32.0f * dm * dm) * (1.0f / HID) + LN_EPS);
            }
            {
                const float m = valid ? a.maskf[node] : 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int sc = set + 2 * c;
                    float v[16], gv[16], bv[16];
                    load_row16(v, h1row + sc * 16);
                    load_vec16(gv, &S.n2g[sc * 16]); load_vec16(bv, &S.n2b[sc * 16]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = ((v[j] - mean) * rstd * gv[j] + bv[j]) * m;
                    store_row16(h1row + sc * 16, v);
                    if (nproj > 0) publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                }
                named_bar(3, 256);
                coop_store(a.hV + row0 * HID, nrows);       // (runs under the first projection GEMM)
                if (a.hV_copy) coop_store(a.hV_copy + row0 * HID, nrows);
            }
            TN_ESTAMP(20);
            // ---- projections: D0 / D1 alternate; the MMA warp runs one projection ahead of this read-out
            for (int p = 0; p < nproj; ++p) {
                const int b = p & 1;
                const ProjSpec ps = a.proj[p];
                const float* tab = (ps.table >= 0 && valid) ? wb + ps.table + (size_t)a.S[node] * HID : nullptr;
                named_bar(3, 256);                          // the tile buffer is free (previous coalesced store has read it)
                wait_d(b);
                TN_ESTAMP(21 + 2 * p);
                uint32_t accb[2][16];
                tmem_ld16(lane_addr + (b ? COL_D1 : COL_D0) + set * 16, accb[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int sc = set + 2 * c;
                    uint32_t (&acc)[16] = accb[c & 1];
                    float bv[16], v[16];
                    load_vec16(bv, &S.pb[p][sc * 16]);
                    tmem_wait_ld();
                    if (c + 1 < 4) tmem_ld16(lane_addr + (b ? COL_D1 : COL_D0) + (sc + 2) * 16, accb[(c + 1) & 1]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) + bv[j];
                    if (tab) {
                        float tv[16];
                        load_row16(tv, tab + sc * 16);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += tv[j];
                    }
                    store_row16(h1row + sc * 16, v);
                }
                if (p + 2 < nproj) {                        // lets the MMA warp overwrite this accumulator with projection p + 2
                    tc_fence_before();
                    mbar_arrive(&S.d_free[b]);
                }
                named_bar(3, 256);
                coop_store(ps.out + row0 * HID, nrows);
                TN_ESTAMP(22 + 2 * p);
            }
            named_bar(3, 256);                              // the next tile's x load overwrites the tile buffer
            tc_fence_before();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

cudaError_t tc_node_kernels_init() {
    return cudaFuncSetAttribute(tc_node_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TnSmem) + 1024);
}

void launch_tc_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st) {
    const int num_tiles = (a.N + TN_ROWS - 1) / TN_ROWS;
    const int grid = num_tiles < sms ? num_tiles : sms;
    launch_pdl(tc_node_update_kernel, dim3(grid), dim3(TN_THREADS), sizeof(TnSmem) + 1024, st, a, num_tiles);
}

}  // namespace tmpnn
