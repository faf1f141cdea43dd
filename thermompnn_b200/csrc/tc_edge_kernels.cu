// Tensor-core (tcgen05) generation of the per-edge kernels.
//
// One CTA works on tiles of 128 consecutive edge rows (rows of the [N*48][128] h_E array; tiles ignore residue
// boundaries).  Warp roles of tc_edge_kernel (448 threads; tc_edge_feat_kernel is laid out the same way with its own
// producers):
//   warps 0-7    epilogue, two sets of four: thread t of warp w owns edge row 32(w%4)+t = TMEM lane.  It turns each
//                16-column accumulator sub-chunk into the next GEMM's A operand (bias / gathered projections, GELU,
//                hi/lo split, tcgen05.st); set s takes the sub-chunks of parity s
//   warp 8       one elected lane issues tcgen05.mma (kind::tf32, A from TMEM, B from shared memory); owns TMEM
//   warp 9       lane 0 streams weight chunks (pre-split, pre-swizzled 32 KB images) through a 3-slot ring (bulk copy)
//   warps 10-13  activation producers: h_E chunks by TMA tensor-map copies, neighbours' Q rows by cp.async gathers,
//                per-tile metadata (edge weights, masks, P rows)
//
// TMEM (512 columns): D0 | D1 | A_hi | A_lo, 128 columns each.  The accumulator is double buffered so that
// stage l+1's MMAs (fed chunk by chunk as the epilogue of stage l produces A) run underneath that epilogue.
// fp32 accuracy comes from the 3xTF32 split (tc_common.cuh).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "pdl.cuh"

namespace tmpnn {

using namespace tc;

namespace {

constexpr int TC_ROWS = 128;
constexpr int TC_WSLOTS = 3;
constexpr uint32_t TC_WCHUNK_BYTES = 32768;   // [128 n][32 k] hi (16 KB) | lo (16 KB)

__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Write 32 values as columns [32c, 32c+32) of the next A operand and publish the chunk.  hi plane: the tf32 the tensor
// core would read anyway (truncation); lo plane: the exact remainder (packed: fma(hi, -1, v)), as in publish_A16.
__device__ __forceinline__ void publish_A_chunk(uint32_t lane_addr, int c, const float (&v)[32], uint64_t* a_ready) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        hi[j] = __float_as_uint(v[j]) & 0xffffe000u;
        hi[j + 1] = __float_as_uint(v[j + 1]) & 0xffffe000u;
        float l0, l1;
        upk2(fma2(pk2(__uint_as_float(hi[j]), __uint_as_float(hi[j + 1])), pk2(-1.0f, -1.0f), pk2(v[j], v[j + 1])), l0, l1);
        lo[j] = __float_as_uint(l0);
        lo[j + 1] = __float_as_uint(l1);
    }
    tmem_st32(lane_addr + COL_AHI + c * 32, hi);
    tmem_st32(lane_addr + COL_ALO + c * 32, lo);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(a_ready);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// One kernel, three kinds (same pipeline, different epilogues):
//   UPD  h_E <- LN3(h_E + W13 gelu(W12 gelu(W11b.h_E + P_i + Q_j) + b12) + b13)       protein_mpnn_utils.py:834-838
//        weight image W11b | W12 | W13
//   ENC  g = gelu(W2 gelu(W1b.h_E + P_i + Q_j) + b2); per-residue sums of m_i m_j g     :819-825   image W1b | W2
//   DEC  g = gelu(W2 gelu(P_i + m_i (W1b.h_E + R_j)) + b2); per-residue sums of g       :863-870, 1268-1272
// ENC/DEC write the column sums of every 16-row sub-block of the edge array (a residue is exactly three of them)
// into Spart[residue][0..2][128]; node_update adds the three parts in order, so the neighbour sum is deterministic
// and independent of where the residue sits in the batch.  W3 is applied there, after the sum.
//
// Warp roles: see the top of the file.
// Everything gathered per row (h_E input, Q_j, the residual) flows through ONE ring of [128 rows][32 cols] chunks.
// MMA schedule: K-chunk by K-chunk as the epilogue publishes them (short tail after the last chunk); inside a chunk
// the two small correction products (A_lo.B_hi, A_hi.B_lo) go first, then A_hi.B_hi.
// ---------------------------------------------------------------------------------------------------------
enum TcKind { TC_UPD = 0, TC_ENC = 1, TC_DEC = 2 };

constexpr int T2_THREADS = 352;
constexpr int T2_NW = 3, T2_NACT = 4;       // (a fourth weight slot was measured: no faster, the weight stream is not the limiter)
constexpr int ACT_PITCH = 36;          // floats per row of an activation chunk: 128 B of data + 16 B pad, so that thread-per-row
                                       // 16-byte reads hit 8 different bank groups per quarter-warp (9r mod 8 = r mod 8)
constexpr uint32_t T2_HALF = 16384;   // one [128 n][32 k] tf32 plane

struct __align__(1024) Tc2Smem {
    uint8_t w[T2_NW][2 * T2_HALF];       // 96 KB  weight chunks: B_hi plane | B_lo plane
    float out[2][TC_ROWS * 32];          // 32 KB  UPD: output staging, one slab per epilogue set
    float act[T2_NACT][TC_ROWS * ACT_PITCH];   // 72 KB  gathered activation chunks, row pitch 144 B
    float p[2][4][HID];                  // P rows of the residues the tile touches, double-buffered by tile parity
    float wrow[2][TC_ROWS];
    float mrow[2][TC_ROWS];
    float vec[4][HID];
    float2 lnx[4][TC_ROWS];              // LayerNorm partial (mean, M2) exchange between the epilogue sets
    int jidx[TC_ROWS];
    uint64_t w_full[T2_NW], w_empty[T2_NW];
    uint64_t act_full[T2_NACT], act_empty[T2_NACT];
    uint64_t a_ready[8], d_ready[2];     // a_ready: one per 16-column sub-chunk of the A operand
    uint64_t meta_full[2];               // wrow / mrow / p of a tile (by tile parity) are in place
    uint32_t tmem_base;
};

struct TcArgs {
    alignas(64) CUtensorMap tm_hE;   // h_E as a [N*48][128] fp32 tensor, box [128 rows][32 cols], 128-byte swizzle
    float* hE;              // [N*48][128]; UPD updates it in place
    const int* E_idx;       // [N*48]
    const float* P; int p_stride;   // per-residue projection of h_V[i] (+bias); stride 0 = one shared row
    const float* Q; int q_stride;   // per-residue projection of h_V[j] (R_j for DEC)
    const float* maskf;     // [N]
    const uint8_t* wimg;    // weight image: per matrix 4 chunks of (hi 16 KB | lo 16 KB)
    const float* v0;        // UPD: b12   ENC/DEC: b2
    const float* v1;        // UPD: b13
    const float* v2;        // UPD: LN gain
    const float* v3;        // UPD: LN shift
    float* Spart;           // ENC/DEC: [N][3][128]
    long long* trace;       // optional (TMPNN_TC_TRACE=1): clock64() phase stamps of block 0, 32 slots per tile
    int N, num_tiles;
};

// phase stamp: slot layout per tile = [0..15] epilogue set 0 (warp 0), [16..23] set 1 (warp 4), [24..31] MMA warp
#define TC_STAMP(ti_, slot_) \
    do { if (a.trace && blockIdx.x == 0 && lane == 0 && (ti_) < 16) a.trace[(ti_) * 32 + (slot_)] = clock64(); } while (0)

// Column sums over each 16-row half of a warp (half-warps never straddle a residue: 48 = 3 x 16).
// On return lane l holds, for its half, the sums of columns 2*(l&15) and 2*(l&15)+1 in v[0], v[1].
__device__ __forceinline__ void halfwarp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 8, cnt = 16; s >= 1; s >>= 1, cnt >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
            const float send = up ? v[i] : v[i + cnt];
            const float keep = up ? v[i + cnt] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
}

// this thread's 16 floats (columns 16*half ..) of row r of an activation chunk
__device__ __forceinline__ void load_act16(float (&v)[16], const float* chunk, int r, int half) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t = *reinterpret_cast<const float4*>(chunk + r * ACT_PITCH + half * 16 + j * 4);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
// the same for a chunk written by TMA: dense 128-byte rows, 16-byte units XOR (row & 7)
__device__ __forceinline__ void load_act16_sw(float (&v)[16], const float* chunk, int r, int half) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t = *reinterpret_cast<const float4*>(chunk + r * 32 + (((half * 4 + j) ^ (r & 7)) << 2));
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
// NSETS epilogue sets of four warps: set s handles the 16-column sub-chunks sc with sc % NSETS == s.
// (Four sets -- 16 epilogue warps, two sub-chunks per thread and stage -- measured 8 % SLOWER than two: the epilogue is
// bound by its instruction count, not by per-thread latency, and four sets publish their sub-chunks in two bursts,
// which the MMA warp follows less closely than the finer interleave of two sets.)
template <int KIND, int NSETS, int NAP>
__global__ void __launch_bounds__((4 * NSETS + 2 + NAP) * 32, 1) tc_edge_kernel(const __grid_constant__ TcArgs a) {
    constexpr int STAGES = (KIND == TC_UPD) ? 3 : 2;
    constexpr int W_MMA = 4 * NSETS, W_WPROD = 4 * NSETS + 1, W_APROD = 4 * NSETS + 2;
    constexpr int NTHREADS = (4 * NSETS + 2 + NAP) * 32, EPI_THREADS = 128 * NSETS;
    constexpr int RPL = 4 / NAP;                         // tile rows per activation-producer lane
    constexpr int SPT = 8 / NSETS;                       // sub-chunks per thread and stage
    extern __shared__ uint8_t smem_raw[];
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[13] = clock64();   // kernel entry
    Tc2Smem& S = *reinterpret_cast<Tc2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, num_tiles = a.num_tiles;
    const int total_rows = N * KNB;

    if (threadIdx.x == 0) {
        for (int i = 0; i < T2_NW; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (int i = 0; i < T2_NACT; ++i) { mbar_init(&S.act_full[i], 64 * NAP); mbar_init(&S.act_empty[i], 256); }
        for (int i = 0; i < 2; ++i) mbar_init(&S.meta_full[i], 64 * NAP);
        for (int i = 0; i < 8; ++i) mbar_init(&S.a_ready[i], 128);
        for (int i = 0; i < 2; ++i) mbar_init(&S.d_ready[i], 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < HID; i += NTHREADS) {
        S.vec[0][i] = a.v0[i];
        if (KIND == TC_UPD) { S.vec[1][i] = a.v1[i]; S.vec[2][i] = a.v2[i]; S.vec[3][i] = a.v3[i]; }
    }
    if (warp == W_MMA) tmem_alloc<512>(&S.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    // PDL: everything above touched only constants, shared memory and TMEM.  The weight producer keeps to constants and
    // starts streaming at once; every other warp first waits for the kernels that wrote h_E / P / Q / E_idx.
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[14] = clock64();   // setup done
    pdl_trigger();
    if (warp != W_WPROD) pdl_wait();
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[15] = clock64();   // predecessor complete

    if (warp == W_WPROD) {
        // ---------------- weight producer ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int c = 0; c < 4 * STAGES; ++c, ++it) {
                    const uint32_t s = it % T2_NW, u = it / T2_NW;
                    if (u > 0) mbar_wait(&S.w_empty[s], (u - 1) & 1);
                    mbar_arrive_expect_tx(&S.w_full[s], 2 * T2_HALF);
                    bulk_g2s(S.w[s], a.wimg + (size_t)c * 2 * T2_HALF, 2 * T2_HALF, &S.w_full[s]);
                }
        }
    } else if (warp == W_MMA) {
        // ---------------- MMA issuer: one 16-column sub-chunk (2 K-steps) at a time, as the epilogue publishes them ----
        // The whole warp runs the loop so that every operand is warp-uniform (lives in uniform registers); only the
        // tcgen05 instructions themselves sit under `if (elect_one())`.  Computing operands inside an `if (lane == 0)` region
        // makes the compiler wrap each UTCHMMA in a ~10-instruction uniform-register "waterfall".
        {
            const uint32_t idesc = make_idesc_tf32(128, 128);
            const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);   // a value the compiler KNOWS is warp-uniform
            uint32_t it = 0, g = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
#pragma unroll 1
                for (int l = 0; l < STAGES; ++l, ++g) {
                    const uint32_t d = tmem + ((g & 1) ? COL_D1 : COL_D0);
#pragma unroll 1
                    for (int c = 0; c < 4; ++c, ++it) {
                        const uint32_t sw = it % T2_NW;
                        const uint32_t bh = desc_sw128_lo(smem_u32(S.w[sw]));      // low descriptor word of the B_hi plane
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int sc = 2 * c + half;
                            mbar_wait(&S.a_ready[sc], g & 1);
                            if (sc == 0) TC_STAMP(g / STAGES, 24 + 2 * l);
                            if (half == 0) mbar_wait(&S.w_full[sw], (it / T2_NW) & 1);
                            tc_fence_after();
                            const uint32_t ahi = tmem + COL_AHI + sc * 16, alo = tmem + COL_ALO + sc * 16;
                            const bool first = (c == 0) && (half == 0);
                            // per 16-column sub-chunk: the two small correction products first, then hi.hi; descriptor
                            // offsets are compile-time (16-byte units: K-step 32 B -> +2, second half +4, lo plane +1024)
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
                    if (elect_one()) mma_commit(&S.d_ready[g & 1]);
                    TC_STAMP(g / STAGES, 25 + 2 * l);
                }
        }
    } else if (warp >= W_APROD) {
        // ---------------- activation producers (NAP warps) ----------------
        // Every lane owns RPL rows of a tile (producer warp pw: rows 32 (pw RPL + i) + lane) and copies their 128-byte
        // chunk slices with eight 16-byte cp.async each (h_E rows, or the neighbour's Q row): no index math inside the
        // copy loop.  One producer warp sustains only ~1 chunk per 2.2k cycles (each cp.async touches 32 different
        // lines), which bounded the whole kernel; NAP warps issue in parallel.
        const int pw = warp - W_APROD;
        // (One 128-byte bulk copy per row was measured 3x slower: too fine-grained for the TMA unit.)
        // Chunk order = the order the epilogue consumes them: in(0) | q(t), [residual(t)], in(t+1) for t = 0, 1, ...
        uint32_t k = 0;
        int jrow[RPL];                                       // neighbour of the row this lane owns (row 32 (pw RPL + i) + lane)
        int jq[RPL][8];                                      // neighbour of row 32 (pw RPL + i) + 4 t + lane / 8 (gather pattern)
        auto spread_j = [&]() {
#pragma unroll
            for (int i = 0; i < RPL; ++i)
#pragma unroll
                for (int t = 0; t < 8; ++t) jq[i][t] = __shfl_sync(0xffffffffu, jrow[i], 4 * t + (lane >> 3));
        };
        auto chunk = [&](int row0, int c, bool gather_q, int with_p_buf) {
            const uint32_t s = k % T2_NACT, u = k / T2_NACT;
            if (u > 0) mbar_wait(&S.act_empty[s], (u - 1) & 1);
            const bool tma_lane = !gather_q && pw == 0 && lane == 0;
            if (gather_q) {
                // eight lanes per 128-byte row slice, four rows per instruction: 4 L1 wavefronts per cp.async instead of
                // the 32 of a lane-per-row pattern (the LSU data pipe was the busiest unit of the kernel, ncu: 65-70 %)
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    float* dst = &S.act[s][(32 * (pw * RPL + i) + (lane >> 3)) * ACT_PITCH + (lane & 7) * 4];
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        cp_async16_ca(dst + 4 * t * ACT_PITCH, a.Q + (size_t)jq[i][t] * a.q_stride + c * 32 + (lane & 7) * 4);
                }
            } else if (tma_lane) {
                // h_E rows of a tile are contiguous: ONE tensor-map copy per chunk (rows past the end arrive as zeros)
                mbar_arrive_expect_tx(&S.act_full[s], TC_ROWS * 32 * 4);
                tma_load_2d(S.act[s], &a.tm_hE, c * 32, row0, &S.act_full[s]);
            }
            cp_async_arrive_noinc(&S.act_full[s]);          // fires when this lane's copies have landed
            if (!tma_lane) mbar_arrive(&S.act_full[s]);
            ++k;
        };
        // metadata of a tile (weights / masks per row) + its four input chunks; returns the neighbour rows in jnext
        auto tile_input = [&](int tile, uint32_t mb, int (&jnext)[RPL]) {
            const int row0 = tile * TC_ROWS;
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = lane + 32 * (pw * RPL + i), row = row0 + r;
                int j = N - 1;
                float w = 0.f, mi = 0.f;
                if (row < total_rows) {
                    const int node = row / KNB;
                    mi = a.maskf[node];
                    j = a.E_idx[row];
                    if (j < 0) j = node;                                   // unused slot: a finite dummy
                    else w = (KIND == TC_ENC) ? mi * a.maskf[j] : 1.0f;    // mask_attend (protein_mpnn_utils.py:1232-1233)
                }
                jnext[i] = j; S.wrow[mb][r] = w; S.mrow[mb][r] = mi;
            }
            {   // P rows of the (up to 4) residues the tile touches; then publish the tile's metadata on its own barrier
                const int node_lo = row0 / KNB;
#pragma unroll
                for (int i = pw * RPL; i < (pw + 1) * RPL; ++i) {
                    int node = node_lo + i; node = node < N ? node : N - 1;
                    cp_async16_ca(&S.p[mb][i][lane * 4], a.P + (size_t)node * a.p_stride + lane * 4);
                }
                cp_async_arrive_noinc(&S.meta_full[mb]);    // P rows landed
                mbar_arrive(&S.meta_full[mb]);              // wrow / mrow stores released
            }
            for (int c = 0; c < 4; ++c) chunk(row0, c, false, -1);
        };
        uint32_t ti = 0;
        int tile = blockIdx.x;
        if (tile < num_tiles) tile_input(tile, 0, jrow);
        for (; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int row0 = tile * TC_ROWS;
            spread_j();
            for (int c = 0; c < 4; ++c) chunk(row0, c, true, -1);
            if (KIND == TC_UPD)                              // residual BEFORE the next tile's input: it waits in the ring for
                for (int c = 0; c < 4; ++c) chunk(row0, c, false, -1);   // the LayerNorm pass, which then frees the slots
            int jn[RPL] = {};
            const int next = tile + gridDim.x;
            if (next < num_tiles) tile_input(next, (ti + 1) & 1, jn);
#pragma unroll
            for (int i = 0; i < RPL; ++i) jrow[i] = jn[i];
        }
        cp_async_wait<0>();
    } else {
        // ---------------- epilogue: thread <-> (edge row, sub-chunk parity) ----------------
        const int set = warp >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;                    // 0..127
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        uint32_t ti = 0, g = 0, kc = 0;                    // tile counter, MMA stage counter, activation chunk counter
        auto act_wait = [&](uint32_t k) -> const float* {
            const uint32_t s = k % T2_NACT;
            mbar_wait(&S.act_full[s], (k / T2_NACT) & 1);
            return S.act[s];
        };
        // the slot's next writer may be the TMA unit (async proxy): our ld.shared must be ordered before it
        // (not needed after reading an input chunk: the slot's next writer is the cp.async gather of q)
        auto act_release = [&](uint32_t k, bool next_is_tma = true) {
            if (next_is_tma) fence_proxy_async();
            mbar_arrive(&S.act_empty[k % T2_NACT]);
        };
        // input stage of a tile: its four h_E chunks -> A (each thread converts its 16-column half of every chunk)
        auto input_stage = [&]() {
            for (int c = 0; c < SPT; ++c) {
                const int sc = set + NSETS * c;
                const float* chunk = act_wait(kc + (sc >> 1));
                float v[16];
                load_act16_sw(v, chunk, r, sc & 1);
                act_release(kc + (sc >> 1), false);
                publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
            }
            kc += 4;
        };
        int tile = blockIdx.x;
        const int sb0 = (quad == 0) ? (set ? 16 : 0) : 64;   // only warps 0 and 4 stamp
#define EPI_STAMP(n_) do { if (sb0 < 32) TC_STAMP(ti, sb0 + (n_)); } while (0)
        if (tile < num_tiles) input_stage();
        for (; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int row0 = tile * TC_ROWS;
            const int row = row0 + r;
            const uint32_t mb = ti & 1;
            int node_l = row / KNB - row0 / KNB;
            node_l = node_l > 3 ? 3 : node_l;
            mbar_wait(&S.meta_full[mb], (ti >> 1) & 1);
            const float w_r = S.wrow[mb][r], m_i = S.mrow[mb][r];
            EPI_STAMP(0);
            // ---- stage 1: gathered projections, GELU
            mbar_wait(&S.d_ready[g & 1], (g >> 1) & 1);
            tc_fence_after();
            EPI_STAMP(2);
            {
                const uint32_t dcol = (g & 1) ? COL_D1 : COL_D0;
                uint32_t accb[2][16];                          // accumulator loads run one sub-chunk ahead of the math
                tmem_ld16(lane_addr + dcol + set * 16, accb[0]);
#pragma unroll
                for (int c = 0; c < SPT; ++c) {
                    const int sc = set + NSETS * c;
                    uint32_t (&acc)[16] = accb[c & 1];
                    const float* chunk = act_wait(kc + (sc >> 1));
                    float qv[16], pv[16], v[16];
                    load_act16(qv, chunk, r, sc & 1);
                    act_release(kc + (sc >> 1));
                    load_vec16(pv, &S.p[mb][node_l][sc * 16]);
                    tmem_wait_ld();
                    if (c + 1 < SPT) tmem_ld16(lane_addr + dcol + (sc + NSETS) * 16, accb[(c + 1) & 1]);
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {      // packed pairs: same per-lane operation order as the scalar form
                        const f32x2 a2 = pk2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]));
                        const f32x2 q2 = pk2(qv[j], qv[j + 1]), p2 = pk2(pv[j], pv[j + 1]);
                        const f32x2 pre = (KIND == TC_DEC) ? fma2(pk2(m_i, m_i), add2(a2, q2), p2) : add2(add2(a2, p2), q2);
                        upk2(gelu_fast2(pre), v[j], v[j + 1]);
                    }
                    publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                }
                kc += 4;
            }
            ++g;
            EPI_STAMP(3);
            if (KIND == TC_UPD) {
                // ---- stage 2: + b12, GELU
                mbar_wait(&S.d_ready[g & 1], (g >> 1) & 1);
                tc_fence_after();
                EPI_STAMP(4);
                {
                    const uint32_t dcol = (g & 1) ? COL_D1 : COL_D0;
                    uint32_t accb[2][16];
                    tmem_ld16(lane_addr + dcol + set * 16, accb[0]);
#pragma unroll
                    for (int c = 0; c < SPT; ++c) {
                        const int sc = set + NSETS * c;
                        uint32_t (&acc)[16] = accb[c & 1];
                        float bv[16], v[16];
                        load_vec16(bv, &S.vec[0][sc * 16]);
                        tmem_wait_ld();
                        if (c + 1 < SPT) tmem_ld16(lane_addr + dcol + (sc + NSETS) * 16, accb[(c + 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            upk2(gelu_fast2(add2(pk2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1])), pk2(bv[j], bv[j + 1]))), v[j], v[j + 1]);
                        publish_A16(lane_addr, sc, v, &S.a_ready[sc]);
                    }
                }
                ++g;
                EPI_STAMP(5);
            }
            // ---- last GEMM of this tile done -> A is free: feed the NEXT tile's input so its first GEMM runs under our final stage
            mbar_wait(&S.d_ready[g & 1], (g >> 1) & 1);
            tc_fence_after();
            EPI_STAMP(KIND == TC_UPD ? 6 : 4);
            const uint32_t dcol = (g & 1) ? COL_D1 : COL_D0;
            ++g;
            if (KIND != TC_UPD) {
                if (tile + (int)gridDim.x < num_tiles) input_stage();
                EPI_STAMP(1);
            }
            if (KIND == TC_UPD) {
                // ---- final: + b13 + residual, LayerNorm over the row (two threads per row), direct 64-byte row stores
                float shift = 0.f, s1 = 0.f, s2 = 0.f;
                uint32_t accb[2][16];
                tmem_ld16(lane_addr + dcol + set * 16, accb[0]);
#pragma unroll
                for (int c = 0; c < SPT; ++c) {
                    const int sc = set + NSETS * c;
                    uint32_t (&acc)[16] = accb[c & 1];
                    const float* chunk = act_wait(kc + (sc >> 1));
                    float old[16], bv[16];
                    load_act16_sw(old, chunk, r, sc & 1);
                    act_release(kc + (sc >> 1));
                    load_vec16(bv, &S.vec[1][sc * 16]);
                    tmem_wait_ld();
                    if (c + 1 < SPT) tmem_ld16(lane_addr + dcol + (sc + NSETS) * 16, accb[(c + 1) & 1]);
                    if (c == 0) shift = __uint_as_float(acc[0]) + bv[0] + old[0];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float x = __uint_as_float(acc[j]) + bv[j] + old[j];
                        acc[j] = __float_as_uint(x);
                        const float dlt = x - shift;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                    }
                    tmem_st16(lane_addr + dcol + sc * 16, acc);
                }
                kc += 4;
                tmem_wait_st();
                // this thread's 128/NSETS values: (mean, M2); combine the NSETS partials of the row in set order (Chan et al.)
                constexpr float MV = (float)(HID / NSETS);
                const float mean_a = shift + s1 * (1.0f / MV);
                const float m2_a = fmaxf(s2 - s1 * s1 * (1.0f / MV), 0.f);
                S.lnx[set][r] = make_float2(mean_a, m2_a);
                if (set == 0) EPI_STAMP(8);
                // the residual chunks are consumed: the ring refills with the next tile's input while we convert it
                if (tile + (int)gridDim.x < num_tiles) input_stage();
                EPI_STAMP(1);
                named_bar(3, EPI_THREADS);
                float mean = 0.f, m2 = 0.f;
#pragma unroll
                for (int q = 0; q < NSETS; ++q) mean += S.lnx[q][r].x;
                mean *= (1.0f / NSETS);
#pragma unroll
                for (int q = 0; q < NSETS; ++q) { const float2 o = S.lnx[q][r]; const float dm = o.x - mean; m2 += o.y + MV * dm * dm; }
                const float var = m2 * (1.0f / HID);
                const float rstd = 1.0f / sqrtf(var + LN_EPS);
                float* ost = &S.out[0][0];                      // [128 rows][64 cols], 16-byte chunks XOR-swizzled by row
                for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                    for (int cc = 0; cc < SPT / 2; ++cc) {
                        const int sc = 4 * hf + NSETS * cc + set;   // this thread's sub-chunk(s) of the half
                        uint32_t xv[16];
                        tmem_ld16(lane_addr + dcol + sc * 16, xv);
                        float gv[16], bv[16];
                        load_vec16(gv, &S.vec[2][sc * 16]);
                        load_vec16(bv, &S.vec[3][sc * 16]);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ch = (sc & 3) * 4 + j;            // 16-byte chunk 0..15 inside the 64-column half
                            *reinterpret_cast<float4*>(ost + r * 64 + ((ch ^ (r & 15)) << 2)) =
                                make_float4((__uint_as_float(xv[4 * j]) - mean) * rstd * gv[4 * j] + bv[4 * j],
                                            (__uint_as_float(xv[4 * j + 1]) - mean) * rstd * gv[4 * j + 1] + bv[4 * j + 1],
                                            (__uint_as_float(xv[4 * j + 2]) - mean) * rstd * gv[4 * j + 2] + bv[4 * j + 2],
                                            (__uint_as_float(xv[4 * j + 3]) - mean) * rstd * gv[4 * j + 3] + bv[4 * j + 3]);
                        }
                    }
                    named_bar(3, EPI_THREADS);
                    {   // 128 rows x 256 B: 16 lanes per row
                        const int t = set * 128 + r;
#pragma unroll
                        for (int i = 0; i < 2048 / EPI_THREADS; ++i) {
                            const int idx = i * EPI_THREADS + t, rr = idx >> 4, ch = idx & 15;
                            if (row0 + rr < total_rows)
                                *reinterpret_cast<float4*>(a.hE + (size_t)(row0 + rr) * HID + hf * 64 + ch * 4) =
                                    *reinterpret_cast<const float4*>(ost + rr * 64 + ((ch ^ (rr & 15)) << 2));
                        }
                    }
                    named_bar(3, EPI_THREADS);
                }
            } else {
                // ---- final: + b2, GELU, weight, per-residue column sums of each 16-row half-warp
                const int sb = (row0 + quad * 32) / 16 + (lane >> 4);   // 16-row sub-block of the edge array
                const int node = sb / 3, part = sb - node * 3;
                for (int c = 0; c < SPT; c += 2) {                     // two 16-column sub-chunks per transpose-reduce
                    float gv[32];
                    uint32_t acc2[2][16];                              // both sub-chunks' loads in flight before the first wait
                    tmem_ld16(lane_addr + dcol + (NSETS * c + set) * 16, acc2[0]);
                    tmem_ld16(lane_addr + dcol + (NSETS * (c + 1) + set) * 16, acc2[1]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int sc = NSETS * (c + h) + set;
                        uint32_t (&acc)[16] = acc2[h];
                        float bv[16];
                        load_vec16(bv, &S.vec[0][sc * 16]);
                        if (h == 0) tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float g0, g1;
                            upk2(gelu_fast2(add2(pk2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1])), pk2(bv[j], bv[j + 1]))), g0, g1);
                            gv[h * 16 + j] = (w_r != 0.f) ? w_r * g0 : 0.f;
                            gv[h * 16 + j + 1] = (w_r != 0.f) ? w_r * g1 : 0.f;
                        }
                    }
                    halfwarp_colsum32(gv, lane);       // lane l: local columns 2(l&15), 2(l&15)+1 of the 32 gathered above
                    if (node < N) {
                        const int lc = 2 * (lane & 15), h = lc >> 4;             // which of the two sub-chunks, column inside it
                        const int col = (NSETS * (c + h) + set) * 16 + (lc & 15);
                        *reinterpret_cast<float2*>(a.Spart + ((size_t)node * 3 + part) * HID + col) = make_float2(gv[0], gv[1]);
                    }
                }
            }
            tc_fence_before();
            EPI_STAMP(7);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// edge_feat (tensor cores): per edge 16 positional + 25x16 RBF features (computed by the row's thread straight
// into the A operand) -> edge_embedding (416 -> 128, 13 K-chunks through 4 recycled A slots) -> LayerNorm
// (thread-local) -> W_e + b  => h_E          (protein_mpnn_utils.py:1111-1125, 1142-1179, 1229)
// ---------------------------------------------------------------------------------------------------------
constexpr float RBF_C = 0.8f * 1.2011224087864498f;    // (1 / sigma) * sqrt(log2 e), sigma = 1.25 (protein_mpnn_utils.py:1115-1118)
__constant__ int8_t c_tc_pair_a[NPAIR] = {1, 0, 2, 3, 4, 1, 1, 1, 1, 0, 0, 0, 4, 4, 3, 0, 2, 3, 4, 2, 3, 4, 2, 3, 2};
__constant__ int8_t c_tc_pair_b[NPAIR] = {1, 0, 2, 3, 4, 0, 2, 3, 4, 2, 3, 4, 2, 3, 2, 1, 1, 1, 1, 0, 0, 0, 4, 4, 3};

struct __align__(1024) TcFeatSmem {
    uint8_t wring[TC_WSLOTS][TC_WCHUNK_BYTES];   // 96 KB
    float out[TC_ROWS * 64];                     // 32 KB output staging: [128 rows][64 cols], 16-byte chunks XOR (row & 15)
    float ajT[16][TC_ROWS];                      // neighbour atoms, component-major: [3*atom + xyz][row]
    float ai[4][16];                             // atoms of the (up to 4) residues the tile touches
    float dnb[TC_ROWS];
    int posq[TC_ROWS];
    float pos_table[66 * 16];
    float vec[3][HID];                           // LN gain, LN shift, W_e bias
    float muc[16];                               // -c mu_t, c = 0.8 sqrt(log2 e)  (RBF centres pre-scaled for the ex2 form)
    float2 lnx[2][TC_ROWS];
    uint64_t w_full[TC_WSLOTS], w_empty[TC_WSLOTS];
    uint64_t a_ready[4], a_free[4];
    uint64_t d_ready[2];
    uint64_t in_full, in_empty;
    uint32_t tmem_base;
};

struct TcFeatArgs {
    const float* atoms;       // [N][16]
    const int* E_idx;         // [N*48]
    const float* D_nb;        // [N*48]
    const int* ridx;
    const int* cenc;
    const float* pos_table;   // [66][16]
    const float* rbf_mu;      // [16]
    const uint8_t* wimg;      // 13 + 4 chunks
    const float* ln_g;
    const float* ln_b;
    const float* b_e;
    float* hE;
    int N, num_tiles;
};

// Warp roles as in tc_edge_kernel: warps 0-7 epilogue (set s = warp/4 handles the 32-column chunks c with c%2 == s),
// warp 8 MMA issuer, warp 9 weight producer, warp 10 per-row input producer.
__global__ void __launch_bounds__(T2_THREADS, 1) tc_edge_feat_kernel(TcFeatArgs a) {
    constexpr int K1_CHUNKS = EDGE_IN / 32;   // 13
    extern __shared__ uint8_t smem_raw[];
    TcFeatSmem& S = *reinterpret_cast<TcFeatSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, num_tiles = a.num_tiles;
    const int total_rows = N * KNB;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_WSLOTS; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&S.a_ready[i], 128); mbar_init(&S.a_free[i], 1); }
        for (int i = 0; i < 2; ++i) mbar_init(&S.d_ready[i], 1);
        mbar_init(&S.in_full, 32);
        mbar_init(&S.in_empty, 256);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < HID; i += T2_THREADS) { S.vec[0][i] = a.ln_g[i]; S.vec[1][i] = a.ln_b[i]; S.vec[2][i] = a.b_e[i]; }
    for (int i = threadIdx.x; i < 66 * 16; i += T2_THREADS) S.pos_table[i] = a.pos_table[i];
    if (threadIdx.x < 16) S.muc[threadIdx.x] = -RBF_C * a.rbf_mu[threadIdx.x];
    if (warp == 8) tmem_alloc<512>(&S.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    pdl_trigger();                                        // (see tc_edge_kernel)
    if (warp != 9) pdl_wait();

    if (warp == 9) {
        if (lane == 0) {                                  // weight producer
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int c = 0; c < K1_CHUNKS + 4; ++c, ++it) {
                    const uint32_t slot = it % TC_WSLOTS, use = it / TC_WSLOTS;
                    if (use > 0) mbar_wait(&S.w_empty[slot], (use - 1) & 1);
                    mbar_arrive_expect_tx(&S.w_full[slot], TC_WCHUNK_BYTES);
                    bulk_g2s(S.wring[slot], a.wimg + (size_t)c * TC_WCHUNK_BYTES, TC_WCHUNK_BYTES, &S.w_full[slot]);
                }
        }
    } else if (warp == 8) {
        {                                                 // MMA issuer (whole warp runs the loop; one elected lane issues)
            // GEMM1 (416 deep): hi.hi products accumulate in D0, the small corrections (lo.hi, hi.lo) in D1 -- with 13
            // K-chunks the truncating accumulator would otherwise lose them against the large partial sums (this GEMM was
            // the whole end-to-end error of the tensor-core path).  The epilogue adds D0 + D1.  GEMM2 (W_e) -> D1.
            const uint32_t idesc = make_idesc_tf32(128, 128);
            const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);   // warp-uniform for the compiler
            uint32_t it = 0, apar = 0;                        // apar: bit s = parity of a_ready[s]'s next completion
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
#pragma unroll 1
                for (int l = 0; l < 2; ++l) {
                    const int nch = l == 0 ? K1_CHUNKS : 4;
                    const uint32_t d_main = tmem + (l == 0 ? COL_D0 : COL_D1), d_corr = tmem + COL_D1;
#pragma unroll 1
                    for (int c = 0; c < nch; ++c, ++it) {
                        const uint32_t slot = it % TC_WSLOTS, as = c & 3;
                        mbar_wait(&S.a_ready[as], (apar >> as) & 1);
                        apar ^= 1u << as;
                        mbar_wait(&S.w_full[slot], (it / TC_WSLOTS) & 1);
                        tc_fence_after();
                        const uint32_t bh = desc_sw128_lo(smem_u32(S.wring[slot]));
                        const uint32_t ahi = tmem + COL_AHI + as * 32, alo = tmem + COL_ALO + as * 32;
                        if (elect_one()) {
                            if (c == 0) mma_tf32_ts_lh_init(d_corr, alo, bh, DESC_SW128_HI, idesc);
                            else mma_tf32_ts_lh(d_corr, alo, bh, DESC_SW128_HI, idesc);
                            mma_tf32_ts_lh(d_corr, ahi, bh + 1024, DESC_SW128_HI, idesc);
#pragma unroll
                            for (int s = 1; s < 4; ++s) {
                                mma_tf32_ts_lh(d_corr, alo + s * 8, bh + s * 2, DESC_SW128_HI, idesc);
                                mma_tf32_ts_lh(d_corr, ahi + s * 8, bh + 1024 + s * 2, DESC_SW128_HI, idesc);
                            }
                            if (l == 0 && c == 0) mma_tf32_ts_lh_init(d_main, ahi, bh, DESC_SW128_HI, idesc);
                            else mma_tf32_ts_lh(d_main, ahi, bh, DESC_SW128_HI, idesc);
#pragma unroll
                            for (int s = 1; s < 4; ++s) mma_tf32_ts_lh(d_main, ahi + s * 8, bh + s * 2, DESC_SW128_HI, idesc);
                            mma_commit(&S.w_empty[slot]);
                            mma_commit(&S.a_free[as]);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) mma_commit(&S.d_ready[l]);
                }
        }
    } else if (warp == 10) {
        uint32_t ti = 0;                                  // per-row inputs
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int row0 = tile * TC_ROWS;
            const int node_lo = row0 / KNB;
            if (ti > 0) mbar_wait(&S.in_empty, (ti - 1) & 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = lane + 32 * i, row = row0 + r;
                int self = row / KNB; self = self < N ? self : N - 1;
                int j = -1;
                float dn = 0.f;
                if (row < total_rows) { j = a.E_idx[row]; if (j >= 0) dn = a.D_nb[row]; }
                if (j < 0) j = self;                       // unused slot: a well-defined dummy edge, never read downstream
                S.dnb[r] = dn;
                int q = a.ridx[self] - a.ridx[j] + 32;
                q = q < 0 ? 0 : (q > 64 ? 64 : q);
                S.posq[r] = (a.cenc[self] == a.cenc[j]) ? q : 65;          // PositionalEncodings, protein_mpnn_utils.py:904-905
                const float4* src = reinterpret_cast<const float4*>(a.atoms + (size_t)j * 16);
                const float4 a0 = __ldg(src), a1 = __ldg(src + 1), a2 = __ldg(src + 2), a3 = __ldg(src + 3);
                const float v[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
#pragma unroll
                for (int k = 0; k < 16; ++k) S.ajT[k][r] = v[k];
            }
            if (lane < 4) {
                int node = node_lo + lane; node = node < N ? node : N - 1;
                const float4* src = reinterpret_cast<const float4*>(a.atoms + (size_t)node * 16);
                float4* dst = reinterpret_cast<float4*>(S.ai[lane]);
                dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = __ldg(src + 3);
            }
            mbar_arrive(&S.in_full);
        }
    } else {
        const int set = warp >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        uint32_t ti = 0, pub = 0;      // pub: bit s = parity of a_free[s]'s next completion, bit 8+s = slot s used before
        auto publish = [&](int slot, const float (&v)[32]) {
            if ((pub >> (8 + slot)) & 1) {                       // wait until the MMAs that read this slot are done
                mbar_wait(&S.a_free[slot], (pub >> slot) & 1);
                pub ^= 1u << slot;
            }
            pub |= 1u << (8 + slot);
            tc_fence_after();
            publish_A_chunk(lane_addr, slot, v, &S.a_ready[slot]);
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int row0 = tile * TC_ROWS;
            const int row = row0 + r;
            int node_l = row / KNB - row0 / KNB;
            node_l = node_l > 3 ? 3 : node_l;
            mbar_wait(&S.in_full, ti & 1);
            // ---- features -> A, 32 columns at a time; this set takes the chunks of its parity
            for (int c = set; c < K1_CHUNKS; c += 2) {
                float v[32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c == 0 && h == 0) {
                        const float* pt = &S.pos_table[S.posq[r] * 16];
#pragma unroll
                        for (int t = 0; t < 16; ++t) v[t] = pt[t];
                    } else {
                        const int p = (c == 0) ? 0 : 2 * c - 1 + h;
                        float D;
                        if (p == 0) {
                            D = S.dnb[r];                  // the masked top-k distance itself (protein_mpnn_utils.py:1142)
                        } else {
                            const int ia = 3 * c_tc_pair_a[p], ib = 3 * c_tc_pair_b[p];
                            const float dx = __fsub_rn(S.ai[node_l][ia], S.ajT[ib][r]);
                            const float dy = __fsub_rn(S.ai[node_l][ia + 1], S.ajT[ib + 1][r]);
                            const float dz = __fsub_rn(S.ai[node_l][ia + 2], S.ajT[ib + 2][r]);
                            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                            D = __fsqrt_rn(__fadd_rn(d2, 1e-6f));
                        }
                        // rbf_t = exp(-((D - mu_t) / 1.25)^2) = 2^(-(c D - c mu_t)^2), c = 0.8 sqrt(log2 e); two per instruction
                        const float cD = D * RBF_C;
                        const f32x2 cd2 = pk2(cD, cD);
#pragma unroll
                        for (int t = 0; t < 16; t += 2) {
                            const f32x2 u = add2(cd2, pk2(S.muc[t], S.muc[t + 1]));     // muc = -c mu
                            float w0, w1, e0, e1;
                            upk2(mul2(u, u), w0, w1);
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-w0));
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-w1));
                            v[h * 16 + t] = e0;
                            v[h * 16 + t + 1] = e1;
                        }
                    }
                }
                publish(c & 3, v);
            }
            mbar_arrive(&S.in_empty);
            // ---- norm_edges: LayerNorm over the 128 accumulator columns (two threads per row) -> A of the W_e GEMM
            mbar_wait(&S.d_ready[0], ti & 1);
            tc_fence_after();
            {
                float shift = 0.f, s1 = 0.f, s2 = 0.f;
                for (int c = set; c < 4; c += 2) {
                    uint32_t acc[32], cor[32];
                    tmem_ld32(lane_addr + COL_D0 + c * 32, acc);
                    tmem_ld32(lane_addr + COL_D1 + c * 32, cor);
                    tmem_wait_ld();
                    if (c == set) shift = __uint_as_float(acc[0]) + __uint_as_float(cor[0]);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = __uint_as_float(acc[j]) + __uint_as_float(cor[j]);
                        acc[j] = __float_as_uint(x);
                        const float dlt = x - shift;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                    }
                    tmem_st32(lane_addr + COL_D0 + c * 32, acc);
                }
                tmem_wait_st();
                const float mean_a = shift + s1 * (1.0f / 64.0f);
                const float m2_a = fmaxf(s2 - s1 * s1 * (1.0f / 64.0f), 0.f);
                S.lnx[set][r] = make_float2(mean_a, m2_a);
                tc_fence_before();
                named_bar(3, 256);       // partner's statistics visible; every row has finished reading D1 (the W_e GEMM overwrites it)
                tc_fence_after();
                const float2 ob = S.lnx[set ^ 1][r];
                const float dm = mean_a - ob.x;
                const float mean = 0.5f * (mean_a + ob.x);
                const float var = (m2_a + ob.y + 32.0f * dm * dm) * (1.0f / HID);
                const float rstd = 1.0f / sqrtf(var + LN_EPS);
                for (int c = set; c < 4; c += 2) {
                    uint32_t acc[32];
                    tmem_ld32(lane_addr + COL_D0 + c * 32, acc);
                    tmem_wait_ld();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = (__uint_as_float(acc[j]) - mean) * rstd * S.vec[0][c * 32 + j] + S.vec[1][c * 32 + j];
                    publish(c, v);
                }
            }
            // ---- W_e output + bias -> staging (two 64-column halves) -> coalesced store
            mbar_wait(&S.d_ready[1], ti & 1);
            tc_fence_after();
            for (int hf = 0; hf < 2; ++hf) {
                const int c = 2 * hf + set;                     // this thread's 32-column chunk of the half
                uint32_t acc[32];
                tmem_ld32(lane_addr + COL_D1 + c * 32, acc);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ch = set * 8 + j;                 // 16-byte chunk 0..15 inside the 64-column half
                    *reinterpret_cast<float4*>(S.out + r * 64 + ((ch ^ (r & 15)) << 2)) =
                        make_float4(__uint_as_float(acc[4 * j]) + S.vec[2][c * 32 + 4 * j], __uint_as_float(acc[4 * j + 1]) + S.vec[2][c * 32 + 4 * j + 1],
                                    __uint_as_float(acc[4 * j + 2]) + S.vec[2][c * 32 + 4 * j + 2], __uint_as_float(acc[4 * j + 3]) + S.vec[2][c * 32 + 4 * j + 3]);
                }
                named_bar(3, 256);
                {
                    const int t = set * 128 + r;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int idx = i * 256 + t, rr = idx >> 4, ch = idx & 15;
                        if (row0 + rr < total_rows)
                            *reinterpret_cast<float4*>(a.hE + (size_t)(row0 + rr) * HID + hf * 64 + ch * 4) =
                                *reinterpret_cast<const float4*>(S.out + rr * 64 + ((ch ^ (rr & 15)) << 2));
                    }
                }
                named_bar(3, 256);
            }
            tc_fence_before();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

// Tensor map of h_E for the activation producers.  cuTensorMapEncodeTiled is a driver entry point, fetched through the
// runtime so the library needs no link-time libcuda.  The last encoding is cached per thread (one context per thread).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static thread_local cudaError_t t_tc_error = cudaSuccess;

cudaError_t tc_take_error() { cudaError_t e = t_tc_error; t_tc_error = cudaSuccess; return e; }

static bool hE_tensor_map(CUtensorMap* out, const float* hE, long long rows) {
    static thread_local const float* c_ptr = nullptr;
    static thread_local long long c_rows = -1;
    static thread_local CUtensorMap c_map;
    if (c_ptr != hE || c_rows != rows) {
        if (!g_encode_tiled) return false;
        const cuuint64_t dims[2] = {(cuuint64_t)HID, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)HID * sizeof(float)};
        const cuuint32_t box[2] = {32, (cuuint32_t)TC_ROWS};
        const cuuint32_t estr[2] = {1, 1};
        if (g_encode_tiled(&c_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(hE), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
        c_ptr = hE; c_rows = rows;
    }
    *out = c_map;
    return true;
}

constexpr int UPD_SETS = 2, MSG_SETS = 2;   // epilogue sets (4 warps each) per kernel kind
constexpr int UPD_NAP = 4, MSG_NAP = 4;     // activation-producer warps

cudaError_t tc_kernels_init() {
    cudaError_t e;
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if ((e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres)) != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const int bytes = (int)sizeof(Tc2Smem) + 1024;
    if ((e = cudaFuncSetAttribute(tc_edge_kernel<TC_UPD, UPD_SETS, UPD_NAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tc_edge_kernel<TC_ENC, MSG_SETS, MSG_NAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tc_edge_kernel<TC_DEC, MSG_SETS, MSG_NAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_edge_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcFeatSmem) + 1024);
}

void launch_tc_edge_feat(const float* atoms, const int* E_idx, const float* D_nb, const int* ridx, const int* cenc,
                         const float* pos_table, const float* rbf_mu, const float* wimg, const float* ln_g, const float* ln_b,
                         const float* b_e, float* hE, int N, int sms, cudaStream_t st) {
    TcFeatArgs a{};
    a.atoms = atoms; a.E_idx = E_idx; a.D_nb = D_nb; a.ridx = ridx; a.cenc = cenc; a.pos_table = pos_table; a.rbf_mu = rbf_mu;
    a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.ln_g = ln_g; a.ln_b = ln_b; a.b_e = b_e; a.hE = hE; a.N = N;
    a.num_tiles = (N * KNB + TC_ROWS - 1) / TC_ROWS;
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    launch_pdl(tc_edge_feat_kernel, dim3(grid), dim3(T2_THREADS), sizeof(TcFeatSmem) + 1024, st, a);
}

static TcArgs tc_args(float* hE, const int* E_idx, const float* P, int p_stride, const float* Q, int q_stride, const float* maskf,
                      const float* wimg, int N) {
    TcArgs a{};
    if (N <= 0) return a;
    if (!hE_tensor_map(&a.tm_hE, hE, (long long)N * KNB)) { t_tc_error = cudaErrorInvalidValue; a.N = 0; a.num_tiles = 0; return a; }
    a.hE = hE; a.E_idx = E_idx; a.P = P; a.p_stride = p_stride; a.Q = Q; a.q_stride = q_stride; a.maskf = maskf;
    a.wimg = reinterpret_cast<const uint8_t*>(wimg);
    a.N = N;
    a.num_tiles = (N * KNB + TC_ROWS - 1) / TC_ROWS;
    return a;
}

void launch_tc_edge_upd(float* hE, const int* E_idx, const float* P, const float* Q, const float* maskf, const float* wimg,
                        const float* b12, const float* b13, const float* ln_g, const float* ln_b, int N, int sms, cudaStream_t st,
                        long long* trace) {
    TcArgs a = tc_args(hE, E_idx, P, HID, Q, HID, maskf, wimg, N);
    if (a.num_tiles == 0) return;
    a.v0 = b12; a.v1 = b13; a.v2 = ln_g; a.v3 = ln_b; a.trace = trace;
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    launch_pdl(tc_edge_kernel<TC_UPD, UPD_SETS, UPD_NAP>, dim3(grid), dim3((4 * UPD_SETS + 2 + UPD_NAP) * 32), sizeof(Tc2Smem) + 1024, st, a);
}

void launch_tc_msg(bool dec, const float* hE, const int* E_idx, const float* P, int p_stride, const float* Q, int q_stride,
                   const float* maskf, const float* wimg, const float* b2, float* Spart, int N, int sms, cudaStream_t st,
                   long long* trace) {
    TcArgs a = tc_args(const_cast<float*>(hE), E_idx, P, p_stride, Q, q_stride, maskf, wimg, N);
    if (a.num_tiles == 0) return;
    a.v0 = b2; a.Spart = Spart; a.trace = trace;
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    if (dec) launch_pdl(tc_edge_kernel<TC_DEC, MSG_SETS, MSG_NAP>, dim3(grid), dim3((4 * MSG_SETS + 2 + MSG_NAP) * 32), sizeof(Tc2Smem) + 1024, st, a);
    else launch_pdl(tc_edge_kernel<TC_ENC, MSG_SETS, MSG_NAP>, dim3(grid), dim3((4 * MSG_SETS + 2 + MSG_NAP) * 32), sizeof(Tc2Smem) + 1024, st, a);
}

}  // namespace tmpnn
