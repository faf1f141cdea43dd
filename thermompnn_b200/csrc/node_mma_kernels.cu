// node_update for small and medium batches: warp-level tensor-core form (mma.sync m16n8k8, 3xTF32) inside a 4-CTA cluster.
//
//   dh = (W3.Ssum + cnt*b3)/30 ; h = LN1(h_V + dh) ; h = LN2(h + FFN(h)) ; h *= mask          protein_mpnn_utils.py:826-832, 871-880
//   then the per-residue projections the next edge kernels gather (W1a.h, W1c.h, W11a.h, W11c.h; engine.cu).
//
// Why this form: for one 300-residue protein (the headline configuration) node_update is a LATENCY chain, six times per
// call -- 19 clusters, each a dependent sequence W3 -> LN -> FFN-in -> FFN-out -> LN -> projections over 0.8 MB of
// weights.  The FP32 SIMT cluster kernel spent 10 of its 17 us in the four GEMM phases (clock64 trace,
// profiles/r02_node_trace.md); on the warp-level tensor cores each phase is 12-48 mma instructions per warp.
//
// A cluster of four CTAs owns a tile of 16 residues = one m16 MMA tile.  CTA q loads only its quarter of the weights
// (requested up front, before the dependency wait) and keeps it resident:
//   W3      K-split   CTA q: input columns [32q, 32q+32)            -> partial [16][128]  -> cluster all-reduce (DSMEM)
//   W_in    N-split   CTA q: hidden block q (columns [128q, 128q+128))                     -> stays local ...
//   W_out   K-split   CTA q: rows [128q, 128q+128) = exactly the hidden block it produced  -> partial -> all-reduce
//   proj    N-split   CTA q: projection q
// The all-reduces add the four partials in rank order, so the result does not depend on timing; a residue's result does
// not depend on what else is in the batch (every (row, column) runs the same instruction sequence wherever its tile sits).
//
// Inside a CTA: 16 warps, warp w owns output columns [8w, 8w+8) (one n8 block); lane (g = lane/4, t = lane%4) holds
// rows {g, g+8} x columns {8w+2t, 8w+2t+1} of every [16][128] result (the m16n8 accumulator layout).
// 3xTF32: x = hi + lo with hi = x & 0xffffe000 (exactly what the tensor core reads of x) and lo = x - hi (exact);
// D = A_lo.B_hi + A_hi.B_lo (small terms, own accumulator) + A_hi.B_hi, fp32 accumulation; dropped lo.lo <= 2^-20.
// Weight chunks ([32 k][128 n] fp32, 16 KB) sit in shared memory with their 16-byte granules XOR-swizzled by (k & 3) so
// that a B-fragment load (rows k0+t, columns 8w+g) is bank-conflict free.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "pdl.cuh"

namespace tmpnn {

namespace {

constexpr int NM_THREADS = 512;
constexpr int NM_SLOTS = 9;          // resident weight chunks: W3 | W_in x4 | W_out x4; the projections reuse slots 1..4
constexpr int NM_T = 16;             // residues per cluster tile
constexpr int NM_P = HID + 4;        // row pitch of activation tiles: rows g -> banks 4g + t, conflict-free A-fragment loads

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split3(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// acc_m / acc_c += A[16][32 k] . W[32 k][n-block of this warp]   (one weight chunk: four k8 steps)
//   A: activation tile in shared memory, row pitch NM_P, k offset already applied;  W: swizzled chunk.
__device__ __forceinline__ void mma_chunk(float (&acc_m)[4], float (&acc_c)[4], const float* __restrict__ A,
                                          const float* __restrict__ W, int g, int t, int ncol) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int k0 = 8 * ks;
        uint32_t ah[4], al[4], bh[2], bl[2];
        split3(A[g * NM_P + k0 + t], ah[0], al[0]);
        split3(A[(g + 8) * NM_P + k0 + t], ah[1], al[1]);
        split3(A[g * NM_P + k0 + t + 4], ah[2], al[2]);
        split3(A[(g + 8) * NM_P + k0 + t + 4], ah[3], al[3]);
        const int col = ncol ^ (t << 3);                    // (k & 3) == t for both rows below
        split3(W[(k0 + t) * HID + col], bh[0], bl[0]);
        split3(W[(k0 + t + 4) * HID + col], bh[1], bl[1]);
        mma_tf32(acc_c, al, bh[0], bh[1]);
        mma_tf32(acc_c, ah, bl[0], bl[1]);
        mma_tf32(acc_m, ah, bh[0], bh[1]);
    }
}

}  // namespace

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NM_THREADS, 1) node_update_mma_kernel(NodeUpdArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                              // [NM_SLOTS][4096]
    float* x = ring + NM_SLOTS * WCHUNK_FLOATS;      // [16][NM_P]  neighbour sum, then LN1 output, then the final h
    float* hid = x + NM_T * NM_P;                    // [16][NM_P]  this CTA's block of the FFN hidden layer
    float* xr0 = hid + NM_T * NM_P;                  // [16][128]   this CTA's partials, read by the other three CTAs
    float* xr1 = xr0 + NM_T * HID;                   //             (one buffer per all-reduce: no second cluster barrier)
    float2* red = reinterpret_cast<float2*>(xr1 + NM_T * HID);   // [2 LayerNorms][16 warps][16 rows] (mean, M2) partials
    float* cnt_s = reinterpret_cast<float*>(red + 2 * 16 * NM_T);   // [16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ncol = 8 * warp + g;                   // B-fragment column of this lane
    const int c0 = 8 * warp + 2 * t;                 // this lane's two output columns: c0, c0 + 1; rows g and g + 8
    const int node0 = (blockIdx.x >> 2) * NM_T;
    const float* wb = a.weights;
#define NM_STAMP(i_) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[i_] = clock64(); } while (0)
    NM_STAMP(0);

    const bool has_proj = a.nproj > 0 && q < a.nproj;
    auto issue = [&](int k, int slot) {                // chunk k of this CTA -> slot (swizzled); always one commit group
        const float* src = wb + a.w.stream + (size_t)(4 * k + q) * WCHUNK_FLOATS;
        float* dst = ring + slot * WCHUNK_FLOATS;
#pragma unroll
        for (int i = 0; i < WCHUNK_FLOATS / 4 / NM_THREADS; ++i) {
            const int gi = threadIdx.x + i * NM_THREADS;               // granule index: row kk = gi / 32, granule gn = gi % 32
            const int kk = gi >> 5, gn = gi & 31;
            cp_async16(dst + kk * HID + ((gn ^ ((kk & 3) << 1)) << 2), src + gi * 4);
        }
        cp_async_commit();
    };
    pdl_trigger();
#pragma unroll
    for (int k = 0; k < NM_SLOTS; ++k) issue(k, k);
    // This lane's slices of the bias / LayerNorm vectors: constants, fetched before the dependency wait (each would
    // otherwise be an exposed L2 round trip in the middle of the chain -- cluster.sync invalidates L1)
    auto ld2 = [&](int64_t off) { return __ldg(reinterpret_cast<const float2*>(wb + off + c0)); };
    const float2 c_b3 = ld2(a.w.b3), c_n1g = ld2(a.w.n1.g), c_n1b = ld2(a.w.n1.b), c_bin = ld2(a.w.bin + q * HID), c_bout = ld2(a.w.bout),
                 c_n2g = ld2(a.w.n2.g), c_n2b = ld2(a.w.n2.b);
    const float2 c_pb = (has_proj && a.proj[q].bias >= 0) ? ld2(a.proj[q].bias) : make_float2(0.f, 0.f);
    NM_STAMP(1);
    pdl_wait();                                      // everything below reads earlier kernels' output
    NM_STAMP(2);

    // ---- inputs: neighbour sums (three 16-row partials per residue, added in order), edge-weight counts, h_V
    for (int i = threadIdx.x; i < NM_T * (HID / 4); i += NM_THREADS) {
        const int r = i / (HID / 4), c4 = (i % (HID / 4)) * 4, node = node0 + r;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (node < a.N) {
            const float* sp = a.Spart + (size_t)node * 3 * HID + c4;
            const float4 p0 = *reinterpret_cast<const float4*>(sp), p1 = *reinterpret_cast<const float4*>(sp + HID),
                         p2 = *reinterpret_cast<const float4*>(sp + 2 * HID);
            s = make_float4((p0.x + p1.x) + p2.x, (p0.y + p1.y) + p2.y, (p0.z + p1.z) + p2.z, (p0.w + p1.w) + p2.w);
        }
        *reinterpret_cast<float4*>(x + r * NM_P + c4) = s;
    }
    if (threadIdx.x < NM_T) cnt_s[threadIdx.x] = (node0 + threadIdx.x < a.N) ? a.cnt_pre[node0 + threadIdx.x] : 0.f;
    float hv[4];                                     // h_V at (g, c0), (g, c0+1), (g+8, c0), (g+8, c0+1)
    float2 tv0 = make_float2(0.f, 0.f), tv1 = tv0;   // rows of the per-letter table this CTA's projection adds (decoder Q)
    {
        const int n0 = node0 + g, n1 = node0 + g + 8;
        if (has_proj && a.proj[q].table >= 0) {      // requested now, used at the very end
            if (n0 < a.N) tv0 = __ldg(reinterpret_cast<const float2*>(wb + a.proj[q].table + (size_t)a.S[n0] * HID + c0));
            if (n1 < a.N) tv1 = __ldg(reinterpret_cast<const float2*>(wb + a.proj[q].table + (size_t)a.S[n1] * HID + c0));
        }
        const float2 v0 = n0 < a.N ? *reinterpret_cast<const float2*>(a.hV + (size_t)n0 * HID + c0) : make_float2(0.f, 0.f);
        const float2 v1 = n1 < a.N ? *reinterpret_cast<const float2*>(a.hV + (size_t)n1 * HID + c0) : make_float2(0.f, 0.f);
        hv[0] = v0.x; hv[1] = v0.y; hv[2] = v1.x; hv[3] = v1.y;
    }
    // sum of the four CTAs' partials, in rank order
    auto cluster_allreduce = [&](float (&v)[4], float* xr) {
        *reinterpret_cast<float2*>(xr + g * HID + c0) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(xr + (g + 8) * HID + c0) = make_float2(v[2], v[3]);
        if (xr == xr0) NM_STAMP(13);
        cluster.sync();
        if (xr == xr0) NM_STAMP(14);
        float2 p[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float* rp = cluster.map_shared_rank(xr, r);
            p[r][0] = *reinterpret_cast<const float2*>(rp + g * HID + c0);
            p[r][1] = *reinterpret_cast<const float2*>(rp + (g + 8) * HID + c0);
        }
        v[0] = (p[0][0].x + p[1][0].x) + (p[2][0].x + p[3][0].x);
        v[1] = (p[0][0].y + p[1][0].y) + (p[2][0].y + p[3][0].y);
        v[2] = (p[0][1].x + p[1][1].x) + (p[2][1].x + p[3][1].x);
        v[3] = (p[0][1].y + p[1][1].y) + (p[2][1].y + p[3][1].y);
    };
    // LayerNorm over the 128 columns of rows g and g + 8: every warp reduces its 8 columns to (mean, M2) per row, the 16
    // partials are combined in warp order (Chan et al.: equal counts) -- ONE block barrier; `red` is a fresh buffer per call
    auto layernorm = [&](float (&v)[4], const float2 gg, const float2 bb, float2* red) {
        float s0 = v[0] + v[1], s1 = v[2] + v[3];
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float mw0 = s0 * 0.125f, mw1 = s1 * 0.125f;
        float e0 = v[0] - mw0, e1 = v[1] - mw0, e2 = v[2] - mw1, e3 = v[3] - mw1;
        float q0 = fmaf(e0, e0, e1 * e1), q1 = fmaf(e2, e2, e3 * e3);
        q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
        if (t == 0) { red[warp * NM_T + g] = make_float2(mw0, q0); red[warp * NM_T + g + 8] = make_float2(mw1, q1); }
        __syncthreads();
        float m0 = 0.f, m1 = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) { m0 += red[w * NM_T + g].x; m1 += red[w * NM_T + g + 8].x; }
        m0 *= (1.0f / 16.0f); m1 *= (1.0f / 16.0f);
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            const float2 a0 = red[w * NM_T + g], a1 = red[w * NM_T + g + 8];
            const float d0 = a0.x - m0, d1 = a1.x - m1;
            v0 += fmaf(8.0f * d0, d0, a0.y);
            v1 += fmaf(8.0f * d1, d1, a1.y);
        }
        const float r0 = 1.0f / sqrtf(v0 * (1.0f / HID) + LN_EPS), r1 = 1.0f / sqrtf(v1 * (1.0f / HID) + LN_EPS);
        v[0] = (v[0] - m0) * r0 * gg.x + bb.x; v[1] = (v[1] - m0) * r0 * gg.y + bb.y;
        v[2] = (v[2] - m1) * r1 * gg.x + bb.x; v[3] = (v[3] - m1) * r1 * gg.y + bb.y;
    };
    auto store_tile = [&](float* dst, const float (&v)[4]) {
        *reinterpret_cast<float2*>(dst + g * NM_P + c0) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(dst + (g + 8) * NM_P + c0) = make_float2(v[2], v[3]);
    };

    float am[4], ac[4], h[4];
    NM_STAMP(3);
    // ---- W3 (this CTA: input columns 32q..)
    cp_async_wait<NM_SLOTS - 1>();                   // chunk 0
    __syncthreads();                                 // ... of every thread; also publishes x / cnt_s
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
    mma_chunk(am, ac, x + 32 * q, ring, g, t, ncol);
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i] += ac[i];
    NM_STAMP(4);
    cluster_allreduce(am, xr0);
    NM_STAMP(5);
    {
        const float2 b3 = c_b3;
        const float cg0 = cnt_s[g], cg1 = cnt_s[g + 8];
        h[0] = hv[0] + fmaf(cg0, b3.x, am[0]) / MSG_SCALE; h[1] = hv[1] + fmaf(cg0, b3.y, am[1]) / MSG_SCALE;
        h[2] = hv[2] + fmaf(cg1, b3.x, am[2]) / MSG_SCALE; h[3] = hv[3] + fmaf(cg1, b3.y, am[3]) / MSG_SCALE;
    }
    layernorm(h, c_n1g, c_n1b, red); // (x is free: every warp passed the barrier inside after its W3 reads)
    NM_STAMP(6);
    store_tile(x, h);
    // ---- FFN 128 -> 512: this CTA's hidden block q
    cp_async_wait<NM_SLOTS - 5>();                   // chunks 1..4
    __syncthreads();                                 // + x
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) mma_chunk(am, ac, x + 32 * s, ring + (1 + s) * WCHUNK_FLOATS, g, t, ncol);
    {
        const float2 bi = c_bin;
        float v[4] = {gelu_erf(am[0] + ac[0] + bi.x), gelu_erf(am[1] + ac[1] + bi.y), gelu_erf(am[2] + ac[2] + bi.x),
                      gelu_erf(am[3] + ac[3] + bi.y)};
        store_tile(hid, v);
    }
    NM_STAMP(7);
    // ---- FFN 512 -> 128: rows 128q.. = the block above
    cp_async_wait<0>();                              // chunks 5..8
    __syncthreads();                                 // + hid; slots 1..4 are free
    if (has_proj)
#pragma unroll
        for (int s = 0; s < 4; ++s) issue(9 + s, 1 + s);
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) mma_chunk(am, ac, hid + 32 * s, ring + (5 + s) * WCHUNK_FLOATS, g, t, ncol);
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i] += ac[i];
    NM_STAMP(8);
    cluster_allreduce(am, xr1);
    NM_STAMP(9);
    {
        const float2 bo = c_bout;
        h[0] += am[0] + bo.x; h[1] += am[1] + bo.y; h[2] += am[2] + bo.x; h[3] += am[3] + bo.y;
    }
    layernorm(h, c_n2g, c_n2b, red + 16 * NM_T);
    NM_STAMP(10);
    {
        const int n0 = node0 + g, n1 = node0 + g + 8;
        const float m0 = n0 < a.N ? a.maskf[n0] : 0.f, m1 = n1 < a.N ? a.maskf[n1] : 0.f;
        h[0] *= m0; h[1] *= m0; h[2] *= m1; h[3] *= m1;
        store_tile(x, h);
        if (q == 0) {
            if (n0 < a.N) {
                *reinterpret_cast<float2*>(a.hV + (size_t)n0 * HID + c0) = make_float2(h[0], h[1]);
                if (a.hV_copy) *reinterpret_cast<float2*>(a.hV_copy + (size_t)n0 * HID + c0) = make_float2(h[0], h[1]);
            }
            if (n1 < a.N) {
                *reinterpret_cast<float2*>(a.hV + (size_t)n1 * HID + c0) = make_float2(h[2], h[3]);
                if (a.hV_copy) *reinterpret_cast<float2*>(a.hV_copy + (size_t)n1 * HID + c0) = make_float2(h[2], h[3]);
            }
        }
    }
    // ---- projections: CTA q computes projection q
    if (has_proj) {
        cp_async_wait<0>();
        __syncthreads();                             // projection chunks + x
#pragma unroll
        for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_chunk(am, ac, x + 32 * s, ring + (1 + s) * WCHUNK_FLOATS, g, t, ncol);
        const ProjSpec ps = a.proj[q];
        const float2 b = c_pb;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int node = node0 + g + 8 * half;
            if (node < a.N) {
                float2 v = make_float2(am[2 * half] + ac[2 * half] + b.x, am[2 * half + 1] + ac[2 * half + 1] + b.y);
                const float2 tv = half ? tv1 : tv0;
                v.x += tv.x; v.y += tv.y;
                *reinterpret_cast<float2*>(ps.out + (size_t)node * HID + c0) = v;
            }
        }
    }
    NM_STAMP(11);
    cp_async_wait<0>();
    cluster.sync();      // no CTA may exit while another can still read its shared memory
    NM_STAMP(12);
}

// ---------------------------------------------------------------------------------------------------------------------
// head, small batches:  x = [dec3 | dec2 | W_s[S]] (384) -> y = LA centre tap . x + b -> (ReLU, Linear) x3 -> affine  => T[N][21]
// (transfer_model.py:82-83, 96-116, 140-155).  The same 4-CTA cluster per 16-residue tile, ONE all-reduce:
//   LA     N-split   CTA q: output columns [96q, 96q+96) over the full K = 384 (every CTA holds the whole x tile)
//   MLP1   K-split   CTA q: exactly those 96 columns of relu(y) are its K slice -> partial [16][64] -> cluster all-reduce
//   MLP2, MLP3       tiny (64 -> 32 -> 21): computed by every CTA from the reduced z1, no exchange; CTA 0 stores
// (The FP32 form this replaces K-split every phase: six all-reduces, 25 us for one 300-residue protein.)
// The CTA's weights (184 KB: LA slice [384 k][96 n] | W1 slice [96][64] | W2 [64][32] | W3 [32][32]) are requested in one
// go before the dependency wait; engine.cu stores them pre-swizzled (16-byte granules XOR (k & 3), see mma_chunk).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int HM_XP = 3 * HID + 4;    // x tile row pitch (388: conflict-free A-fragment loads)
constexpr int HM_YP = 96 + 4, HM_Z1P = 64 + 4, HM_Z2P = 32 + 4;
constexpr int HM_WLA = 384 * 96, HM_W1 = 96 * 64, HM_W2 = 64 * 32, HM_W3 = 32 * 32;
constexpr int HM_WFLOATS = HM_WLA + HM_W1 + HM_W2 + HM_W3;     // 46080 floats per CTA

// acc += A[16][8 nk] . W[8 nk][NC] restricted to this warp's n8 block; A row pitch ap, W row pitch NC (swizzled)
template <int NC>
__device__ __forceinline__ void mma_rows(float (&acc_m)[4], float (&acc_c)[4], const float* __restrict__ A, int ap,
                                         const float* __restrict__ W, int nk8, int g, int t, int ncol) {
#pragma unroll 4
    for (int ks = 0; ks < nk8; ++ks) {
        const int k0 = 8 * ks;
        uint32_t ah[4], al[4], bh[2], bl[2];
        split3(A[g * ap + k0 + t], ah[0], al[0]);
        split3(A[(g + 8) * ap + k0 + t], ah[1], al[1]);
        split3(A[g * ap + k0 + t + 4], ah[2], al[2]);
        split3(A[(g + 8) * ap + k0 + t + 4], ah[3], al[3]);
        const int col = ncol ^ (t << 3);
        split3(W[(k0 + t) * NC + col], bh[0], bl[0]);
        split3(W[(k0 + t + 4) * NC + col], bh[1], bl[1]);
        mma_tf32(acc_c, al, bh[0], bh[1]);
        mma_tf32(acc_c, ah, bl[0], bl[1]);
        mma_tf32(acc_m, ah, bh[0], bh[1]);
    }
}
}  // namespace

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NM_THREADS, 1) head_mma_kernel(HeadArgs a, int64_t wstream) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    float* wla = smem;                               // [384][96]
    float* w1 = wla + HM_WLA;                        // [96][64]
    float* w2 = w1 + HM_W1;                          // [64][32]
    float* w3 = w2 + HM_W2;                          // [32][32]  (columns >= 21 are zero)
    float* x = w3 + HM_W3;                           // [16][HM_XP]
    float* y = x + NM_T * HM_XP;                     // [16][HM_YP]   relu(y) columns [96q, 96q+96)
    float* xr = y + NM_T * HM_YP;                    // [16][64]      this CTA's partial of z1, read by the other three
    float* z1 = xr + NM_T * 64;                      // [16][HM_Z1P]
    float* z2 = z1 + NM_T * HM_Z1P;                  // [16][HM_Z2P]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int node0 = (blockIdx.x >> 2) * NM_T;
    const float* wb = a.weights;

    pdl_trigger();
    {   // the CTA's whole weight block: 11520 16-byte granules, linear copy (already swizzled in global memory)
        const float* src = wb + wstream + (size_t)q * HM_WFLOATS;
        for (int i = threadIdx.x; i < HM_WFLOATS / 4; i += NM_THREADS) cp_async16(smem + i * 4, src + i * 4);
        cp_async_commit();
    }
    // constants of this lane, fetched before the dependency wait
    const int c0 = 8 * warp + 2 * t;                 // this lane's output columns c0, c0 + 1 inside a phase's n8 block
    const float2 c_la = warp < 12 ? __ldg(reinterpret_cast<const float2*>(wb + a.la_b + 96 * q + c0)) : make_float2(0.f, 0.f);
    const float c_sc = __ldg(wb + a.ddg), c_sh = __ldg(wb + a.ddg + 1);
    const float2 c_m2 = warp < 4 ? __ldg(reinterpret_cast<const float2*>(wb + a.m2_b + c0)) : make_float2(0.f, 0.f);
    float2 c_m3 = make_float2(0.f, 0.f);
    if (warp < 3) { c_m3.x = c0 < 21 ? __ldg(wb + a.m3_b + c0) : 0.f; c_m3.y = c0 + 1 < 21 ? __ldg(wb + a.m3_b + c0 + 1) : 0.f; }
    const float c_m1a = __ldg(wb + a.m1_b + (threadIdx.x & 63));                    // z1 reduction: thread <-> column tid % 64
    pdl_wait();
    // ---- x tile: [dec_last | dec_prev | W_s[S]] of the 16 residues
    for (int i = threadIdx.x; i < NM_T * 96; i += NM_THREADS) {
        const int r = i / 96, c4 = (i % 96) * 4, node = node0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (node < a.N) {
            const int blk = c4 >> 7, cc = c4 & 127;
            const float* src = blk == 0 ? a.dec_last + (size_t)node * HID : blk == 1 ? a.dec_prev + (size_t)node * HID
                                        : wb + a.Ws + (size_t)a.S[node] * HID;
            v = *reinterpret_cast<const float4*>(src + cc);
        }
        *reinterpret_cast<float4*>(x + r * HM_XP + c4) = v;
    }
    cp_async_wait<0>();
    __syncthreads();
    float am[4], ac[4];
    // ---- LightAttention centre tap: columns [96q, 96q+96), 12 n8 blocks on warps 0..11, K = 384
    if (warp < 12) {
#pragma unroll
        for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
        mma_rows<96>(am, ac, x, HM_XP, wla, 48, g, t, 8 * warp + g);
        // both_out starts with a ReLU (transfer_model.py:67-71)
        *reinterpret_cast<float2*>(y + g * HM_YP + c0) = make_float2(fmaxf(am[0] + ac[0] + c_la.x, 0.f), fmaxf(am[1] + ac[1] + c_la.y, 0.f));
        *reinterpret_cast<float2*>(y + (g + 8) * HM_YP + c0) = make_float2(fmaxf(am[2] + ac[2] + c_la.x, 0.f), fmaxf(am[3] + ac[3] + c_la.y, 0.f));
    }
    __syncthreads();
    // ---- both_out.1 (384 -> 64): this CTA's K slice = its 96 columns of relu(y); 8 n8 blocks on warps 0..7
    if (warp < 8) {
#pragma unroll
        for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
        mma_rows<64>(am, ac, y, HM_YP, w1, 12, g, t, 8 * warp + g);
        *reinterpret_cast<float2*>(xr + g * 64 + c0) = make_float2(am[0] + ac[0], am[1] + ac[1]);
        *reinterpret_cast<float2*>(xr + (g + 8) * 64 + c0) = make_float2(am[2] + ac[2], am[3] + ac[3]);
    }
    cluster.sync();
    {   // z1 = relu(sum of the four partials, in rank order, + b1): 1024 values, two per thread
        const float* r0 = cluster.map_shared_rank(xr, 0);
        const float* r1 = cluster.map_shared_rank(xr, 1);
        const float* r2 = cluster.map_shared_rank(xr, 2);
        const float* r3 = cluster.map_shared_rank(xr, 3);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = threadIdx.x + j * NM_THREADS, r = i >> 6, c = i & 63;
            z1[r * HM_Z1P + c] = fmaxf(((r0[i] + r1[i]) + (r2[i] + r3[i])) + c_m1a, 0.f);
        }
    }
    __syncthreads();
    // ---- both_out.3 (64 -> 32) and both_out.5 (32 -> 21): every CTA computes them; 4 resp. 3 n8 blocks
    if (warp < 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
        mma_rows<32>(am, ac, z1, HM_Z1P, w2, 8, g, t, 8 * warp + g);
        *reinterpret_cast<float2*>(z2 + g * HM_Z2P + c0) = make_float2(fmaxf(am[0] + ac[0] + c_m2.x, 0.f), fmaxf(am[1] + ac[1] + c_m2.y, 0.f));
        *reinterpret_cast<float2*>(z2 + (g + 8) * HM_Z2P + c0) = make_float2(fmaxf(am[2] + ac[2] + c_m2.x, 0.f), fmaxf(am[3] + ac[3] + c_m2.y, 0.f));
    }
    __syncthreads();
    if (warp < 3 && q == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) am[i] = ac[i] = 0.f;
        mma_rows<32>(am, ac, z2, HM_Z2P, w3, 4, g, t, 8 * warp + g);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int node = node0 + g + 8 * half;
            if (node < a.N) {                        // ddg_out: Linear(1, 1) on each logit (transfer_model.py:110-111)
                if (c0 < 21) a.table[(size_t)node * 21 + c0] = fmaf(am[2 * half] + ac[2 * half] + c_m3.x, c_sc, c_sh);
                if (c0 + 1 < 21) a.table[(size_t)node * 21 + c0 + 1] = fmaf(am[2 * half + 1] + ac[2 * half + 1] + c_m3.y, c_sc, c_sh);
            }
        }
    }
    cluster.sync();      // no CTA may exit while another can still read its shared memory
}

static size_t smem_head_mma() {
    return sizeof(float) * (HM_WFLOATS + NM_T * HM_XP + NM_T * HM_YP + NM_T * 64 + NM_T * HM_Z1P + NM_T * HM_Z2P);
}

static size_t smem_node_mma() {
    return sizeof(float) * (NM_SLOTS * WCHUNK_FLOATS + 2 * NM_T * NM_P + 2 * NM_T * HID + 2 * 2 * 16 * NM_T + NM_T);
}

cudaError_t node_mma_kernels_init() {
    cudaError_t e = cudaFuncSetAttribute(node_update_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_mma());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(head_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_mma());
}

void launch_head_mma(const HeadArgs& a, int64_t wstream, cudaStream_t st) {
    launch_pdl(head_mma_kernel, dim3(4 * ((a.N + NM_T - 1) / NM_T)), dim3(NM_THREADS), smem_head_mma(), st, a, wstream);
}

void launch_node_update_mma(const NodeUpdArgs& a, cudaStream_t st) {
    launch_pdl(node_update_mma_kernel, dim3(4 * ((a.N + NM_T - 1) / NM_T)), dim3(NM_THREADS), smem_node_mma(), st, a);
}

}  // namespace tmpnn
