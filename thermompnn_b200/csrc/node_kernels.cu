// Per-residue kernels: node_update (tail of every Enc/Dec layer) and head.
//
// These are skinny GEMM chains (a few residues x 128..512 x 128) over ~0.8 MB of weights per call: latency- and
// L2->SM-stream-bound, not FLOP-bound.  A block has 512 threads = four groups of 128 (q = tid/128, n = tid%128 = output
// column).  Work proceeds in STEPS of four 16 KB weight chunks ([32 k][128 n]) -- one chunk per group -- so that the
// groups run four different matrices (FFN column blocks, projections) or four K-slices of one matrix (W3, W_out) side
// by side: 13 steps per call instead of 52 sequential chunks.  The next step's 64 KB is prefetched (cp.async) under the
// current step; one block barrier per step.  The host lays the weights out step-major, group-minor (engine.cu).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "pdl.cuh"

namespace tmpnn {

constexpr int NK_THREADS = 512;
constexpr int STEP_FLOATS = 4 * WCHUNK_FLOATS;   // four chunks = 64 KB

struct StepStream {
    const float* base;
    float* ring;      // 2 x STEP_FLOATS
    int total, s;
    __device__ __forceinline__ void issue(int step) {
        if (step < total) {
            const float* src = base + (size_t)step * STEP_FLOATS;
            float* dst = ring + (step & 1) * STEP_FLOATS;
#pragma unroll
            for (int i = 0; i < STEP_FLOATS / 4 / NK_THREADS; ++i) {
                const int e = (threadIdx.x + i * NK_THREADS) * 4;
                cp_async16(dst + e, src + e);
            }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void start(const float* b, float* r, int t) {
        base = b; ring = r; total = t; s = 0;
        issue(0);
    }
    // Block-wide.  Returns this step's four chunks; publishes every shared-memory write made before the call.
    __device__ __forceinline__ const float* next() {
        cp_async_wait<0>();
        __syncthreads();          // step s has landed; everyone is done with step s-1 (whose half is reused below)
        issue(s + 1);
        const float* p = ring + (s & 1) * STEP_FLOATS;
        ++s;
        return p;
    }
};

// acc[t] += sum_{kk<32} x[t][kk] * W[kk][n]      (x: this group's 32-column slice of the input rows)
template <int NT>
__device__ __forceinline__ void step_fma(float (&acc)[NT], const float* __restrict__ W, const float* __restrict__ x, int ldx, int n) {
#pragma unroll
    for (int kk = 0; kk < WCHUNK_K; kk += 4) {
        const float w0 = W[kk * HID + n], w1 = W[(kk + 1) * HID + n], w2 = W[(kk + 2) * HID + n], w3 = W[(kk + 3) * HID + n];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const float4 xv = *reinterpret_cast<const float4*>(x + t * ldx + kk);
            acc[t] = fmaf(xv.x, w0, acc[t]);
            acc[t] = fmaf(xv.y, w1, acc[t]);
            acc[t] = fmaf(xv.z, w2, acc[t]);
            acc[t] = fmaf(xv.w, w3, acc[t]);
        }
    }
}

// One matrix, K split four ways: group q takes rows [(q*steps + s)*32, +32) in step s.  Result: res[j] for residue
// t = q + 4j, column n (the four partials are added in a fixed order).
template <int NT>
__device__ __forceinline__ void phase_ksplit(float (&res)[NT / 4], StepStream& ws, int steps, const float* xin, int ldx, float* part) {
    const int q = threadIdx.x >> 7, n = threadIdx.x & 127;
    float acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    for (int s = 0; s < steps; ++s) {
        const float* W = ws.next() + q * WCHUNK_FLOATS;
        step_fma<NT>(acc, W, xin + (q * steps + s) * WCHUNK_K, ldx, n);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) part[(q * NT + t) * HID + n] = acc[t];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NT / 4; ++j) {
        const int t = q + 4 * j;
        res[j] = (part[t * HID + n] + part[(NT + t) * HID + n]) + (part[(2 * NT + t) * HID + n] + part[(3 * NT + t) * HID + n]);
    }
}

// Four matrices side by side (K = 128 each, four steps): group q computes matrix q for all NT residues.
template <int NT>
__device__ __forceinline__ void phase_4mat(float (&acc)[NT], StepStream& ws, const float* xin, int ldx) {
    const int q = threadIdx.x >> 7, n = threadIdx.x & 127;
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    for (int s = 0; s < 4; ++s) {
        const float* W = ws.next() + q * WCHUNK_FLOATS;
        step_fma<NT>(acc, W, xin + s * WCHUNK_K, ldx, n);
    }
}

// ---- register-tiled form of the same arithmetic for the step kernel (NT = 16 or 32) ------------------------------------
// A group's warp covers 16 residues x 32 columns.  step_fma gives every lane one column and all 16 residues: each
// x value is then fetched by a broadcast LDS.128 that moves 16 useful bytes per shared-memory wavefront, and ncu showed
// the kernel bound by exactly that (l1tex data-pipe wavefronts 71 %).  Here lane = (hq = lane / 8, pq = lane % 8) owns
// residues {hq, hq+4, hq+8, ...} (RI of them) and columns [32 w + 4 pq, +4): per four k, 4 weight + 4 x loads of one wavefront
// each feed 64 FMAs (2.5x fewer wavefronts).  Every (residue, column) still runs the identical k-sequential fmaf chain,
// so results equal step_fma's bit for bit.  Row pitches 132 / 516 keep the four residues of a load in different banks.
constexpr int XP = HID + 4, HIDP = 4 * HID + 4;

template <int RI>
__device__ __forceinline__ void tile_fma(float (&acc)[RI][4], const float* __restrict__ W, const float* __restrict__ x, int ldx,
                                         int hq, int c0) {
#pragma unroll
    for (int kk = 0; kk < WCHUNK_K; kk += 4) {
        const float4 w0 = *reinterpret_cast<const float4*>(W + kk * HID + c0);
        const float4 w1 = *reinterpret_cast<const float4*>(W + (kk + 1) * HID + c0);
        const float4 w2 = *reinterpret_cast<const float4*>(W + (kk + 2) * HID + c0);
        const float4 w3 = *reinterpret_cast<const float4*>(W + (kk + 3) * HID + c0);
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            const float4 xv = *reinterpret_cast<const float4*>(x + (hq + 4 * i) * ldx + kk);
            acc[i][0] = fmaf(xv.x, w0.x, acc[i][0]); acc[i][0] = fmaf(xv.y, w1.x, acc[i][0]);
            acc[i][0] = fmaf(xv.z, w2.x, acc[i][0]); acc[i][0] = fmaf(xv.w, w3.x, acc[i][0]);
            acc[i][1] = fmaf(xv.x, w0.y, acc[i][1]); acc[i][1] = fmaf(xv.y, w1.y, acc[i][1]);
            acc[i][1] = fmaf(xv.z, w2.y, acc[i][1]); acc[i][1] = fmaf(xv.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(xv.x, w0.z, acc[i][2]); acc[i][2] = fmaf(xv.y, w1.z, acc[i][2]);
            acc[i][2] = fmaf(xv.z, w2.z, acc[i][2]); acc[i][2] = fmaf(xv.w, w3.z, acc[i][2]);
            acc[i][3] = fmaf(xv.x, w0.w, acc[i][3]); acc[i][3] = fmaf(xv.y, w1.w, acc[i][3]);
            acc[i][3] = fmaf(xv.z, w2.w, acc[i][3]); acc[i][3] = fmaf(xv.w, w3.w, acc[i][3]);
        }
    }
}

// phase_ksplit with the tiled inner loop (NT = 4 RI residues): same partials, same combine order.
// `part` may alias the input rows (the FFN hidden layer): a barrier separates the last read from the first write.
template <int RI>
__device__ __forceinline__ void phase_ksplit_t(float (&res)[RI], StepStream& ws, int steps, const float* xin, int ldx, float* part) {
    constexpr int NT = 4 * RI;
    const int q = threadIdx.x >> 7, n = threadIdx.x & 127, lane = n & 31;
    const int hq = lane >> 3, c0 = (n & 96) + 4 * (lane & 7);
    float acc[RI][4];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    for (int s = 0; s < steps; ++s) {
        const float* W = ws.next() + q * WCHUNK_FLOATS;
        tile_fma<RI>(acc, W, xin + (q * steps + s) * WCHUNK_K, ldx, hq, c0);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RI; ++i)
        *reinterpret_cast<float4*>(part + (q * NT + hq + 4 * i) * HID + c0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RI; ++j) {
        const int t = q + 4 * j;
        res[j] = (part[t * HID + n] + part[(NT + t) * HID + n]) + (part[(2 * NT + t) * HID + n] + part[(3 * NT + t) * HID + n]);
    }
}

// phase_4mat with the tiled inner loop: acc[i][c] = (matrix q . x)[residue hq + 4 i][column c0 + c]
template <int RI>
__device__ __forceinline__ void phase_4mat_t(float (&acc)[RI][4], StepStream& ws, const float* xin, int ldx, int hq, int c0) {
    const int q = threadIdx.x >> 7;
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    for (int s = 0; s < 4; ++s) {
        const float* W = ws.next() + q * WCHUNK_FLOATS;
        tile_fma<RI>(acc, W, xin + s * WCHUNK_K, ldx, hq, c0);
    }
}

// LayerNorm over the 128 columns held by the four warps of a group.  Block-wide barriers inside.
template <int NJ>
__device__ __forceinline__ void group_layernorm(float (&v)[NJ], const float* __restrict__ g, const float* __restrict__ b,
                                                float* red /*[16][NJ]*/) {
    const int kq = threadIdx.x >> 7, n = threadIdx.x & 127, wg = n >> 5, lane = n & 31;
    float mean[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float s = v[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[(kq * 4 + wg) * NJ + j] = s;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NJ; ++j)
        mean[j] = ((red[(kq * 4) * NJ + j] + red[(kq * 4 + 1) * NJ + j]) + (red[(kq * 4 + 2) * NJ + j] + red[(kq * 4 + 3) * NJ + j])) * (1.0f / HID);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float d = v[j] - mean[j];
        float s = d * d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[(kq * 4 + wg) * NJ + j] = s;
    }
    __syncthreads();
    const float gg = __ldg(g + n), bb = __ldg(b + n);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float var = ((red[(kq * 4) * NJ + j] + red[(kq * 4 + 1) * NJ + j]) + (red[(kq * 4 + 2) * NJ + j] + red[(kq * 4 + 3) * NJ + j])) * (1.0f / HID);
        v[j] = (v[j] - mean[j]) * (1.0f / sqrtf(var + LN_EPS)) * gg + bb;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// node_update:  dh = (W3.Ssum + cnt*b3)/30 ; h = LN1(h_V + dh) ; h = LN2(h + FFN(h)) ; h *= mask
// (protein_mpnn_utils.py:826-832, 871-880), then the projections the next kernels gather.
//
// A tile of NT residues is handled by a CLUSTER of four CTAs.  CTA q loads only chunk q of every step (a quarter of
// the 0.8 MB) and keeps it resident (nine chunks requested at once, before the dependency wait; projections reuse slots):
//   W3      K-split   CTA q: input columns [32q, 32q+32)            -> partial [NT][128]  -> cluster all-reduce
//   W_in    4 mats    CTA q: hidden block q (columns [128q, 128q+128))                    -> stays local ...
//   W_out   K-split   CTA q: rows [128q, 128q+128) = exactly the hidden block it produced -> partial -> all-reduce
//   proj    4 mats    CTA q: projection q (with two projections CTAs 0,1 take one each)
// The two all-reduces exchange [NT][128] partials through distributed shared memory and add them in rank order, so
// the result does not depend on timing.  LayerNorms are computed redundantly by the four CTAs.
// Inside a CTA, thread (kq = tid/128, n = tid%128) takes rows [8kq, 8kq+8) of the 32-row chunk and column n.
// ---------------------------------------------------------------------------
constexpr int NC_SLOTS = 9;     // resident weight chunks per CTA of the cluster form
constexpr int NC_THREADS = 512;

// One CTA of the cluster plays one 128-thread group of node_update_step_kernel: the same per-(residue, column) FMA
// chain (step_fma over the 32 rows of each chunk, chunk after chunk) and the same combine order ((p0+p1)+(p2+p3) over
// the four K-slices), so both forms give bit-identical results and the choice between them (by batch size) is
// invisible.  Inside the CTA the tile's residues are dealt to four thread groups (g = tid/128 owns g, g+4, ...).
template <int NT>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NC_THREADS, 1) node_update_kernel(NodeUpdArgs a) {
    constexpr int NJ = NT / 4;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                              // [NC_SLOTS][4096]: W3 | W_in x4 | W_out x4; projections reuse 1..4
    float* x = ring + NC_SLOTS * WCHUNK_FLOATS;      // [NT][XP]   (padded rows, see tile_fma)
    float* hid = x + NT * XP;                        // [NT][XP]   this CTA's block of the FFN hidden layer
    float* xr0 = hid + NT * XP;                      // [NT][128]  this CTA's partials, read by the other three CTAs
    float* xr1 = xr0 + NT * HID;                     //            (one buffer per all-reduce: no second cluster barrier)
    float* red = xr1 + NT * HID;                     // [16][NJ]
    float* cnt_s = red + 16 * NJ;                    // [NT]
    const int g = threadIdx.x >> 7, n = threadIdx.x & 127;
    // (Measured alternatives for the GEMM phases, both slower than one column x NJ residues per thread on all 16
    // warps: 4 residues x 32 columns per warp with 1 x 4 thread tiles, +30 %; four warps holding all NT residues,
    // +120 %.  At a dozen residues per CTA the phases are bound by per-warp latency, so more warps win.)
    const int node0 = (blockIdx.x >> 2) * NT;
    const float* wb = a.weights;
#define NU_STAMP(i_) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[i_] = clock64(); } while (0)
    NU_STAMP(0);

    // This CTA's quarter of the weight stream (chunk q of every step) stays resident: nine chunks are requested up
    // front -- before pdl_wait(), they are constants -- as nine cp.async groups, so the steps of a phase run without
    // barriers or waits between them.  The four projection chunks take slots 1..4 once the FFN input layer is done.
    const bool has_proj = a.nproj > 0 && q < a.nproj;
    auto issue = [&](int k, int slot) {                // chunk k of this CTA -> slot; always one commit group
        const float* src = wb + a.w.stream + (size_t)(4 * k + q) * WCHUNK_FLOATS;
        float* dst = ring + slot * WCHUNK_FLOATS;
#pragma unroll
        for (int i = 0; i < WCHUNK_FLOATS / 4 / NC_THREADS; ++i) {
            const int e = (threadIdx.x + i * NC_THREADS) * 4;
            cp_async16(dst + e, src + e);
        }
        cp_async_commit();
    };
    pdl_trigger();
#pragma unroll
    for (int k = 0; k < NC_SLOTS; ++k) issue(k, k);
    NU_STAMP(1);
    pdl_wait();                                      // everything below reads earlier kernels' output
    NU_STAMP(2);
    // sum of the four CTAs' partials, in rank order (this thread: residues g + 4j, column n)
    auto cluster_allreduce = [&](float (&v)[NJ], float* xr) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) xr[(g + 4 * j) * HID + n] = v[j];
        cluster.sync();
        const float* r0 = cluster.map_shared_rank(xr, 0);
        const float* r1 = cluster.map_shared_rank(xr, 1);
        const float* r2 = cluster.map_shared_rank(xr, 2);
        const float* r3 = cluster.map_shared_rank(xr, 3);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int o = (g + 4 * j) * HID + n;
            v[j] = (r0[o] + r1[o]) + (r2[o] + r3[o]);
        }
    };
    // h_V rows of this thread's residues: requested now, used after the first all-reduce
    float hv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int node = node0 + g + 4 * j;
        hv[j] = node < a.N ? a.hV[(size_t)node * HID + n] : 0.f;
    }

    if (a.Spart) {
        // neighbour sum = the three 16-row partials of the residue, added in order; weight count from E_idx
        for (int i = threadIdx.x; i < NT * HID; i += NC_THREADS) {
            const int node = node0 + i / HID, c = i % HID;
            float s = 0.f;
            if (node < a.N) {
                const float* sp = a.Spart + (size_t)node * 3 * HID + c;
                s = (sp[0] + sp[HID]) + sp[2 * HID];
            }
            x[(i / HID) * XP + c] = s;
        }
        for (int t = threadIdx.x >> 5; t < NT; t += NC_THREADS / 32) {   // one warp per residue
            const int lane = threadIdx.x & 31, node = node0 + t;
            float c = 0.f;
            if (node < a.N) {
                const float mi = a.maskf[node];
                for (int k = lane; k < KNB; k += 32) {
                    const int j = a.E_idx[(size_t)node * KNB + k];
                    if (j >= 0) c += a.enc ? mi * a.maskf[j] : 1.0f;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) cnt_s[t] = c;
        }
    } else {
        for (int i = threadIdx.x; i < NT * HID; i += NC_THREADS) {
            int node = node0 + i / HID;
            x[(i / HID) * XP + i % HID] = node < a.N ? a.Ssum[(size_t)node0 * HID + i] : 0.f;
        }
        if (threadIdx.x < NT) cnt_s[threadIdx.x] = (node0 + threadIdx.x < a.N) ? a.cnt[node0 + threadIdx.x] : 0.f;
    }

    float acc[NJ], h[NJ];
    NU_STAMP(3);
    // ---- W3 (this CTA: input columns 32q..)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
    cp_async_wait<NC_SLOTS - 1>();                   // chunk 0
    __syncthreads();                                 // ... of every thread; also publishes x / cnt_s
    step_fma<NJ>(acc, ring, x + g * XP + 32 * q, 4 * XP, n);
    NU_STAMP(4);
    cluster_allreduce(acc, xr0);
    NU_STAMP(5);
    {
        const float b3 = __ldg(wb + a.w.b3 + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) h[j] = hv[j] + fmaf(cnt_s[g + 4 * j], b3, acc[j]) / MSG_SCALE;
    }
    group_layernorm<NJ>(h, wb + a.w.n1.g, wb + a.w.n1.b, red);        // ends with a barrier: x is free
    NU_STAMP(6);
#pragma unroll
    for (int j = 0; j < NJ; ++j) x[(g + 4 * j) * XP + n] = h[j];
    // ---- FFN 128 -> 512: this CTA's hidden block q
    cp_async_wait<NC_SLOTS - 5>();                   // chunks 1..4
    __syncthreads();                                 // + x
    {
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) step_fma<NJ>(acc, ring + (1 + s) * WCHUNK_FLOATS, x + g * XP + 32 * s, 4 * XP, n);
        const float bi = __ldg(wb + a.w.bin + q * HID + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) hid[(g + 4 * j) * XP + n] = gelu_erf(acc[j] + bi);
    }
    NU_STAMP(7);
    // ---- FFN 512 -> 128: rows 128q.. = the block above
    cp_async_wait<0>();                              // chunks 5..8
    __syncthreads();                                 // + hid; slots 1..4 are free
    if (has_proj)
#pragma unroll
        for (int s = 0; s < 4; ++s) issue(9 + s, 1 + s);
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) step_fma<NJ>(acc, ring + (5 + s) * WCHUNK_FLOATS, hid + g * XP + 32 * s, 4 * XP, n);
    NU_STAMP(8);
    cluster_allreduce(acc, xr1);
    NU_STAMP(9);
    {
        const float bo = __ldg(wb + a.w.bout + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) h[j] += acc[j] + bo;
    }
    group_layernorm<NJ>(h, wb + a.w.n2.g, wb + a.w.n2.b, red);
    NU_STAMP(10);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int t = g + 4 * j, node = node0 + t;
        const float m = node < a.N ? a.maskf[node] : 0.f;
        h[j] *= m;
        x[t * XP + n] = h[j];
        if (node < a.N && q == 0) {
            a.hV[(size_t)node * HID + n] = h[j];
            if (a.hV_copy) a.hV_copy[(size_t)node * HID + n] = h[j];
        }
    }
    // ---- projections: CTA q computes projection q
    if (has_proj) {
        cp_async_wait<0>();
        __syncthreads();                             // projection chunks + x
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) step_fma<NJ>(acc, ring + (1 + s) * WCHUNK_FLOATS, x + g * XP + 32 * s, 4 * XP, n);
        const ProjSpec ps = a.proj[q];
        const float b = ps.bias >= 0 ? __ldg(wb + ps.bias + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int node = node0 + g + 4 * j;
            if (node < a.N) {
                float v = acc[j] + b;
                if (ps.table >= 0) v += __ldg(wb + ps.table + (size_t)a.S[node] * HID + n);
                ps.out[(size_t)node * HID + n] = v;
            }
        }
    }
    NU_STAMP(11);
    cp_async_wait<0>();
    cluster.sync();      // no CTA may exit while another can still read its shared memory
    NU_STAMP(12);
}

// ---------------------------------------------------------------------------
// node_update, single-CTA form for LARGE batches (many tiles per SM: the per-CTA fixed cost of the cluster form --
// cluster barriers, DSMEM round trips -- would dominate): the four thread groups play the role of the four CTAs.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NK_THREADS, 1) node_update_step_kernel(NodeUpdArgs a) {
    constexpr int NJ = NT / 4;
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                              // [2][4][4096]
    static_assert(NT == 16 || NT == 32, "four residues per lane group and tile row");
    float* x = ring + 2 * STEP_FLOATS;               // [NT][XP]    (padded rows, see tile_fma)
    float* hid = x + NT * XP;                        // [NT][HIDP]
    float* part = hid;                               // [4][NT][128] K-split partials, aliasing the hidden layer
    float* red = hid + NT * HIDP;                    // [16][NJ]
    float* cnt_s = red + 16 * NJ;                    // [NT]
    const int q = threadIdx.x >> 7, n = threadIdx.x & 127;
    const int hq = (n & 31) >> 3, c0 = (n & 96) + 4 * (n & 7);       // this thread's residues {hq + 4 i} and columns [c0, c0 + 4)
    const int node0 = blockIdx.x * NT;
    const float* wb = a.weights;
    StepStream ws;
    pdl_trigger();
    ws.start(wb + a.w.stream, ring, a.w.nchunks / 4);
    pdl_wait();
    if (a.Spart) {
        for (int i = threadIdx.x; i < NT * HID; i += NK_THREADS) {
            const int node = node0 + i / HID, c = i % HID;
            float s = 0.f;
            if (node < a.N) {
                const float* sp = a.Spart + (size_t)node * 3 * HID + c;
                s = (sp[0] + sp[HID]) + sp[2 * HID];
            }
            x[(i / HID) * XP + c] = s;
        }
        for (int t = threadIdx.x >> 5; t < NT; t += NK_THREADS / 32) {   // one warp per residue
            const int lane = threadIdx.x & 31, node = node0 + t;
            float c = 0.f;
            if (node < a.N) {
                const float mi = a.maskf[node];
                for (int k = lane; k < KNB; k += 32) {
                    const int j = a.E_idx[(size_t)node * KNB + k];
                    if (j >= 0) c += a.enc ? mi * a.maskf[j] : 1.0f;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) cnt_s[t] = c;
        }
    } else {
        for (int i = threadIdx.x; i < NT * HID; i += NK_THREADS) {
            int node = node0 + i / HID;
            x[(i / HID) * XP + i % HID] = node < a.N ? a.Ssum[(size_t)node0 * HID + i] : 0.f;
        }
        if (threadIdx.x < NT) cnt_s[threadIdx.x] = (node0 + threadIdx.x < a.N) ? a.cnt[node0 + threadIdx.x] : 0.f;
    }
    float res[NJ], h[NJ];
    phase_ksplit_t<NJ>(res, ws, 1, x, XP, part);                           // W3
    {
        const float b3 = __ldg(wb + a.w.b3 + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int node = node0 + q + 4 * j;
            const float hv = node < a.N ? a.hV[(size_t)node * HID + n] : 0.f;
            h[j] = hv + fmaf(cnt_s[q + 4 * j], b3, res[j]) / MSG_SCALE;
        }
    }
    group_layernorm<NJ>(h, wb + a.w.n1.g, wb + a.w.n1.b, red);
#pragma unroll
    for (int j = 0; j < NJ; ++j) x[(q + 4 * j) * XP + n] = h[j];
    {   // FFN 128 -> 512: group q owns hidden columns [128 q, 128 q + 128)
        float acc[NJ][4];
        phase_4mat_t<NJ>(acc, ws, x, XP, hq, c0);
        const float4 bi = __ldg(reinterpret_cast<const float4*>(wb + a.w.bin + q * HID + c0));
#pragma unroll
        for (int i = 0; i < NJ; ++i)
            *reinterpret_cast<float4*>(hid + (hq + 4 * i) * HIDP + q * HID + c0) =
                make_float4(gelu_erf(acc[i][0] + bi.x), gelu_erf(acc[i][1] + bi.y), gelu_erf(acc[i][2] + bi.z), gelu_erf(acc[i][3] + bi.w));
    }
    phase_ksplit_t<NJ>(res, ws, 4, hid, HIDP, part);                       // FFN 512 -> 128
    {
        const float bo = __ldg(wb + a.w.bout + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) h[j] += res[j] + bo;
    }
    group_layernorm<NJ>(h, wb + a.w.n2.g, wb + a.w.n2.b, red);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int t = q + 4 * j, node = node0 + t;
        const float m = node < a.N ? a.maskf[node] : 0.f;
        h[j] *= m;
        x[t * XP + n] = h[j];
        if (node < a.N) {
            a.hV[(size_t)node * HID + n] = h[j];
            if (a.hV_copy) a.hV_copy[(size_t)node * HID + n] = h[j];
        }
    }
    if (a.nproj > 0) {                                                 // group q computes projection q (idle if q >= nproj)
        float acc[NJ][4];
        phase_4mat_t<NJ>(acc, ws, x, XP, hq, c0);
        if (q < a.nproj) {
            const ProjSpec ps = a.proj[q];
            const float4 b = ps.bias >= 0 ? __ldg(reinterpret_cast<const float4*>(wb + ps.bias + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < NJ; ++i) {
                const int node = node0 + hq + 4 * i;
                if (node < a.N) {
                    float4 v = make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
                    if (ps.table >= 0) {
                        const float4 tv = __ldg(reinterpret_cast<const float4*>(wb + ps.table + (size_t)a.S[node] * HID + c0));
                        v.x += tv.x; v.y += tv.y; v.z += tv.z; v.w += tv.w;
                    }
                    *reinterpret_cast<float4*>(ps.out + (size_t)node * HID + c0) = v;
                }
            }
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// head: x = [dec3 | dec2 | W_s[S]] -> LightAttention centre tap -> (ReLU, Linear) x3 -> affine => T[N][21]
// (transfer_model.py:82-83, 96-116, 140-155).  Steps (all K-split): LA column blocks 0..2 (3 steps each, K = 384) |
// both_out.1 padded to 128 columns (3) | both_out.3 padded to K = 128 (1) | both_out.5 padded (1).
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NK_THREADS, 1) head_kernel(HeadArgs a) {
    constexpr int NJ = NT / 4;
    constexpr int XP3 = 3 * HID + 4;             // padded row pitch of the 384-wide inputs (see tile_fma)
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;
    float* x = ring + 2 * STEP_FLOATS;           // [NT][XP3]; later z1 [NT][XP]
    float* y = x + NT * XP3;                     // [NT][XP3]; later z2 [NT][XP]
    float* part = y + NT * XP3;                  // [4][NT][128]
    const int q = threadIdx.x >> 7, n = threadIdx.x & 127;
    const int node0 = blockIdx.x * NT;
    const float* wb = a.weights;
    StepStream ws;
    pdl_trigger();
    ws.start(wb + a.stream, ring, 14);
    pdl_wait();
    for (int i = threadIdx.x; i < NT * HID; i += NK_THREADS) {
        int t = i / HID, c = i % HID, node = node0 + t;
        bool ok = node < a.N;
        x[t * XP3 + c] = ok ? a.dec_last[(size_t)node * HID + c] : 0.f;
        x[t * XP3 + HID + c] = ok ? a.dec_prev[(size_t)node * HID + c] : 0.f;
        x[t * XP3 + 2 * HID + c] = ok ? __ldg(wb + a.Ws + (size_t)a.S[node] * HID + c) : 0.f;
    }
    float res[NJ];
    for (int blk = 0; blk < 3; ++blk) {
        phase_ksplit_t<NJ>(res, ws, 3, x, XP3, part);
        const float b = __ldg(wb + a.la_b + blk * HID + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(q + 4 * j) * XP3 + blk * HID + n] = fmaxf(res[j] + b, 0.f);
    }
    phase_ksplit_t<NJ>(res, ws, 3, y, XP3, part);
    {
        const float b = n < 64 ? __ldg(wb + a.m1_b + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[(q + 4 * j) * XP + n] = fmaxf(res[j] + b, 0.f);   // z1 (columns >= 64 are 0)
    }
    phase_ksplit_t<NJ>(res, ws, 1, x, XP, part);
    {
        const float b = n < 32 ? __ldg(wb + a.m2_b + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(q + 4 * j) * XP + n] = fmaxf(res[j] + b, 0.f);   // z2 (columns >= 32 are 0)
    }
    phase_ksplit_t<NJ>(res, ws, 1, y, XP, part);
    if (n < 21) {
        const float b = __ldg(wb + a.m3_b + n), sc = __ldg(wb + a.ddg), sh = __ldg(wb + a.ddg + 1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int node = node0 + q + 4 * j;
            if (node < a.N) a.table[(size_t)node * 21 + n] = fmaf(res[j] + b, sc, sh);   // ddg_out: Linear(1,1)
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// head, cluster form for SMALL batches (the step form streams all 0.9 MB of head weights through every block, which
// at a few dozen blocks is one long exposed L2 -> SM stream).  As in node_update_kernel, CTA q of a 4-CTA cluster plays
// group q of head_kernel: chunk q of each of the 14 steps, the same per-(residue, column) chains, partials exchanged
// through distributed shared memory and added in rank order -- bit-identical to the step form.
// Weight chunks: 0..8 (LightAttention) requested up front, before the dependency wait; 9..11 take slots 0..2 once
// block 0 is done, 12..13 slots 3..4 once block 1 is done.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NC_THREADS, 1) head_cluster_kernel(HeadArgs a) {
    constexpr int NJ = NT / 4;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                              // [NC_SLOTS][4096]
    float* x = ring + NC_SLOTS * WCHUNK_FLOATS;      // [NT][384]; later z1 [NT][128]
    float* y = x + NT * 3 * HID;                     // [NT][384]; later z2 [NT][128]
    float* xr0 = y + NT * 3 * HID;                   // [NT][128] partials, one buffer per parity of the all-reduce
    float* xr1 = xr0 + NT * HID;
    const int g = threadIdx.x >> 7, n = threadIdx.x & 127;
    const int node0 = (blockIdx.x >> 2) * NT;
    const float* wb = a.weights;
    auto issue = [&](int k, int slot) {              // chunk q of step k -> slot (no commit)
        const float* src = wb + a.stream + (size_t)(4 * k + q) * WCHUNK_FLOATS;
        float* dst = ring + slot * WCHUNK_FLOATS;
#pragma unroll
        for (int i = 0; i < WCHUNK_FLOATS / 4 / NC_THREADS; ++i) {
            const int e = (threadIdx.x + i * NC_THREADS) * 4;
            cp_async16(dst + e, src + e);
        }
    };
    pdl_trigger();
#pragma unroll
    for (int k = 0; k < NC_SLOTS; ++k) { issue(k, k); cp_async_commit(); }
    pdl_wait();
    for (int i = threadIdx.x; i < NT * HID; i += NC_THREADS) {
        int t = i / HID, c = i % HID, node = node0 + t;
        bool ok = node < a.N;
        x[t * 3 * HID + c] = ok ? a.dec_last[(size_t)node * HID + c] : 0.f;
        x[t * 3 * HID + HID + c] = ok ? a.dec_prev[(size_t)node * HID + c] : 0.f;
        x[t * 3 * HID + 2 * HID + c] = ok ? __ldg(wb + a.Ws + (size_t)a.S[node] * HID + c) : 0.f;
    }
    int ar = 0;
    // K-split phase of `steps` steps on input rows xin (row pitch ldx): this CTA takes rows [(q steps + s) 32, +32)
    auto phase = [&](float (&res)[NJ], int steps, int slot0, const float* xin, int ldx) {
        float acc[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
        for (int s = 0; s < steps; ++s)
            step_fma<NJ>(acc, ring + (slot0 + s) * WCHUNK_FLOATS, xin + g * ldx + (q * steps + s) * WCHUNK_K, 4 * ldx, n);
        float* xr = (ar++ & 1) ? xr1 : xr0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) xr[(g + 4 * j) * HID + n] = acc[j];
        cluster.sync();
        const float* r0 = cluster.map_shared_rank(xr, 0);
        const float* r1 = cluster.map_shared_rank(xr, 1);
        const float* r2 = cluster.map_shared_rank(xr, 2);
        const float* r3 = cluster.map_shared_rank(xr, 3);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int o = (g + 4 * j) * HID + n;
            res[j] = (r0[o] + r1[o]) + (r2[o] + r3[o]);
        }
    };
    float res[NJ];
    // ---- LightAttention centre tap, three column blocks of 128 (K = 384)
    cp_async_wait<6>(); __syncthreads();             // chunks 0..2 + x
    phase(res, 3, 0, x, 3 * HID);
    issue(9, 0); issue(10, 1); issue(11, 2); cp_async_commit();      // slots 0..2 are free (cluster barrier above)
    {
        const float b = __ldg(wb + a.la_b + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * 3 * HID + n] = fmaxf(res[j] + b, 0.f);
    }
    cp_async_wait<4>(); __syncthreads();             // chunks 3..5
    phase(res, 3, 3, x, 3 * HID);
    issue(12, 3); issue(13, 4); cp_async_commit();
    {
        const float b = __ldg(wb + a.la_b + HID + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * 3 * HID + HID + n] = fmaxf(res[j] + b, 0.f);
    }
    cp_async_wait<2>(); __syncthreads();             // chunks 6..8
    phase(res, 3, 6, x, 3 * HID);
    {
        const float b = __ldg(wb + a.la_b + 2 * HID + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * 3 * HID + 2 * HID + n] = fmaxf(res[j] + b, 0.f);
    }
    // ---- both_out MLP
    cp_async_wait<1>(); __syncthreads();             // chunks 9..11 + y
    phase(res, 3, 0, y, 3 * HID);
    {
        const float b = n < 64 ? __ldg(wb + a.m1_b + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[(g + 4 * j) * HID + n] = fmaxf(res[j] + b, 0.f);   // z1 (columns >= 64 are 0)
    }
    cp_async_wait<0>(); __syncthreads();             // chunks 12, 13 + z1
    phase(res, 1, 3, x, HID);
    {
        const float b = n < 32 ? __ldg(wb + a.m2_b + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * HID + n] = fmaxf(res[j] + b, 0.f);   // z2 (columns >= 32 are 0)
    }
    __syncthreads();
    phase(res, 1, 4, y, HID);
    if (q == 0 && n < 21) {
        const float b = __ldg(wb + a.m3_b + n), sc = __ldg(wb + a.ddg), sh = __ldg(wb + a.ddg + 1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int node = node0 + g + 4 * j;
            if (node < a.N) a.table[(size_t)node * 21 + n] = fmaf(res[j] + b, sc, sh);   // ddg_out: Linear(1,1)
        }
    }
    cluster.sync();      // no CTA may exit while another can still read its shared memory
}

template <int NT> static size_t smem_node_update() {
    return sizeof(float) * (NC_SLOTS * WCHUNK_FLOATS + 2 * NT * XP + 2 * NT * HID + 16 * (NT / 4) + NT);
}
template <int NT> static size_t smem_node_update_step() {
    return sizeof(float) * (2 * STEP_FLOATS + NT * XP + NT * HIDP + 16 * (NT / 4) + NT);
}
template <int NT> static size_t smem_head_cluster() {
    return sizeof(float) * (NC_SLOTS * WCHUNK_FLOATS + 2 * NT * 3 * HID + 2 * NT * HID);
}
template <int NT> static size_t smem_head() {
    return sizeof(float) * (2 * STEP_FLOATS + 2 * NT * (3 * HID + 4) + 4 * NT * HID);
}

cudaError_t node_kernels_init() {
    cudaError_t e;
    if ((e = node_mma_kernels_init()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(node_update_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update<8>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(node_update_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update<16>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(node_update_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update<12>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(node_update_step_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update_step<16>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(node_update_step_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update_step<32>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head<16>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_cluster<8>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_cluster_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_cluster<12>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_cluster<16>())) != cudaSuccess) return e;
    return cudaSuccess;
}

constexpr int NODE_MMA_MAX = 593;    // residues: up to here the clusters (one 16-residue tile each, 4 CTAs) fit in ONE wave on 148 SMs

bool node_update_uses_tcgen05(int N) { return N >= NODE_MMA_MAX; }

void launch_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st) {
    // Small batches: a 4-CTA cluster per tile of NT residues, NT chosen so that the whole grid is ONE wave (152 CTAs on
    // 148 SMs would take two).  Large batches: the single-CTA step kernel, 16 or 32 residues per block.
    // All forms are bit-identical per residue, so the choice never shows in the results.
    const int N = a.N;
    // Small and medium batches: the warp-level tensor-core cluster form (node_mma_kernels.cu), one 16-residue tile per
    // 4-CTA cluster.  (The FP32 forms below remain for larger batches and for the dev build's SIMT message kernels.)
    // Larger batches: the tcgen05 form (tc_node_kernels.cu), 128-residue tiles, persistent over the SMs.
    if (a.Spart && N < NODE_MMA_MAX) { launch_node_update_mma(a, st); return; }
    if (a.Spart) { launch_tc_node_update(a, sms, st); return; }
    if (4 * ((N + 7) / 8) <= sms) {
        launch_pdl(node_update_kernel<8>, dim3(4 * ((N + 7) / 8)), dim3(NC_THREADS), smem_node_update<8>(), st, a);
    } else if (4 * ((N + 11) / 12) <= sms) {
        launch_pdl(node_update_kernel<12>, dim3(4 * ((N + 11) / 12)), dim3(NC_THREADS), smem_node_update<12>(), st, a);
    } else if (4 * ((N + 15) / 16) <= sms) {
        launch_pdl(node_update_kernel<16>, dim3(4 * ((N + 15) / 16)), dim3(NC_THREADS), smem_node_update<16>(), st, a);
    } else if ((N + 31) / 32 < 2 * sms) {
        launch_pdl(node_update_step_kernel<16>, dim3((N + 15) / 16), dim3(NK_THREADS), smem_node_update_step<16>(), st, a);
    } else {   // at least two full waves of 32-residue blocks: half the weight traffic per residue
        launch_pdl(node_update_step_kernel<32>, dim3((N + 31) / 32), dim3(NK_THREADS), smem_node_update_step<32>(), st, a);
    }
}

void launch_head(const HeadArgs& a, int sms, cudaStream_t st) {
    const int N = a.N;   // same dispatch rule as launch_node_update; all forms bit-identical per residue
    if (4 * ((N + 7) / 8) <= sms) {
        launch_pdl(head_cluster_kernel<8>, dim3(4 * ((N + 7) / 8)), dim3(NC_THREADS), smem_head_cluster<8>(), st, a);
    } else if (4 * ((N + 11) / 12) <= sms) {
        launch_pdl(head_cluster_kernel<12>, dim3(4 * ((N + 11) / 12)), dim3(NC_THREADS), smem_head_cluster<12>(), st, a);
    } else if (4 * ((N + 15) / 16) <= sms) {
        launch_pdl(head_cluster_kernel<16>, dim3(4 * ((N + 15) / 16)), dim3(NC_THREADS), smem_head_cluster<16>(), st, a);
    } else {
        launch_pdl(head_kernel<16>, dim3((N + 15) / 16), dim3(NK_THREADS), smem_head<16>(), st, a);
    }
}

}  // namespace tmpnn
