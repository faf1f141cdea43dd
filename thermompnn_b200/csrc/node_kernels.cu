// Per-residue kernels: the head (FP32 forms) and the node_update dispatcher (its kernels: node_mma_kernels.cu, tc_node_kernels.cu).
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

// ---- register-tiled form of the same arithmetic for the step kernel (NT = 16 or 32) ------------------------------------
// A group's warp covers 16 residues x 32 columns.  step_fma gives every lane one column and all 16 residues: each
// x value is then fetched by a broadcast LDS.128 that moves 16 useful bytes per shared-memory wavefront, and ncu showed
// the kernel bound by exactly that (l1tex data-pipe wavefronts 71 %).  Here lane = (hq = lane / 8, pq = lane % 8) owns
// residues {hq, hq+4, hq+8, ...} (RI of them) and columns [32 w + 4 pq, +4): per four k, 4 weight + 4 x loads of one wavefront
// each feed 64 FMAs (2.5x fewer wavefronts).  Every (residue, column) still runs the identical k-sequential fmaf chain,
// so results equal step_fma's bit for bit.  Row pitches 132 / 516 keep the four residues of a load in different banks.
constexpr int XP = HID + 4;
constexpr int NC_SLOTS = 9;     // resident weight chunks per CTA of the cluster form
constexpr int NC_THREADS = 512;

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
    // this thread's bias values: constants, fetched before the dependency wait (after a cluster.sync, which invalidates
    // L1, each would be an exposed L2 round trip inside the six-phase chain)
    const float c_la0 = __ldg(wb + a.la_b + n), c_la1 = __ldg(wb + a.la_b + HID + n), c_la2 = __ldg(wb + a.la_b + 2 * HID + n);
    const float c_m1 = n < 64 ? __ldg(wb + a.m1_b + n) : 0.f, c_m2 = n < 32 ? __ldg(wb + a.m2_b + n) : 0.f;
    const float c_m3 = n < 21 ? __ldg(wb + a.m3_b + n) : 0.f, c_sc = __ldg(wb + a.ddg), c_sh = __ldg(wb + a.ddg + 1);
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
        const float b = c_la0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * 3 * HID + n] = fmaxf(res[j] + b, 0.f);
    }
    cp_async_wait<4>(); __syncthreads();             // chunks 3..5
    phase(res, 3, 3, x, 3 * HID);
    issue(12, 3); issue(13, 4); cp_async_commit();
    {
        const float b = c_la1;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * 3 * HID + HID + n] = fmaxf(res[j] + b, 0.f);
    }
    cp_async_wait<2>(); __syncthreads();             // chunks 6..8
    phase(res, 3, 6, x, 3 * HID);
    {
        const float b = c_la2;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * 3 * HID + 2 * HID + n] = fmaxf(res[j] + b, 0.f);
    }
    // ---- both_out MLP
    cp_async_wait<1>(); __syncthreads();             // chunks 9..11 + y
    phase(res, 3, 0, y, 3 * HID);
    {
        const float b = c_m1;
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[(g + 4 * j) * HID + n] = fmaxf(res[j] + b, 0.f);   // z1 (columns >= 64 are 0)
    }
    cp_async_wait<0>(); __syncthreads();             // chunks 12, 13 + z1
    phase(res, 1, 3, x, HID);
    {
        const float b = c_m2;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(g + 4 * j) * HID + n] = fmaxf(res[j] + b, 0.f);   // z2 (columns >= 32 are 0)
    }
    __syncthreads();
    phase(res, 1, 4, y, HID);
    if (q == 0 && n < 21) {
        const float b = c_m3, sc = c_sc, sh = c_sh;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int node = node0 + g + 4 * j;
            if (node < a.N) a.table[(size_t)node * 21 + n] = fmaf(res[j] + b, sc, sh);   // ddg_out: Linear(1,1)
        }
    }
    cluster.sync();      // no CTA may exit while another can still read its shared memory
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
    if ((e = cudaFuncSetAttribute(head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head<16>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_cluster<8>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_cluster_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_cluster<12>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head_cluster<16>())) != cudaSuccess) return e;
    return cudaSuccess;
}

constexpr int NODE_MMA_MAX = 593;    // residues: up to here the clusters (one 16-residue tile each, 4 CTAs) fit in ONE wave on 148 SMs

bool node_update_uses_tcgen05(int N) { return N >= NODE_MMA_MAX; }
bool head_uses_tcgen05(int N) { return N >= NODE_MMA_MAX; }     // same boundary: above it the FP32 head is one long step-form wave

void launch_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st) {
    // Small batches: the warp-level tensor-core cluster form (node_mma_kernels.cu), one 16-residue tile per 4-CTA cluster,
    // the whole grid one wave.  Larger batches: the tcgen05 form (tc_node_kernels.cu), 128-residue tiles, persistent
    // over the SMs.  Inside a form a residue's result does not depend on the rest of the batch; the two forms agree to
    // fp32 rounding.
    if (a.N < NODE_MMA_MAX) launch_node_update_mma(a, st);
    else launch_tc_node_update(a, sms, st);
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
