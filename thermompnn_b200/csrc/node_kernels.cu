// Per-residue kernels: node_update (tail of every Enc/Dec layer) and head.
//
// These are skinny GEMMs (a few residues x 128..512 x 128) over ~0.8 MB of weights per
// call, so they are bound by how fast one SM can pull weights out of L2, not by FMAs.
// Each block therefore keeps SEVEN 16 KB weight chunks in flight through an 8-slot
// shared-memory ring (cp.async), and 512 threads split every 32-row chunk four ways
// along k (kq = tid/128 owns 8 rows, n = tid%128 owns one output column).
#include "common.cuh"
#include "kernels.h"

namespace tmpnn {

constexpr int NK_THREADS = 512;
constexpr int NK_RING = 8;
constexpr int NK_DIST = NK_RING - 1;

struct NodeStream {
    const float* base;
    float* ring;
    int total;
    __device__ __forceinline__ void issue(int g) {
        if (g < total) {
            const float* src = base + (size_t)g * WCHUNK_FLOATS;
            float* dst = ring + (g % NK_RING) * WCHUNK_FLOATS;
#pragma unroll
            for (int i = 0; i < WCHUNK_FLOATS / 4 / NK_THREADS; ++i) {
                int e = (threadIdx.x + i * NK_THREADS) * 4;
                cp_async16(dst + e, src + e);
            }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void start(const float* b, float* r, int t) {
        base = b; ring = r; total = t;
#pragma unroll
        for (int g = 0; g < NK_DIST; ++g) issue(g);
    }
    __device__ __forceinline__ const float* get(int g) {
        cp_async_wait<NK_DIST - 1>();
        __syncthreads();
        issue(g + NK_DIST);
        return ring + (g % NK_RING) * WCHUNK_FLOATS;
    }
};

// res[j] (residue t = kq + 4j, column n) = sum_k xin[t][k] * W[k][n] over the next `nchunks` chunks.
template <int NT>
__device__ __forceinline__ void node_phase(float (&res)[NT / 4], NodeStream& ws, int& g, int nchunks,
                                           const float* xin, int ldx, float* part) {
    const int kq = threadIdx.x >> 7, n = threadIdx.x & 127;
    float acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
        const float* W = ws.get(g++);
        const float* wp = W + (kq * 8) * HID + n;
        const float* xp = xin + c * WCHUNK_K + kq * 8;
#pragma unroll
        for (int kk = 0; kk < 8; kk += 4) {
            float w0 = wp[kk * HID], w1 = wp[(kk + 1) * HID], w2 = wp[(kk + 2) * HID], w3 = wp[(kk + 3) * HID];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float4 xv = *reinterpret_cast<const float4*>(xp + t * ldx + kk);
                acc[t] = fmaf(xv.x, w0, acc[t]);
                acc[t] = fmaf(xv.y, w1, acc[t]);
                acc[t] = fmaf(xv.z, w2, acc[t]);
                acc[t] = fmaf(xv.w, w3, acc[t]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) part[(kq * NT + t) * HID + n] = acc[t];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NT / 4; ++j) {
        int t = kq + 4 * j;
        res[j] = (part[t * HID + n] + part[(NT + t) * HID + n]) + (part[(2 * NT + t) * HID + n] + part[(3 * NT + t) * HID + n]);
    }
}

// LayerNorm over the 128 columns held by the four warps that share kq.  Block-wide barriers inside.
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
// Stream order: W3 | W_in column blocks 0..3 | W_out | projections.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NK_THREADS, 1) node_update_kernel(NodeUpdArgs a) {
    constexpr int NJ = NT / 4;
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                              // [8][4096]
    float* x = ring + NK_RING * WCHUNK_FLOATS;       // [NT][128]
    float* hid = x + NT * HID;                       // [NT][512]
    float* part = hid + NT * 4 * HID;                // [4][NT][128]
    float* red = part + 4 * NT * HID;                // [16][NJ]
    const int kq = threadIdx.x >> 7, n = threadIdx.x & 127;
    const int node0 = blockIdx.x * NT;
    const float* wb = a.weights;
    NodeStream ws;
    ws.start(wb + a.w.stream, ring, a.w.nchunks);
    float* cnt_s = red + 16 * NJ;                      // [NT]
    if (a.Spart) {
        // neighbour sum = the three 16-row partials of the residue, added in order; weight count from E_idx
        for (int i = threadIdx.x; i < NT * HID; i += NK_THREADS) {
            const int node = node0 + i / HID, c = i % HID;
            float s = 0.f;
            if (node < a.N) {
                const float* sp = a.Spart + (size_t)node * 3 * HID + c;
                s = (sp[0] + sp[HID]) + sp[2 * HID];
            }
            x[i] = s;
        }
        if (threadIdx.x < NT * 32) {                   // one warp per residue
            const int t = threadIdx.x >> 5, lane = threadIdx.x & 31, node = node0 + t;
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
            x[i] = node < a.N ? a.Ssum[(size_t)node0 * HID + i] : 0.f;
        }
        if (threadIdx.x < NT) cnt_s[threadIdx.x] = (node0 + threadIdx.x < a.N) ? a.cnt[node0 + threadIdx.x] : 0.f;
    }
    int g = 0;
    float res[NJ], h[NJ];
    node_phase<NT>(res, ws, g, 4, x, HID, part);
    {
        const float b3 = __ldg(wb + a.w.b3 + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int node = node0 + kq + 4 * j;
            float c = cnt_s[kq + 4 * j];
            float hv = node < a.N ? a.hV[(size_t)node * HID + n] : 0.f;
            h[j] = hv + fmaf(c, b3, res[j]) / MSG_SCALE;
        }
    }
    group_layernorm<NJ>(h, wb + a.w.n1.g, wb + a.w.n1.b, red);
#pragma unroll
    for (int j = 0; j < NJ; ++j) x[(kq + 4 * j) * HID + n] = h[j];
    for (int blk = 0; blk < 4; ++blk) {
        node_phase<NT>(res, ws, g, 4, x, HID, part);
        const float bi = __ldg(wb + a.w.bin + blk * HID + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) hid[(kq + 4 * j) * 4 * HID + blk * HID + n] = gelu_erf(res[j] + bi);
    }
    node_phase<NT>(res, ws, g, 16, hid, 4 * HID, part);
    {
        const float bo = __ldg(wb + a.w.bout + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) h[j] += res[j] + bo;
    }
    group_layernorm<NJ>(h, wb + a.w.n2.g, wb + a.w.n2.b, red);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        int t = kq + 4 * j, node = node0 + t;
        float m = node < a.N ? a.maskf[node] : 0.f;
        h[j] *= m;
        x[t * HID + n] = h[j];
        if (node < a.N) {
            a.hV[(size_t)node * HID + n] = h[j];
            if (a.hV_copy) a.hV_copy[(size_t)node * HID + n] = h[j];
        }
    }
    for (int p = 0; p < a.nproj; ++p) {
        node_phase<NT>(res, ws, g, 4, x, HID, part);
        const ProjSpec ps = a.proj[p];
        const float b = ps.bias >= 0 ? __ldg(wb + ps.bias + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int node = node0 + kq + 4 * j;
            if (node < a.N) {
                float v = res[j] + b;
                if (ps.table >= 0) v += __ldg(wb + ps.table + (size_t)a.S[node] * HID + n);
                ps.out[(size_t)node * HID + n] = v;
            }
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// head: x = [dec3 | dec2 | W_s[S]] -> LightAttention centre tap -> (ReLU, Linear) x3 -> affine => T[N][21]
// (transfer_model.py:82-83, 96-116, 140-155).  Stream: LA column blocks 0..2 (12 chunks each) |
// both_out.1 padded to 128 columns (12) | both_out.3 padded (2) | both_out.5 padded (1).
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NK_THREADS, 1) head_kernel(HeadArgs a) {
    constexpr int NJ = NT / 4;
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;
    float* x = ring + NK_RING * WCHUNK_FLOATS;   // [NT][384]; later z1 [NT][128]
    float* y = x + NT * 3 * HID;                 // [NT][384]; later z2 [NT][128]
    float* part = y + NT * 3 * HID;              // [4][NT][128]
    const int kq = threadIdx.x >> 7, n = threadIdx.x & 127;
    const int node0 = blockIdx.x * NT;
    const float* wb = a.weights;
    NodeStream ws;
    ws.start(wb + a.stream, ring, 51);
    for (int i = threadIdx.x; i < NT * HID; i += NK_THREADS) {
        int t = i / HID, c = i % HID, node = node0 + t;
        bool ok = node < a.N;
        x[t * 3 * HID + c] = ok ? a.dec_last[(size_t)node * HID + c] : 0.f;
        x[t * 3 * HID + HID + c] = ok ? a.dec_prev[(size_t)node * HID + c] : 0.f;
        x[t * 3 * HID + 2 * HID + c] = ok ? __ldg(wb + a.Ws + (size_t)a.S[node] * HID + c) : 0.f;
    }
    int g = 0;
    float res[NJ];
    for (int blk = 0; blk < 3; ++blk) {
        node_phase<NT>(res, ws, g, 12, x, 3 * HID, part);
        const float b = __ldg(wb + a.la_b + blk * HID + n);
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(kq + 4 * j) * 3 * HID + blk * HID + n] = fmaxf(res[j] + b, 0.f);
    }
    node_phase<NT>(res, ws, g, 12, y, 3 * HID, part);
    {
        const float b = n < 64 ? __ldg(wb + a.m1_b + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[(kq + 4 * j) * HID + n] = fmaxf(res[j] + b, 0.f);   // z1 (columns >= 64 are 0)
    }
    node_phase<NT>(res, ws, g, 2, x, HID, part);
    {
        const float b = n < 32 ? __ldg(wb + a.m2_b + n) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) y[(kq + 4 * j) * HID + n] = fmaxf(res[j] + b, 0.f);   // z2
    }
    node_phase<NT>(res, ws, g, 1, y, HID, part);
    if (n < 21) {
        const float b = __ldg(wb + a.m3_b + n), sc = __ldg(wb + a.ddg), sh = __ldg(wb + a.ddg + 1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int node = node0 + kq + 4 * j;
            if (node < a.N) a.table[(size_t)node * 21 + n] = fmaf(res[j] + b, sc, sh);   // ddg_out: Linear(1,1)
        }
    }
    cp_async_wait<0>();
}

template <int NT> static size_t smem_node_update() {
    return sizeof(float) * (NK_RING * WCHUNK_FLOATS + NT * HID + NT * 4 * HID + 4 * NT * HID + 16 * (NT / 4) + NT);
}
template <int NT> static size_t smem_head() {
    return sizeof(float) * (NK_RING * WCHUNK_FLOATS + 2 * NT * 3 * HID + 4 * NT * HID);
}

cudaError_t node_kernels_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(node_update_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update<8>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(node_update_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node_update<16>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head<8>())) != cudaSuccess) return e;
    return cudaSuccess;
}

void launch_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st) {
    if (a.N > 16 * sms) {
        node_update_kernel<16><<<(a.N + 15) / 16, NK_THREADS, smem_node_update<16>(), st>>>(a);
    } else {
        node_update_kernel<8><<<(a.N + 7) / 8, NK_THREADS, smem_node_update<8>(), st>>>(a);
    }
}

void launch_head(const HeadArgs& a, cudaStream_t st) {
    head_kernel<8><<<(a.N + 7) / 8, NK_THREADS, smem_head<8>(), st>>>(a);
}

}  // namespace tmpnn
