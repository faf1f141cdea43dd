// FP32 SIMT generation of the per-edge kernels (6x8 register tiles, cp.async weight ring).  DEVELOPMENT ONLY: compiled
// into the library only with -DTMPNN_DEV (bring-up and error localisation of the tensor-core kernels through
// TMPNN_TC=<mask>); the product library does not contain these kernels.
#ifdef TMPNN_DEV
#include "common.cuh"
#include "kernels.h"

namespace tmpnn {

__constant__ int8_t c_pair_a[NPAIR] = {1, 0, 2, 3, 4, 1, 1, 1, 1, 0, 0, 0, 4, 4, 3, 0, 2, 3, 4, 2, 3, 4, 2, 3, 2};
__constant__ int8_t c_pair_b[NPAIR] = {1, 0, 2, 3, 4, 0, 2, 3, 4, 2, 3, 4, 2, 3, 2, 1, 1, 1, 1, 0, 0, 0, 4, 4, 3};

// ---------------------------------------------------------------------------
// edge_feat
// ---------------------------------------------------------------------------
template <int NPT>
__global__ void __launch_bounds__(128 * NPT)
edge_feat_kernel(const float* __restrict__ atoms, const int* __restrict__ E_idx, const float* __restrict__ D_nb,
                 const int* __restrict__ ridx, const int* __restrict__ cenc, const float* __restrict__ pos_table,
                 const float* __restrict__ rbf_mu, const float* __restrict__ wstream, const float* __restrict__ ln_g,
                 const float* __restrict__ ln_b, const float* __restrict__ b_e, float* __restrict__ hE, int N) {
    constexpr int TM = Tile<NPT>::TM, THREADS = Tile<NPT>::THREADS;
    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;                              // [128][TM]; its head doubles as feat[2][32][TM]
    float* ring = A_s + HID * TM;                   // [3][4096]
    float* red = ring + WRING * WCHUNK_FLOATS;      // [4][TM]
    float* atoms_j = red + 4 * TM;                  // [TM][16]
    float* atoms_i = atoms_j + TM * 16;             // [NPT][16]
    float* dnb = atoms_i + NPT * 16;                // [TM]
    int* posq = reinterpret_cast<int*>(dnb + TM);   // [TM]
    float* mu = reinterpret_cast<float*>(posq + TM);  // [16]

    const ThreadPos tp;
    const int node0 = blockIdx.x * NPT;
    WeightStream<THREADS> ws;
    ws.start(wstream, ring, 17);

    if (threadIdx.x < 16) mu[threadIdx.x] = rbf_mu[threadIdx.x];
    for (int m = threadIdx.x; m < TM; m += THREADS) {
        int node = node0 + m / KNB;
        int self = node < N ? node : N - 1;
        int j = node < N ? E_idx[(size_t)node * KNB + m % KNB] : -1;
        float dn = (j >= 0) ? D_nb[(size_t)node * KNB + m % KNB] : 0.f;
        if (j < 0) j = self;  // unused slot: a well-defined dummy edge (never contributes downstream)
        dnb[m] = dn;
        int off = ridx[self] - ridx[j];
        int q = off + 32;
        q = q < 0 ? 0 : (q > 64 ? 64 : q);
        posq[m] = (cenc[self] == cenc[j]) ? q : 65;  // PositionalEncodings, protein_mpnn_utils.py:904-905
        const float4* src = reinterpret_cast<const float4*>(atoms + (size_t)j * 16);
        float4* dst = reinterpret_cast<float4*>(atoms_j + m * 16);
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = __ldg(src + 3);
    }
    for (int t = threadIdx.x; t < NPT * 16; t += THREADS) {
        int node = node0 + t / 16;
        atoms_i[t] = atoms[(size_t)(node < N ? node : N - 1) * 16 + (t & 15)];
    }
    __syncthreads();

    auto feat_chunk = [&](int c, float* dst) {  // 32 feature columns of every edge row: dst[32][TM]
        for (int item = threadIdx.x; item < 2 * TM; item += THREADS) {
            int h = item / TM, m = item - h * TM;
            float* o = dst + h * 16 * TM + m;
            if (c == 0 && h == 0) {
                const float* pt = pos_table + posq[m] * 16;
#pragma unroll
                for (int t = 0; t < 16; ++t) o[t * TM] = __ldg(pt + t);
            } else {
                int p = (c == 0) ? 0 : 2 * c - 1 + h;
                float D;
                if (p == 0) {
                    D = dnb[m];  // the masked top-k distance itself (protein_mpnn_utils.py:1142)
                } else {
                    const float* a = atoms_i + (m / KNB) * 16 + 3 * c_pair_a[p];
                    const float* b = atoms_j + m * 16 + 3 * c_pair_b[p];
                    float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
                    float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    D = __fsqrt_rn(__fadd_rn(d2, 1e-6f));
                }
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    float z = __fdiv_rn(D - mu[t], 1.25f);
                    o[t * TM] = expf(-(z * z));
                }
            }
        }
    };

    float acc[6][8];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    int g = 0;
    feat_chunk(0, A_s);
    for (int c = 0; c < EDGE_IN / WCHUNK_K; ++c) {
        const float* W = ws.get(g++);
        if (c + 1 < EDGE_IN / WCHUNK_K) feat_chunk(c + 1, A_s + ((c + 1) & 1) * WCHUNK_K * TM);
        fma_chunk<TM>(acc, A_s + (c & 1) * WCHUNK_K * TM, W, tp);
    }
    tile_layernorm<TM>(acc, ln_g, ln_b, red, tp);   // norm_edges; its first barrier also retires the feat buffers
    store_as_A<TM>(A_s, acc, tp);
    gemm_stage128<NPT>(acc, A_s, ws, g, tp);        // W_e
    const int node = node0 + tp.wm;
    if (node < N) {
        float bb[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) bb[c] = __ldg(b_e + tp.col(c));
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float* o = hE + ((size_t)node0 * KNB + tp.m0 + r) * HID + tp.n0;
            *reinterpret_cast<float4*>(o) = make_float4(acc[r][0] + bb[0], acc[r][1] + bb[1], acc[r][2] + bb[2], acc[r][3] + bb[3]);
            *reinterpret_cast<float4*>(o + 16) = make_float4(acc[r][4] + bb[4], acc[r][5] + bb[5], acc[r][6] + bb[6], acc[r][7] + bb[7]);
        }
    }
}

// Load the [TM][128] row-major h_E tile into A_s[k][m] (transposed), zero-filling rows past N.
template <int NPT>
__device__ __forceinline__ void load_hE_tile(float* A_s, const float* __restrict__ hE, int node0, int N) {
    constexpr int TM = Tile<NPT>::TM, THREADS = Tile<NPT>::THREADS;
    const float* base = hE + (size_t)node0 * KNB * HID;
    const int rows_ok = (N - node0) * KNB;  // rows < rows_ok exist
    for (int it = threadIdx.x; it < TM * 16; it += THREADS) {
        int m = it % TM, k8 = it / TM;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (m < rows_ok) {
            v0 = ldg4(base + (size_t)m * HID + k8 * 8);
            v1 = ldg4(base + (size_t)m * HID + k8 * 8 + 4);
        }
        float* o = A_s + (k8 * 8) * TM + m;
        o[0] = v0.x; o[TM] = v0.y; o[2 * TM] = v0.z; o[3 * TM] = v0.w;
        o[4 * TM] = v1.x; o[5 * TM] = v1.y; o[6 * TM] = v1.z; o[7 * TM] = v1.w;
    }
}

// ---------------------------------------------------------------------------
// msg: message MLP up to the second GELU, then the (masked) neighbour sum.
//   ENC: pre = W1b.h_E + P_i + Q_j                    weight of edge = m_i m_j   (protein_mpnn_utils.py:819-825)
//   DEC: pre = P_i + m_i (W1b.h_E + R_j)              weight = 1                  (:863-870, 1268-1272)
// Outputs Ssum[i] = sum_k w g2_ik and cnt[i] = sum_k w; node_update applies W3.
// ---------------------------------------------------------------------------
template <int NPT, bool DEC>
__global__ void __launch_bounds__(128 * NPT)
msg_kernel(const float* __restrict__ hE, const int* __restrict__ E_idx, const float* __restrict__ P, int p_stride,
           const float* __restrict__ Q, int q_stride, const float* __restrict__ maskf,
           const float* __restrict__ wstream, const float* __restrict__ b2, float* __restrict__ Ssum,
           float* __restrict__ cnt, int N) {
    constexpr int TM = Tile<NPT>::TM, THREADS = Tile<NPT>::THREADS;
    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;
    float* ring = A_s + HID * TM;
    int* jidx = reinterpret_cast<int*>(ring + WRING * WCHUNK_FLOATS);  // [TM]

    const ThreadPos tp;
    const int node0 = blockIdx.x * NPT;
    WeightStream<THREADS> ws;
    ws.start(wstream, ring, 8);
    for (int m = threadIdx.x; m < TM; m += THREADS) {
        int node = node0 + m / KNB;
        jidx[m] = node < N ? E_idx[(size_t)node * KNB + m % KNB] : -1;
    }
    load_hE_tile<NPT>(A_s, hE, node0, N);

    const int node = node0 + tp.wm;
    const int nodec = node < N ? node : N - 1;
    float acc[6][8];
    int g = 0;
    gemm_stage128<NPT>(acc, A_s, ws, g, tp);
    {
        float p[8];
        float4 p0 = ldg4(P + (size_t)nodec * p_stride + tp.n0), p1 = ldg4(P + (size_t)nodec * p_stride + tp.n0 + 16);
        p[0] = p0.x; p[1] = p0.y; p[2] = p0.z; p[3] = p0.w; p[4] = p1.x; p[5] = p1.y; p[6] = p1.z; p[7] = p1.w;
        const float mi = DEC ? maskf[nodec] : 1.0f;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            int j = jidx[tp.m0 + r];
            if (j < 0) j = nodec;
            float4 q0 = ldg4(Q + (size_t)j * q_stride + tp.n0), q1 = ldg4(Q + (size_t)j * q_stride + tp.n0 + 16);
            float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v = DEC ? fmaf(mi, acc[r][c] + q[c], p[c]) : (acc[r][c] + p[c] + q[c]);
                acc[r][c] = gelu_erf(v);
            }
        }
    }
    __syncthreads();  // everyone is done reading h_E from A_s
    store_as_A<TM>(A_s, acc, tp);
    gemm_stage128<NPT>(acc, A_s, ws, g, tp);
    {
        float bb[8], sum[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { bb[c] = __ldg(b2 + tp.col(c)); sum[c] = 0.f; }
        const float mi = maskf[nodec];
        float wsum = 0.f;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            int j = jidx[tp.m0 + r];
            float w = (j < 0) ? 0.f : (DEC ? 1.0f : mi * __ldg(maskf + j));
            wsum += w;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float gl = gelu_erf(acc[r][c] + bb[c]);
                sum[c] += (w != 0.f) ? w * gl : 0.f;
            }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
            wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        }
        if (tp.ty == 0 && node < N) {
            float* o = Ssum + (size_t)node * HID + tp.n0;
            *reinterpret_cast<float4*>(o) = make_float4(sum[0], sum[1], sum[2], sum[3]);
            *reinterpret_cast<float4*>(o + 16) = make_float4(sum[4], sum[5], sum[6], sum[7]);
            if (tp.wn == 0 && tp.tx == 0) cnt[node] = wsum;
        }
    }
}

// ---------------------------------------------------------------------------
// edge_upd: h_E <- LN3(h_E + W13 gelu(W12 gelu(W11b.h_E + P_i + Q_j) + b12) + b13)   (protein_mpnn_utils.py:834-838)
// ---------------------------------------------------------------------------
template <int NPT>
__global__ void __launch_bounds__(128 * NPT)
edge_upd_kernel(float* __restrict__ hE, const int* __restrict__ E_idx, const float* __restrict__ P,
                const float* __restrict__ Q, const float* __restrict__ wstream, const float* __restrict__ b12,
                const float* __restrict__ b13, const float* __restrict__ ln_g, const float* __restrict__ ln_b, int N) {
    constexpr int TM = Tile<NPT>::TM, THREADS = Tile<NPT>::THREADS;
    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;
    float* ring = A_s + HID * TM;
    float* red = ring + WRING * WCHUNK_FLOATS;          // [4][TM]
    int* jidx = reinterpret_cast<int*>(red + 4 * TM);   // [TM]

    const ThreadPos tp;
    const int node0 = blockIdx.x * NPT;
    WeightStream<THREADS> ws;
    ws.start(wstream, ring, 12);
    for (int m = threadIdx.x; m < TM; m += THREADS) {
        int node = node0 + m / KNB;
        jidx[m] = node < N ? E_idx[(size_t)node * KNB + m % KNB] : -1;
    }
    load_hE_tile<NPT>(A_s, hE, node0, N);

    const int node = node0 + tp.wm;
    const int nodec = node < N ? node : N - 1;
    float acc[6][8];
    int g = 0;
    gemm_stage128<NPT>(acc, A_s, ws, g, tp);
    {
        float4 p0 = ldg4(P + (size_t)nodec * HID + tp.n0), p1 = ldg4(P + (size_t)nodec * HID + tp.n0 + 16);
        float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            int j = jidx[tp.m0 + r];
            if (j < 0) j = nodec;
            float4 q0 = ldg4(Q + (size_t)j * HID + tp.n0), q1 = ldg4(Q + (size_t)j * HID + tp.n0 + 16);
            float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = gelu_erf(acc[r][c] + p[c] + q[c]);
        }
    }
    __syncthreads();
    store_as_A<TM>(A_s, acc, tp);
    gemm_stage128<NPT>(acc, A_s, ws, g, tp);
    {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float b = __ldg(b12 + tp.col(c));
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[r][c] = gelu_erf(acc[r][c] + b);
        }
    }
    __syncthreads();
    store_as_A<TM>(A_s, acc, tp);
    gemm_stage128<NPT>(acc, A_s, ws, g, tp);
    {
        float bb[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) bb[c] = __ldg(b13 + tp.col(c));
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
            if (node < N) {
                const float* src = hE + ((size_t)node0 * KNB + tp.m0 + r) * HID + tp.n0;
                o0 = *reinterpret_cast<const float4*>(src);
                o1 = *reinterpret_cast<const float4*>(src + 16);
            }
            acc[r][0] += bb[0] + o0.x; acc[r][1] += bb[1] + o0.y; acc[r][2] += bb[2] + o0.z; acc[r][3] += bb[3] + o0.w;
            acc[r][4] += bb[4] + o1.x; acc[r][5] += bb[5] + o1.y; acc[r][6] += bb[6] + o1.z; acc[r][7] += bb[7] + o1.w;
        }
    }
    tile_layernorm<TM>(acc, ln_g, ln_b, red, tp);
    if (node < N) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float* o = hE + ((size_t)node0 * KNB + tp.m0 + r) * HID + tp.n0;
            *reinterpret_cast<float4*>(o) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            *reinterpret_cast<float4*>(o + 16) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <int NPT> static size_t smem_edge_feat() {
    constexpr int TM = Tile<NPT>::TM;
    return sizeof(float) * (HID * TM + WRING * WCHUNK_FLOATS + 4 * TM + TM * 16 + NPT * 16 + TM + TM + 16);
}
template <int NPT> static size_t smem_msg() {
    constexpr int TM = Tile<NPT>::TM;
    return sizeof(float) * (HID * TM + WRING * WCHUNK_FLOATS + TM);
}
template <int NPT> static size_t smem_edge_upd() {
    constexpr int TM = Tile<NPT>::TM;
    return sizeof(float) * (HID * TM + WRING * WCHUNK_FLOATS + 4 * TM + TM);
}

template <typename F>
static cudaError_t set_smem(F* f, size_t bytes) {
    return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t simt_kernels_init() {
    cudaError_t e;
#define TMPNN_INIT_NPT(NPT)                                                                         \
    if ((e = set_smem(edge_feat_kernel<NPT>, smem_edge_feat<NPT>())) != cudaSuccess) return e;       \
    if ((e = set_smem(msg_kernel<NPT, false>, smem_msg<NPT>())) != cudaSuccess) return e;            \
    if ((e = set_smem(msg_kernel<NPT, true>, smem_msg<NPT>())) != cudaSuccess) return e;             \
    if ((e = set_smem(edge_upd_kernel<NPT>, smem_edge_upd<NPT>())) != cudaSuccess) return e;
    TMPNN_INIT_NPT(1) TMPNN_INIT_NPT(2) TMPNN_INIT_NPT(3) TMPNN_INIT_NPT(4)
#undef TMPNN_INIT_NPT
    return cudaSuccess;
}

#define TMPNN_DISPATCH_NPT(npt, CALL)  \
    switch (npt) {                      \
        case 1: { constexpr int NPT = 1; CALL; } break; \
        case 2: { constexpr int NPT = 2; CALL; } break; \
        case 3: { constexpr int NPT = 3; CALL; } break; \
        default: { constexpr int NPT = 4; CALL; } break; \
    }

void launch_edge_feat(int npt, const float* atoms, const int* E_idx, const float* D_nb, const int* ridx,
                      const int* cenc, const float* pos_table, const float* rbf_mu, const float* wstream,
                      const float* ln_g, const float* ln_b, const float* b_e, float* hE, int N, cudaStream_t st) {
    TMPNN_DISPATCH_NPT(npt, (edge_feat_kernel<NPT><<<(N + NPT - 1) / NPT, 128 * NPT, smem_edge_feat<NPT>(), st>>>(
        atoms, E_idx, D_nb, ridx, cenc, pos_table, rbf_mu, wstream, ln_g, ln_b, b_e, hE, N)));
}

void launch_msg(int npt, bool dec, const float* hE, const int* E_idx, const float* P, int p_stride, const float* Q,
                int q_stride, const float* maskf, const float* wstream, const float* b2, float* Ssum, float* cnt,
                int N, cudaStream_t st) {
    if (dec) {
        TMPNN_DISPATCH_NPT(npt, (msg_kernel<NPT, true><<<(N + NPT - 1) / NPT, 128 * NPT, smem_msg<NPT>(), st>>>(
            hE, E_idx, P, p_stride, Q, q_stride, maskf, wstream, b2, Ssum, cnt, N)));
    } else {
        TMPNN_DISPATCH_NPT(npt, (msg_kernel<NPT, false><<<(N + NPT - 1) / NPT, 128 * NPT, smem_msg<NPT>(), st>>>(
            hE, E_idx, P, p_stride, Q, q_stride, maskf, wstream, b2, Ssum, cnt, N)));
    }
}

void launch_edge_upd(int npt, float* hE, const int* E_idx, const float* P, const float* Q, const float* wstream,
                     const float* b12, const float* b13, const float* ln_g, const float* ln_b, int N,
                     cudaStream_t st) {
    TMPNN_DISPATCH_NPT(npt, (edge_upd_kernel<NPT><<<(N + NPT - 1) / NPT, 128 * NPT, smem_edge_upd<NPT>(), st>>>(
        hE, E_idx, P, Q, wstream, b12, b13, ln_g, ln_b, N)));
}

}  // namespace tmpnn

#endif  // TMPNN_DEV
