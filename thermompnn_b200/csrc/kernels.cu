// thermompnn_b200 CUDA kernels (sm_100a), FP32 SIMT generation.
//
// Pipeline for a ragged batch of proteins (N residues in total, 48 neighbour slots each):
//
//   knn (first part) backbone atoms + virtual Cb, residue -> protein range      protein_mpnn_utils.py:1131-1138
//   knn             48 nearest CA neighbours per residue, ascending            :1101-1109
//   edge_feat       25x16 RBF + positional -> 416->128 -> LN -> W_e  => h_E    :1111-1125, 1142-1179, 1229
//   msg<ENC>        per edge W1b.h_E + P_i + Q_j -> GELU -> W2 -> GELU, masked sum over neighbours  :819-825
//   node_update     W3 on the sum, /30, LN1, FFN, LN2, mask; projections for the next kernels       :826-832
//   edge_upd        h_E = LN3(h_E + W13 gelu(W12 gelu(W11.[..])))              :834-838
//   msg<DEC>        same skeleton on [h_V_i | m_i (h_E, h_S_j, h_V_j)]         :859-870, 1268-1272
//   head            [dec3|dec2|h_S] -> LightAttention centre tap -> MLP -> 21 letters  transfer_model.py:96-116
//
// Algebra used (all exact in real arithmetic, validated against the oracle to <3e-5 in fp32):
//   W1.[h_V_i | h_E_ik | h_V_j] = W1a.h_V_i + W1b.h_E_ik + W1c.h_V_j   -> the two h_V terms are per-residue
//   projections (P_i, Q_j) computed once by node_update and gathered per edge;
//   sum_k w_ik (W3 g_ik + b3) = W3 (sum_k w_ik g_ik) + b3 sum_k w_ik    -> W3 moves after the neighbour sum;
//   W1c.W_s[S_j] is a 21-row table per decoder layer.
#include "common.cuh"
#include "kernels.h"
#include "pdl.cuh"

#include <cfloat>

namespace tmpnn {

// Atom order inside the 16-float atom record: N, Ca, C, O, Cb.
// Pair order = column order of edge_embedding (protein_mpnn_utils.py:1142-1167); (A of i, B of j).
__constant__ int8_t c_pair_a[NPAIR] = {1, 0, 2, 3, 4, 1, 1, 1, 1, 0, 0, 0, 4, 4, 3, 0, 2, 3, 4, 2, 3, 4, 2, 3, 2};
__constant__ int8_t c_pair_b[NPAIR] = {1, 0, 2, 3, 4, 0, 2, 3, 4, 2, 3, 4, 2, 3, 2, 1, 1, 1, 1, 0, 0, 0, 4, 4, 3};

// ---------------------------------------------------------------------------
// knn (+ node preparation): one warp per residue; the row of masked distances lives in shared memory.  The K-th smallest distance is
// found by a bitwise (radix) search on the float pattern (all values are >= 0, so uint order == float order): 31 rounds
// of "count keys below the trial value" with a warp-wide integer reduction.  The K winners (ties at the threshold: lowest
// index first) are compacted with ballots and ordered by (distance, index) with a 48-element rank sort, which gives
// exactly torch.topk's ascending output with self first.
// ---------------------------------------------------------------------------
// NPL > 0: proteins of at most 32 NPL residues -- the row of distances lives in registers (NPL keys per lane), so the 31
// selection rounds touch no memory and all the row's loads are in flight at once.  NPL = 0: any length, row in smem.
template <int NPL>
__global__ void knn_kernel(const int* __restrict__ offsets, int B, const float* __restrict__ X,
                           const uint8_t* __restrict__ mask_u8, int N, int lmax, float* __restrict__ atoms,
                           float* __restrict__ maskf, int* __restrict__ E_idx, float* __restrict__ D_nb) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_row[];
    const int warps = blockDim.x >> 5;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * warps + w;
    if (i >= N) return;  // no block-level barriers below
    const int rowlen = NPL > 0 ? 0 : lmax;
    float* d = s_row + (size_t)w * (rowlen + 2 * KNB);
    float* selv = d + rowlen;                             // [48] selected distances
    int* seli = reinterpret_cast<int*>(selv + KNB);       // [48] selected local indices
    // ---- this residue's protein: offsets[b] <= i < offsets[b+1]  (all lanes probe the sorted offsets at once)
    int nle = 0;
    for (int b = lane; b <= B; b += 32) nle += (offsets[b] <= i) ? 1 : 0;
    nle = __reduce_add_sync(0xffffffffu, nle);
    const int s = offsets[nle - 1], L = offsets[nle] - s;
    // ---- node preparation (what the other kernels read): N, CA, C, O and the virtual C-beta (protein_mpnn_utils.py:1133-1136);
    // lane d < 12 holds coordinate d of the residue
    const float xd = X[(size_t)i * 12 + (lane < 12 ? lane : 0)];
    float x9[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) x9[q] = __shfl_sync(0xffffffffu, xd, q);
    const float xi = x9[3], yi = x9[4], zi = x9[5];
    const float mi = mask_u8[i] ? 1.0f : 0.0f;
    if (lane < 16) {
        float v = xd;
        if (lane >= 12) {
            float b[3], cc[3], a[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) { b[q] = __fsub_rn(x9[3 + q], x9[q]); cc[q] = __fsub_rn(x9[6 + q], x9[3 + q]); }
            a[0] = __fsub_rn(__fmul_rn(b[1], cc[2]), __fmul_rn(b[2], cc[1]));
            a[1] = __fsub_rn(__fmul_rn(b[2], cc[0]), __fmul_rn(b[0], cc[2]));
            a[2] = __fsub_rn(__fmul_rn(b[0], cc[1]), __fmul_rn(b[1], cc[0]));
            const int q = lane - 12;
            v = 0.f;
            if (q < 3) {
                const float aq = q == 0 ? a[0] : (q == 1 ? a[1] : a[2]), bq = q == 0 ? b[0] : (q == 1 ? b[1] : b[2]);
                const float cq = q == 0 ? cc[0] : (q == 1 ? cc[1] : cc[2]), caq = q == 0 ? x9[3] : (q == 1 ? x9[4] : x9[5]);
                float t = __fadd_rn(__fmul_rn(-0.58273431f, aq), __fmul_rn(0.56802827f, bq));
                t = __fsub_rn(t, __fmul_rn(0.54067466f, cq));
                v = __fadd_rn(t, caq);
            }
        }
        atoms[(size_t)i * 16 + lane] = v;
        if (lane == 0) maskf[i] = mi;
    }
    // ---- masked CA-CA distances of the row (inputs read directly: other warps are still writing atoms / maskf)
    auto dist = [&](int j, float& m2) -> float {
        const float* aj = X + (size_t)(s + j) * 12 + 3;
        float dx = __fsub_rn(aj[0], xi), dy = __fsub_rn(aj[1], yi), dz = __fsub_rn(aj[2], zi);
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        m2 = mi * (mask_u8[s + j] ? 1.0f : 0.0f);
        return __fmul_rn(m2, __fsqrt_rn(__fadd_rn(d2, 1e-6f)));
    };
    const int K = L < KNB ? L : KNB;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int out_pos = 0, eq_seen = 0;
    // one 32-wide slice of the row: take keys below T, and the first need_eq keys equal to T in index order
    auto compact = [&](uint32_t key, int j, uint32_t T, int need_eq) {
        const bool is_lt = key < T, is_eq = key == T;
        const uint32_t b_eq = __ballot_sync(0xffffffffu, is_eq);
        const bool take = is_lt || (is_eq && (eq_seen + __popc(b_eq & lt_mask)) < need_eq);
        const uint32_t b_take = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int pos = out_pos + __popc(b_take & lt_mask);
            selv[pos] = __uint_as_float(key);
            seli[pos] = j;
        }
        out_pos += __popc(b_take);
        eq_seen += __popc(b_eq);
    };
    if (NPL > 0) {
        uint32_t key[NPL > 0 ? NPL : 1];
        float m2r[NPL > 0 ? NPL : 1];
        float dmax = 0.f;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int j = lane + 32 * q;
            float D = 0.f;
            m2r[q] = 1.0f;
            if (j < L) { D = dist(j, m2r[q]); dmax = fmaxf(dmax, D); }
            key[q] = __float_as_uint(D);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
#pragma unroll
        for (int q = 0; q < NPL; ++q) {      // masked pairs pushed to the row maximum (:1105-1106); padding never selected
            const float D = __fadd_rn(__uint_as_float(key[q]), __fmul_rn(1.0f - m2r[q], dmax));
            key[q] = (lane + 32 * q < L) ? __float_as_uint(D) : 0xffffffffu;
        }
        uint32_t T = 0;                       // threshold = K-th smallest key (bitwise search, all keys >= 0)
        for (int bit = 30; bit >= 0; --bit) {
            const uint32_t trial = T | (1u << bit);
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < NPL; ++q) cnt += (key[q] < trial) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt < K) T = trial;
        }
        int n_less = 0;
#pragma unroll
        for (int q = 0; q < NPL; ++q) n_less += (key[q] < T) ? 1 : 0;
        n_less = __reduce_add_sync(0xffffffffu, n_less);
        const int need_eq = K - n_less;
#pragma unroll
        for (int q = 0; q < NPL; ++q)
            if (32 * q < L) compact(key[q], lane + 32 * q, T, need_eq);
    } else {
        float dmax = 0.f;
        for (int j = lane; j < L; j += 32) {
            float m2;
            const float D = dist(j, m2);
            d[j] = D;
            dmax = fmaxf(dmax, D);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        for (int j = lane; j < L; j += 32) {
            float m2 = mi * (mask_u8[s + j] ? 1.0f : 0.0f);
            d[j] = __fadd_rn(d[j], __fmul_rn(1.0f - m2, dmax));      // masked pairs pushed to the row maximum (:1105-1106)
        }
        __syncwarp();
        uint32_t T = 0;
        for (int bit = 30; bit >= 0; --bit) {
            const uint32_t trial = T | (1u << bit);
            int cnt = 0;
            for (int j = lane; j < L; j += 32) cnt += (__float_as_uint(d[j]) < trial) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt < K) T = trial;
        }
        int n_less = 0;
        for (int j = lane; j < L; j += 32) n_less += (__float_as_uint(d[j]) < T) ? 1 : 0;
        n_less = __reduce_add_sync(0xffffffffu, n_less);
        const int need_eq = K - n_less;                       // how many keys equal to T to take, in index order
        for (int base = 0; base < L; base += 32) {
            const int j = base + lane;
            compact(j < L ? __float_as_uint(d[j]) : 0xffffffffu, j, T, need_eq);
        }
    }
    __syncwarp();
    // ---- order by (distance, index): rank sort of K <= 48 entries
    for (int e = lane; e < K; e += 32) {
        const float v = selv[e];
        const int idx = seli[e];
        int rank = 0;
        for (int o = 0; o < K; ++o) {
            const float ov = selv[o];
            const int oi = seli[o];
            rank += (ov < v || (ov == v && oi < idx)) ? 1 : 0;
        }
        E_idx[(size_t)i * KNB + rank] = s + idx;
        D_nb[(size_t)i * KNB + rank] = v;
    }
    for (int r = K + lane; r < KNB; r += 32) {
        E_idx[(size_t)i * KNB + r] = -1;
        D_nb[(size_t)i * KNB + r] = 0.f;
    }
}

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

cudaError_t kernels_init() {
    cudaError_t e;
#define TMPNN_INIT_NPT(NPT)                                                                         \
    if ((e = set_smem(edge_feat_kernel<NPT>, smem_edge_feat<NPT>())) != cudaSuccess) return e;       \
    if ((e = set_smem(msg_kernel<NPT, false>, smem_msg<NPT>())) != cudaSuccess) return e;            \
    if ((e = set_smem(msg_kernel<NPT, true>, smem_msg<NPT>())) != cudaSuccess) return e;             \
    if ((e = set_smem(edge_upd_kernel<NPT>, smem_edge_upd<NPT>())) != cudaSuccess) return e;
    TMPNN_INIT_NPT(1) TMPNN_INIT_NPT(2) TMPNN_INIT_NPT(3) TMPNN_INIT_NPT(4)
#undef TMPNN_INIT_NPT
    return cudaFuncSetAttribute(knn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

int knn_max_len() { return 200 * 1024 / 4 - 2 * KNB; }

void launch_knn(const int* offsets, int B, const float* X, const uint8_t* mask, int N, int lmax, float* atoms, float* maskf,
                int* E_idx, float* D_nb, cudaStream_t st) {
    int warps = 8;
    const int rowlen = lmax <= 512 ? 0 : lmax;            // register-resident rows up to 512 residues
    const size_t per_warp = (size_t)(rowlen + 2 * KNB) * 4;
    while (warps > 1 && (size_t)warps * per_warp > 200 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * per_warp;
    const dim3 grid((N + warps - 1) / warps), block(warps * 32);
    if (lmax <= 128) launch_pdl(knn_kernel<4>, grid, block, smem, st, offsets, B, X, mask, N, lmax, atoms, maskf, E_idx, D_nb);
    else if (lmax <= 256) launch_pdl(knn_kernel<8>, grid, block, smem, st, offsets, B, X, mask, N, lmax, atoms, maskf, E_idx, D_nb);
    else if (lmax <= 512) launch_pdl(knn_kernel<16>, grid, block, smem, st, offsets, B, X, mask, N, lmax, atoms, maskf, E_idx, D_nb);
    else launch_pdl(knn_kernel<0>, grid, block, smem, st, offsets, B, X, mask, N, lmax, atoms, maskf, E_idx, D_nb);
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
