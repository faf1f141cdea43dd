// thermompnn_b200: first kernel of a call -- node preparation + the k-NN graph (FP32 / integer SIMT).
//
// Pipeline for a ragged batch of proteins (N residues in total, 48 neighbour slots each):
//
//   knn (first part) backbone atoms + virtual Cb, float mask, sanitised S      protein_mpnn_utils.py:1131-1138
//   knn             48 nearest CA neighbours per residue, ascending            :1101-1109
//   tc_edge_feat    25x16 RBF + positional -> 416->128 -> LN -> W_e  => h_E    :1111-1125, 1142-1179, 1229   (tc_edge_kernels.cu)
//   tc_enc_msg      per edge W1b.h_E + P_i + Q_j -> GELU -> W2 -> GELU, masked sum over neighbours  :819-825
//   node_update     W3 on the sum, /30, LN1, FFN, LN2, mask; projections for the next kernels       :826-832  (node_kernels.cu)
//   tc_edge_upd     h_E = LN3(h_E + W13 gelu(W12 gelu(W11.[..])))              :834-838
//   tc_dec_msg      same skeleton on [h_V_i | m_i (h_E, h_S_j, h_V_j)]         :859-870, 1268-1272
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
                           const uint8_t* __restrict__ mask_u8, const int* __restrict__ S_in, int N, int lmax,
                           float* __restrict__ atoms, float* __restrict__ maskf, int* __restrict__ S_out,
                           int* __restrict__ err_flag, int* __restrict__ E_idx, float* __restrict__ D_nb,
                           float* __restrict__ cnt_enc, float* __restrict__ cnt_dec) {
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
        if (lane == 0) {
            maskf[i] = mi;
            // letters outside 0..20 would index W_s / Tseq out of bounds downstream: clamp, and tell the host
            // (the host-pointer path has already validated S; this covers TMPNN_IN_DEVICE callers)
            const int sv = S_in[i];
            const bool bad = sv < 0 || sv > 20;
            S_out[i] = bad ? 20 : sv;
            if (bad) atomicOr(err_flag, 1);
        }
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
    // Sum of the edge weights of the residue, which every node update needs (b3 * sum_k w_ik): m_i m_j over the selected
    // neighbours for the encoder (mask_attend, protein_mpnn_utils.py:1232-1233), the slot count for the decoder.
    // (Sums of 0/1 values: exact in any order.)
    float c = 0.f;
    for (int e = lane; e < K; e += 32) c += mi * (mask_u8[s + seli[e]] ? 1.0f : 0.0f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) { cnt_enc[i] = c; cnt_dec[i] = (float)K; }
}

cudaError_t kernels_init() {
    return cudaFuncSetAttribute(knn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
int knn_max_len() { return 200 * 1024 / 4 - 2 * KNB; }

void launch_knn(const int* offsets, int B, const float* X, const uint8_t* mask, const int* S_in, int N, int lmax, float* atoms,
                float* maskf, int* S_out, int* err_flag, int* E_idx, float* D_nb, float* cnt_enc, float* cnt_dec, cudaStream_t st) {
    int warps = 8;
    const int rowlen = lmax <= 512 ? 0 : lmax;            // register-resident rows up to 512 residues
    const size_t per_warp = (size_t)(rowlen + 2 * KNB) * 4;
    while (warps > 1 && (size_t)warps * per_warp > 200 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * per_warp;
    const dim3 grid((N + warps - 1) / warps), block(warps * 32);
    if (lmax <= 128) launch_pdl(knn_kernel<4>, grid, block, smem, st, offsets, B, X, mask, S_in, N, lmax, atoms, maskf, S_out, err_flag, E_idx, D_nb, cnt_enc, cnt_dec);
    else if (lmax <= 256) launch_pdl(knn_kernel<8>, grid, block, smem, st, offsets, B, X, mask, S_in, N, lmax, atoms, maskf, S_out, err_flag, E_idx, D_nb, cnt_enc, cnt_dec);
    else if (lmax <= 512) launch_pdl(knn_kernel<16>, grid, block, smem, st, offsets, B, X, mask, S_in, N, lmax, atoms, maskf, S_out, err_flag, E_idx, D_nb, cnt_enc, cnt_dec);
    else launch_pdl(knn_kernel<0>, grid, block, smem, st, offsets, B, X, mask, S_in, N, lmax, atoms, maskf, S_out, err_flag, E_idx, D_nb, cnt_enc, cnt_dec);
}

}  // namespace tmpnn
