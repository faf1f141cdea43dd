// Per-residue output kernels outside the default fast path:
//
//   head_generic   TransferModel's head for ANY cfg.model.{num_final_layers, lightattn, hidden_dims}
//                  (transfer_model.py:45-73, 96-116): x = [dec_last | dec_prev | ... | W_s[S]] ->
//                  [LightAttention centre tap] -> (ReLU, Linear)* -> ddg_out  => T[N][21].
//                  The shipped configuration (2 final layers, LightAttention, [64, 32]) runs the specialised kernels of
//                  node_kernels.cu instead; this kernel exists so that other checkpoints / configs run on the GPU too.
//   log_probs      ProteinMPNN's own output layer, log_softmax(W_out . h_V + b) (protein_mpnn_utils.py:1275-1276), the
//                  quantity ProteinMPNNBaseline reads (analysis/thermompnn_benchmarking.py:38-65).
//
// Both are small dense per-residue chains (FP32 CUDA cores); weights are read as nn.Linear stores them ([out][in]).
#include "common.cuh"
#include "kernels.h"
#include "pdl.cuh"

namespace tmpnn {

constexpr int HG_THREADS = 256;
constexpr int HG_RT = 8;          // residues per block
constexpr int HG_MAXD = 512;      // widest layer: 128 * 3 + 128 inputs, hidden_dims <= 512

// One warp per output neuron j (strided): lanes split the input dimension, the 8 residues of the tile share every
// weight load; partial sums are combined with a fixed xor-shuffle tree, so the result never depends on the batch.
__global__ void __launch_bounds__(HG_THREADS) head_generic_kernel(HeadGenArgs a) {
    __shared__ float buf[2][HG_RT][HG_MAXD];
    pdl_trigger();
    pdl_wait();
    const int node0 = blockIdx.x * HG_RT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* wb = a.weights;
    const int din = HID * a.num_final + HID;
    for (int i = threadIdx.x; i < HG_RT * din; i += HG_THREADS) {
        const int t = i / din, c = i - t * din, node = node0 + t;
        float v = 0.f;
        if (node < a.N) {
            const int blk = c / HID, cc = c - blk * HID;
            v = blk < a.num_final ? a.dec[blk][(size_t)node * HID + cc] : __ldg(wb + a.Ws + (size_t)a.S[node] * HID + cc);
        }
        buf[0][t][c] = v;
    }
    __syncthreads();
    int cur = 0;
    for (int l = 0; l < a.nlayers; ++l) {
        const int K = a.in_dim[l], M = a.out_dim[l];
        const bool relu_in = l > 0 || a.relu_first;       // both_out = (ReLU, Linear)*; LightAttention takes x as is
        const float* W = wb + a.w[l];
        const float* bias = wb + a.b[l];
        for (int j = warp; j < M; j += HG_THREADS / 32) {
            float acc[HG_RT];
#pragma unroll
            for (int t = 0; t < HG_RT; ++t) acc[t] = 0.f;
            for (int k = lane; k < K; k += 32) {
                const float w = __ldg(W + (size_t)j * K + k);
#pragma unroll
                for (int t = 0; t < HG_RT; ++t) {
                    float x = buf[cur][t][k];
                    if (relu_in) x = fmaxf(x, 0.f);
                    acc[t] = fmaf(x, w, acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < HG_RT; ++t)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
            if (lane < HG_RT) {
                float v = acc[0];
#pragma unroll
                for (int t = 1; t < HG_RT; ++t) v = lane == t ? acc[t] : v;
                buf[cur ^ 1][lane][j] = v + __ldg(bias + j);
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    // ddg_out: Linear(1, 1) on each of the 21 logits (transfer_model.py:110-111)
    const float sc = __ldg(wb + a.ddg), sh = __ldg(wb + a.ddg + 1);
    for (int i = threadIdx.x; i < HG_RT * 21; i += HG_THREADS) {
        const int t = i / 21, c = i - t * 21, node = node0 + t;
        if (node < a.N) a.table[(size_t)node * 21 + c] = fmaf(buf[cur][t][c], sc, sh);
    }
}

// One warp per residue: lane a < 21 owns letter a (its W_out row comes from shared memory), then a warp-wide
// max / sum for the log-softmax, written the way torch does it: x - max - log(sum(exp(x - max))).
constexpr int LP_THREADS = 256;
__global__ void __launch_bounds__(LP_THREADS) log_probs_kernel(LogProbArgs a) {
    __shared__ float Ws[21][HID + 1];
    __shared__ float bs[21];
    pdl_trigger();
    for (int i = threadIdx.x; i < 21 * HID; i += LP_THREADS) Ws[i / HID][i % HID] = __ldg(a.weights + a.Wout + i);
    if (threadIdx.x < 21) bs[threadIdx.x] = __ldg(a.weights + a.bout + threadIdx.x);
    pdl_wait();
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int node = blockIdx.x * (LP_THREADS / 32) + warp; node < a.N; node += gridDim.x * (LP_THREADS / 32)) {
        const float* h = a.hV + (size_t)node * HID;
        const float4 hv = *reinterpret_cast<const float4*>(h + 4 * lane);      // lane l holds h[4l .. 4l+3]
        const int row = lane < 21 ? lane : 20;                                  // lanes >= 21 only serve the shuffles
        float s = 0.f;
        for (int k = 0; k < HID; k += 4) {
            const float4 x = make_float4(__shfl_sync(0xffffffffu, hv.x, k >> 2), __shfl_sync(0xffffffffu, hv.y, k >> 2),
                                         __shfl_sync(0xffffffffu, hv.z, k >> 2), __shfl_sync(0xffffffffu, hv.w, k >> 2));
            s = fmaf(x.x, Ws[row][k], s); s = fmaf(x.y, Ws[row][k + 1], s);
            s = fmaf(x.z, Ws[row][k + 2], s); s = fmaf(x.w, Ws[row][k + 3], s);
        }
        const float logit = lane < 21 ? s + bs[row] : -INFINITY;
        float mx = logit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float e = lane < 21 ? expf(logit - mx) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane < 21) a.out[(size_t)node * 21 + lane] = (logit - mx) - logf(e);
    }
}

// out_m[i][c] = sum_k h[i][k] W_m[k][c]: 16 residues per block, thread = output column (m = tid / 128), FP32.
struct ProjArgs { const float* weights; const float* h; int nmat; int64_t w[3]; float* out[3]; int N; };
constexpr int PJ_RT = 16;
__global__ void __launch_bounds__(384) proj_kernel(ProjArgs a) {
    __shared__ float hs[PJ_RT][HID];
    pdl_trigger();
    pdl_wait();
    const int node0 = blockIdx.x * PJ_RT, m = threadIdx.x >> 7, c = threadIdx.x & 127;
    for (int i = threadIdx.x; i < PJ_RT * HID; i += 384) {
        const int node = node0 + i / HID;
        hs[i / HID][i % HID] = node < a.N ? a.h[(size_t)node * HID + i % HID] : 0.f;
    }
    __syncthreads();
    if (m >= a.nmat) return;
    const float* W = a.weights + a.w[m];
    float acc[PJ_RT];
#pragma unroll
    for (int t = 0; t < PJ_RT; ++t) acc[t] = 0.f;
    for (int k = 0; k < HID; ++k) {
        const float wv = __ldg(W + (size_t)k * HID + c);
#pragma unroll
        for (int t = 0; t < PJ_RT; ++t) acc[t] = fmaf(hs[t][k], wv, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < PJ_RT; ++t)
        if (node0 + t < a.N) a.out[m][(size_t)(node0 + t) * HID + c] = acc[t];
}

cudaError_t head_kernels_init() { return cudaSuccess; }   // static shared memory only

void launch_proj(const float* weights, const float* h, int nmat, const int64_t* w_kn, float* const* out, int N, cudaStream_t st) {
    ProjArgs a{};
    a.weights = weights; a.h = h; a.nmat = nmat; a.N = N;
    for (int m = 0; m < nmat; ++m) { a.w[m] = w_kn[m]; a.out[m] = out[m]; }
    launch_pdl(proj_kernel, dim3((N + PJ_RT - 1) / PJ_RT), dim3(384), 0, st, a);
}

void launch_head_generic(const HeadGenArgs& a, cudaStream_t st) {
    launch_pdl(head_generic_kernel, dim3((a.N + HG_RT - 1) / HG_RT), dim3(HG_THREADS), 0, st, a);
}

void launch_log_probs(const LogProbArgs& a, cudaStream_t st) {
    const int per_block = LP_THREADS / 32;
    int grid = (a.N + per_block - 1) / per_block;
    if (grid > 148 * 8) grid = 148 * 8;
    launch_pdl(log_probs_kernel, dim3(grid), dim3(LP_THREADS), 0, st, a);
}

}  // namespace tmpnn
