// Host-visible launch interface of the kernel files: knn_kernel.cu, tc_edge_kernels.cu (edge features, messages, edge update),
// node_mma_kernels.cu / tc_node_kernels.cu / tc_head_kernels.cu (per-residue kernels, small / large batches; node_kernels.cu
// picks the form), head_kernels.cu (generic head, log_probs).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "weights.h"

namespace tmpnn {

struct ProjSpec {      // out[i] = h_V[i] . W (+ bias) (+ table[S_i]); W is the next 4 chunks of the layer's stream
    int64_t bias;      // offset or -1
    int64_t table;     // [21][128] offset or -1
    float* out;        // [N][128]
};

struct NodeUpdArgs {
    const float* weights;
    NodeUpdW w;
    const float* Spart;  // [N][3][128] per-16-row-block partial sums (tensor-core message kernel)
    const float* cnt_pre; // [N] sum of the edge weights per residue, computed once per call by the knn kernel
    int64_t tc_img;      // tcgen05 form: weight images in the order that kernel consumes them (engine.cu: Prep::node_upd)
    float* hV;           // [N][128] in/out
    float* hV_copy;      // optional second destination (decoder states kept for the head)
    const float* maskf;  // [N]
    const int* S;        // [N]
    ProjSpec proj[4];
    int nproj;
    int N;
    long long* trace;    // optional (-DTMPNN_DEV build + TMPNN_TC_TRACE=1): clock64() phase stamps of block 0
};

struct HeadArgs {
    const float* weights;
    int64_t Ws, la_b, m1_b, m2_b, m3_b, ddg;
    const float* dec_last;  // [N][128] output of the last decoder layer
    const float* dec_prev;  // [N][128] output of the one before
    const int* S;
    float* table;           // [N][21]
    int N;
};

// Generic head (any cfg.model.{num_final_layers, lightattn, hidden_dims}, transfer_model.py:45-73,96-116)
struct HeadGenArgs {
    const float* weights;
    int nlayers;                    // LightAttention centre tap (no ReLU in front) + the both_out Linears (ReLU in front)
    int in_dim[TMPNN_MAX_HIDDEN + 2], out_dim[TMPNN_MAX_HIDDEN + 2];
    int64_t w[TMPNN_MAX_HIDDEN + 2], b[TMPNN_MAX_HIDDEN + 2];
    int relu_first;                 // 0: layer 0 is LightAttention (applied to the raw input)
    int64_t Ws, ddg;
    const float* dec[3];            // decoder states, last layer first; only the first num_final are read
    int num_final;
    const int* S;
    float* table;                   // [N][21]
    int N;
};

// ProteinMPNN's own output layer: log_softmax(W_out . h_V + b)   (protein_mpnn_utils.py:1275-1276)
struct LogProbArgs {
    const float* weights;
    int64_t Wout, bout;
    const float* hV;                // [N][128] last decoder state
    float* out;                     // [N][21]
    int N;
};

cudaError_t kernels_init();
cudaError_t head_kernels_init();
cudaError_t node_kernels_init();
cudaError_t node_mma_kernels_init();
void launch_node_update_mma(const NodeUpdArgs& a, cudaStream_t st);   // warp-level tensor-core cluster form (node_mma_kernels.cu)
cudaError_t tc_node_kernels_init();
void launch_tc_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st);   // tcgen05 form for large batches (tc_node_kernels.cu)
cudaError_t tc_kernels_init();
cudaError_t tc_take_error();   // sticky host-side error of the tensor-core launchers (tensor-map encoding), cleared on read
int knn_max_len();

// first kernel of a call: per-residue atoms (N, CA, C, O, virtual C-beta) and float mask, then the k-NN graph
// (also writes S_out = S clamped to 0..20 and raises *err_flag when it had to clamp)
void launch_knn(const int* offsets, int B, const float* X, const uint8_t* mask, const int* S_in, int N, int lmax, float* atoms,
                float* maskf, int* S_out, int* err_flag, int* E_idx, float* D_nb, float* cnt_enc, float* cnt_dec, cudaStream_t st);
void launch_head_generic(const HeadGenArgs& a, cudaStream_t st);
void launch_log_probs(const LogProbArgs& a, cudaStream_t st);
// out_m[i] = h[i] . W_m for up to three [128 k][128 n] matrices (unconditional_probs: the decoder layers' neighbour projections
// of the ENCODER state, computed once)
void launch_proj(const float* weights, const float* h, int nmat, const int64_t* w_kn, float* const* out, int N, cudaStream_t st);
void launch_tc_edge_upd(float* hE, const int* E_idx, const float* P, const float* Q, const float* maskf, const float* wimg,
                        const float* b12, const float* b13, const float* ln_g, const float* ln_b, int N, int sms, cudaStream_t st,
                        long long* trace = nullptr);
void launch_tc_edge_feat(const float* atoms, const int* E_idx, const float* D_nb, const int* ridx, const int* cenc,
                         const float* pos_table, const float* rbf_mu, const float* wimg, const float* ln_g, const float* ln_b,
                         const float* b_e, float* hE, int N, int sms, cudaStream_t st);
void launch_tc_msg(bool dec, const float* hE, const int* E_idx, const float* P, int p_stride, const float* Q, int q_stride,
                   const float* maskf, const float* wimg, const float* b2, float* Spart, int N, int sms, cudaStream_t st,
                   long long* trace = nullptr);
void launch_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st);
bool node_update_uses_tcgen05(int N);   // which form launch_node_update picks for a batch of N residues
void launch_head_mma(const HeadArgs& a, int64_t wstream, cudaStream_t st);   // small batches: node_mma_kernels.cu
bool head_uses_tcgen05(int N);           // large batches: tc_head_kernels.cu
cudaError_t tc_head_kernels_init();
void launch_tc_head(const HeadArgs& a, int64_t tc_img, int sms, cudaStream_t st);

}  // namespace tmpnn
