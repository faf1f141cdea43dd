/*
 * thermompnn_b200 -- C ABI of the B200-native ThermoMPNN site-saturation ddG path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * Python host layer (thermompnn_b200/engine.py, ctypes) binds exactly these
 * symbols; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * What each entry point replaces in the reference (paths relative to the
 * upstream repo root):
 *
 *   tmpnn_create / tmpnn_destroy   model.cuda() / module lifetime      analysis/custom_inference.py:67-68
 *   tmpnn_weight_*                 the state_dict key/shape contract     thermompnn_benchmarking.py:78-84,
 *   tmpnn_load_weights             load_state_dict of the checkpoint     train_thermompnn.py:30-32, transfer_model.py:17-37
 *   tmpnn_run_batch                TransferModel.forward minus the host  transfer_model.py:75-121
 *                                  parsing: ProteinFeatures + 3 EncLayer protein_mpnn_utils.py:1084-1180,
 *                                  + 3 DecLayer + LightAttention/MLP     794-893, 1222-1277; transfer_model.py:96-155
 *                                  head for all 21 letters per residue
 *   tmpnn_run_batch_ex             the same + ProteinMPNN's own log_probs  protein_mpnn_utils.py:1275-1277; analysis/
 *                                  (ProteinMPNNBaseline)                   thermompnn_benchmarking.py:38-65
 *   tmpnn_model_cfg, tmpnn_cfg_*   cfg.model.{num_final_layers, lightattn, transfer_model.py:45-73
 *                                  hidden_dims}: which head the checkpoint carries
 *   tmpnn_fetch                    forward hooks on prot_mpnn.* used for layer-wise parity (tests only)
 *
 * Inputs are what tied_featurize (protein_mpnn_utils.py:353-605) hands to
 * ProteinMPNN.forward, flattened over a ragged batch of B independent proteins
 * (the reference itself is strictly B=1; a batch here is defined as "identical
 * to running each protein alone"):
 *
 *   offsets[B+1]   int32   node range of protein b is [offsets[b], offsets[b+1]);  N = offsets[B].
 *                          ALWAYS a host pointer (the host sizes grids from it), even with TMPNN_IN_DEVICE
 *   X[N][4][3]     float   backbone N, CA, C, O; zeros where the residue is missing
 *   S[N]           int32   residue letter index in "ACDEFGHIKLMNPQRSTVWYX"
 *   residue_idx[N] int32   100*(chain-1) + position
 *   chain_enc[N]   int32   1-based chain number
 *   mask[N]        uint8   1 = all four backbone atoms present
 *
 * Output:
 *
 *   out_table[N][21] float  T[i][a] = w*z_a(i) + b, the head output for letter a at residue i.
 *                           ddG(i, wt->a) = T[i][a] - T[i][S[i]]  (transfer_model.py:113-114).
 *
 * Error handling: every function returns 0 on success or a negative TMPNN_E_*
 * code; nothing throws across the boundary.  tmpnn_last_error() returns a
 * human-readable message for the most recent failure on that context (or, with
 * ctx == NULL, for the most recent failed tmpnn_create on this thread).
 * Threading: one context per device; a context is not thread-safe; distinct
 * contexts may be used concurrently.  There is no CPU fallback: without a CUDA
 * device tmpnn_create fails with TMPNN_E_CUDA.
 */
#ifndef THERMOMPNN_B200_H
#define THERMOMPNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMPNN_ABI_VERSION 2

#define TMPNN_OK            0
#define TMPNN_E_INVALID    -1   /* bad argument (NULL pointer, negative size, unsorted offsets, ...) */
#define TMPNN_E_CUDA       -2   /* CUDA runtime error, see tmpnn_last_error */
#define TMPNN_E_NOWEIGHTS  -3   /* tmpnn_run_batch before tmpnn_load_weights */
#define TMPNN_E_NOTFOUND   -4   /* unknown tensor name */
#define TMPNN_E_RANGE      -5   /* S / chain / buffer size out of range */

/* tmpnn_run_batch flags */
#define TMPNN_IN_DEVICE    1u   /* X, S, residue_idx, chain_enc, mask are device pointers on the context's device */
#define TMPNN_OUT_DEVICE   2u   /* out_table is a device pointer */
#define TMPNN_NO_SYNC      4u   /* do not synchronise the stream before returning (device in/out only) */
#define TMPNN_UNCONDITIONAL 8u  /* tmpnn_run_batch_ex: out_log_probs = ProteinMPNN.unconditional_probs (no sequence input to the
                                   decoder, protein_mpnn_utils.py:1557-1587) instead of the sequence-conditioned log_probs;
                                   out_table must be NULL */

typedef struct tmpnn_ctx tmpnn_ctx;

int tmpnn_abi_version(void);

/* Lifetime. */
int  tmpnn_create(int device, tmpnn_ctx** out);
void tmpnn_destroy(tmpnn_ctx* ctx);
const char* tmpnn_last_error(const tmpnn_ctx* ctx);

/* Weight contract: the reference checkpoint tensors (names as in
 * models/thermoMPNN_default.pt with the Lightning "model." prefix stripped; the
 * LightAttention feature convolution is taken as its centre tap
 * "light_attention.feature_convolution.weight_center" [384,384]) packed
 * row-major into one float blob at the offsets this table reports. */
int     tmpnn_weight_count(void);
int     tmpnn_weight_info(int index, const char** name, int64_t* offset, int32_t* rows, int32_t* cols);
int64_t tmpnn_weight_blob_floats(void);
int     tmpnn_load_weights(tmpnn_ctx* ctx, const float* host_blob, int64_t n_floats);

/* Model configuration = the head a checkpoint carries (transfer_model.py:45-73: hid_sizes = [128*num_final_layers + 128]
 * + hidden_dims + [21], optional LightAttention in front).  The default (tmpnn_default_cfg: 2 final layers,
 * LightAttention, hidden_dims [64, 32] = models/thermoMPNN_default.pt) runs the specialised head kernels; any other
 * configuration runs a generic FP32 head kernel.  has_head = 0 describes a bare ProteinMPNN (ProteinMPNNBaseline,
 * thermompnn_benchmarking.py:38-44): the weight contract then ends after prot_mpnn.* and only log_probs can be asked for.
 * The tmpnn_weight_* functions above and tmpnn_load_weights are the default-cfg forms of these. */
#define TMPNN_MAX_HIDDEN 8
typedef struct tmpnn_model_cfg {
    int32_t num_final_layers;              /* 0..3 decoder states fed to the head, last layer first */
    int32_t lightattn;                     /* 1: LightAttention (centre tap, see above) before the MLP */
    int32_t n_hidden;                      /* 0..TMPNN_MAX_HIDDEN */
    int32_t hidden_dims[TMPNN_MAX_HIDDEN]; /* each 1..512 */
    int32_t has_head;                      /* 0: ProteinMPNN only */
} tmpnn_model_cfg;
void    tmpnn_default_cfg(tmpnn_model_cfg* cfg);
int     tmpnn_cfg_weight_count(const tmpnn_model_cfg* cfg);                 /* < 0: invalid cfg */
int     tmpnn_cfg_weight_info(const tmpnn_model_cfg* cfg, int index, const char** name, int64_t* offset, int32_t* rows, int32_t* cols);
int64_t tmpnn_cfg_weight_blob_floats(const tmpnn_model_cfg* cfg);
int     tmpnn_load_weights_cfg(tmpnn_ctx* ctx, const tmpnn_model_cfg* cfg, const float* host_blob, int64_t n_floats);

/* The hot path.  stream: a cudaStream_t to run on, or NULL for the context's own stream.
 * With TMPNN_IN_DEVICE the letters S cannot be validated on the host: the first kernel clamps anything outside 0..20
 * to 20 ('X') and raises a flag that turns the call (or, with TMPNN_NO_SYNC, the next synchronising call on the
 * context) into TMPNN_E_RANGE.
 * Calls on one context are ordered against each other even when they use different streams (each call waits for the
 * previous call's completion event before touching the context's buffers). */
int tmpnn_run_batch(tmpnn_ctx* ctx, int32_t B, const int32_t* offsets,
                    const float* X, const int32_t* S, const int32_t* residue_idx,
                    const int32_t* chain_enc, const uint8_t* mask,
                    float* out_table, uint32_t flags, void* stream);

/* tmpnn_run_batch plus ProteinMPNN's own per-residue log-probabilities (protein_mpnn_utils.py:1275-1277):
 * out_log_probs[N][21] = log_softmax(W_out . h_V(last decoder layer) + b).  Either output pointer may be NULL (not
 * wanted); out_table must be NULL for a has_head = 0 model.  TMPNN_OUT_DEVICE applies to both. */
int tmpnn_run_batch_ex(tmpnn_ctx* ctx, int32_t B, const int32_t* offsets,
                       const float* X, const int32_t* S, const int32_t* residue_idx,
                       const int32_t* chain_enc, const uint8_t* mask,
                       float* out_table, float* out_log_probs, uint32_t flags, void* stream);

/* Replay mode: with graphs enabled (default 1) a call whose shape (B, offsets, N), pointers and flags equal the previous
 * call's on this context replays a captured CUDA graph of the whole kernel chain instead of launching 18 kernels. */
int tmpnn_set_graphs(tmpnn_ctx* ctx, int enabled);

/* Introspection for tests and measurement. */
/* Copy an intermediate of the LAST run to host memory.  Names: "E_idx" int32 [N,48] (-1 = unused slot),
 * "D_nb" f32 [N,48], "h_E" f32 [N,48,128] (encoder output), "h_V" f32 [N,128] (node state after the last stage run),
 * "dec0" / "dec1" / "dec2" f32 [N,128] (the decoder states, dec2 = last layer), "table" f32 [N,21],
 * "log_probs" f32 [N,21] (when the last run asked for them). */
int tmpnn_fetch(tmpnn_ctx* ctx, const char* name, void* host_dst, int64_t bytes);
/* Stop after a pipeline stage (for layer-wise parity): 0 = run everything. */
int tmpnn_set_debug_stop(tmpnn_ctx* ctx, int stage);
/* Per-kernel device timing with CUDA events on the launching stream. */
int tmpnn_set_profiling(tmpnn_ctx* ctx, int enabled);
int tmpnn_profile_count(tmpnn_ctx* ctx);
int tmpnn_profile_get(tmpnn_ctx* ctx, int index, const char** kernel_name, double* total_ms, int64_t* launches);
int tmpnn_profile_reset(tmpnn_ctx* ctx);
/* Kernels launched by this context since creation (or the last profile reset). */
int64_t tmpnn_launch_count(const tmpnn_ctx* ctx);
/* How many of the calls since then replayed a captured graph. */
int64_t tmpnn_graph_launch_count(const tmpnn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* THERMOMPNN_B200_H */
