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

#define TMPNN_ABI_VERSION 1

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

/* The hot path.  stream: a cudaStream_t to run on, or NULL for the context's own stream. */
int tmpnn_run_batch(tmpnn_ctx* ctx, int32_t B, const int32_t* offsets,
                    const float* X, const int32_t* S, const int32_t* residue_idx,
                    const int32_t* chain_enc, const uint8_t* mask,
                    float* out_table, uint32_t flags, void* stream);

/* Introspection for tests and measurement. */
/* Copy an intermediate of the LAST run to host memory.  Names: "E_idx" int32 [N,48] (-1 = unused slot),
 * "D_nb" f32 [N,48], "h_E" f32 [N,48,128] (encoder output), "h_V_enc" f32 [N,128],
 * "dec1" / "dec2" f32 [N,128] (the last two decoder states), "table" f32 [N,21]. */
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

#ifdef __cplusplus
}
#endif
#endif /* THERMOMPNN_B200_H */
