// Weight contract and device layout.
//
// RAW: the reference checkpoint tensors (models/thermoMPNN_default.pt, names with the
// Lightning "model." prefix stripped, see train_thermompnn.py:30-32 / transfer_model.py:42-73),
// packed row-major into one host blob; tmpnn_weight_info() reports this table.
//
// DEV: what the kernels read.  nn.Linear weights [out,in] are transposed to [k=in][n=out]
// rows ("KN"), W1/W11 are split into their h_V[i] / h_E / h_V[j] column blocks
// (protein_mpnn_utils.py:819-822: h_EV = [h_V_i | h_E | h_V_j]), and the matrices one kernel
// streams are stored back to back in the order it consumes them.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/thermompnn_b200.h"

namespace tmpnn {

constexpr int KNB = 48;     // neighbour slots per residue (checkpoint num_edges, transfer_model.py:26)
constexpr int HID = 128;    // hidden width
constexpr int NRBF = 16;
constexpr int NPAIR = 25;
constexpr int EDGE_IN = 416;  // 16 positional + 25*16 RBF (protein_mpnn_utils.py:1098)
constexpr float LN_EPS = 1e-5f;
constexpr float MSG_SCALE = 30.0f;  // EncLayer/DecLayer scale (protein_mpnn_utils.py:795,843)

struct RawTensor {
    std::string name;
    int rows, cols;
    int64_t offset;
};

inline tmpnn_model_cfg default_model_cfg() {
    tmpnn_model_cfg c{};
    c.num_final_layers = 2; c.lightattn = 1; c.n_hidden = 2; c.hidden_dims[0] = 64; c.hidden_dims[1] = 32; c.has_head = 1;
    return c;
}
inline bool cfg_valid(const tmpnn_model_cfg& c) {
    if (c.has_head != 0 && c.has_head != 1) return false;
    if (!c.has_head) return true;
    if (c.num_final_layers < 0 || c.num_final_layers > 3 || c.n_hidden < 0 || c.n_hidden > TMPNN_MAX_HIDDEN) return false;
    if (c.lightattn != 0 && c.lightattn != 1) return false;
    for (int i = 0; i < c.n_hidden; ++i)
        if (c.hidden_dims[i] < 1 || c.hidden_dims[i] > 512) return false;
    return true;
}
inline bool cfg_is_default(const tmpnn_model_cfg& c) {
    return c.has_head == 1 && c.num_final_layers == 2 && c.lightattn == 1 && c.n_hidden == 2 && c.hidden_dims[0] == 64 &&
           c.hidden_dims[1] == 32;
}
inline int head_in_dim(const tmpnn_model_cfg& c) { return HID * c.num_final_layers + HID; }   // transfer_model.py:58

inline std::vector<RawTensor> build_raw_table(const tmpnn_model_cfg& cfg) {
    std::vector<RawTensor> t;
    int64_t off = 0;
    auto add = [&](const std::string& n, int r, int c) {
        t.push_back({n, r, c, off});
        off += (int64_t)r * c;
    };
    const std::string pm = "prot_mpnn.";
    add(pm + "features.embeddings.linear.weight", 16, 66);
    add(pm + "features.embeddings.linear.bias", 1, 16);
    add(pm + "features.edge_embedding.weight", 128, 416);
    add(pm + "features.norm_edges.weight", 1, 128);
    add(pm + "features.norm_edges.bias", 1, 128);
    add(pm + "W_e.weight", 128, 128);
    add(pm + "W_e.bias", 1, 128);
    add(pm + "W_s.weight", 21, 128);
    for (int l = 0; l < 3; ++l) {
        std::string p = pm + "encoder_layers." + std::to_string(l) + ".";
        for (const char* n : {"norm1", "norm2", "norm3"}) {
            add(p + n + ".weight", 1, 128);
            add(p + n + ".bias", 1, 128);
        }
        add(p + "W1.weight", 128, 384); add(p + "W1.bias", 1, 128);
        add(p + "W2.weight", 128, 128); add(p + "W2.bias", 1, 128);
        add(p + "W3.weight", 128, 128); add(p + "W3.bias", 1, 128);
        add(p + "W11.weight", 128, 384); add(p + "W11.bias", 1, 128);
        add(p + "W12.weight", 128, 128); add(p + "W12.bias", 1, 128);
        add(p + "W13.weight", 128, 128); add(p + "W13.bias", 1, 128);
        add(p + "dense.W_in.weight", 512, 128); add(p + "dense.W_in.bias", 1, 512);
        add(p + "dense.W_out.weight", 128, 512); add(p + "dense.W_out.bias", 1, 128);
    }
    for (int l = 0; l < 3; ++l) {
        std::string p = pm + "decoder_layers." + std::to_string(l) + ".";
        for (const char* n : {"norm1", "norm2"}) {
            add(p + n + ".weight", 1, 128);
            add(p + n + ".bias", 1, 128);
        }
        add(p + "W1.weight", 128, 512); add(p + "W1.bias", 1, 128);
        add(p + "W2.weight", 128, 128); add(p + "W2.bias", 1, 128);
        add(p + "W3.weight", 128, 128); add(p + "W3.bias", 1, 128);
        add(p + "dense.W_in.weight", 512, 128); add(p + "dense.W_in.bias", 1, 512);
        add(p + "dense.W_out.weight", 128, 512); add(p + "dense.W_out.bias", 1, 128);
    }
    add(pm + "W_out.weight", 21, 128);
    add(pm + "W_out.bias", 1, 21);
    if (!cfg.has_head) return t;
    // head (transfer_model.py:58-73): hid_sizes = [128 F + 128] + hidden_dims + [21]; both_out = (ReLU, Linear) per pair
    const int din = head_in_dim(cfg);
    if (cfg.lightattn) {
        add("light_attention.feature_convolution.weight_center", din, din);
        add("light_attention.feature_convolution.bias", 1, din);
    }
    int prev = din;
    for (int i = 0; i <= cfg.n_hidden; ++i) {
        const int out = i < cfg.n_hidden ? cfg.hidden_dims[i] : 21;
        const std::string p = "both_out." + std::to_string(2 * i + 1);
        add(p + ".weight", out, prev);
        add(p + ".bias", 1, out);
        prev = out;
    }
    add("ddg_out.weight", 1, 1);
    add("ddg_out.bias", 1, 1);
    return t;
}

// ---- device layout (offsets in floats into one allocation) ------------------
struct LayerNormW { int64_t g, b; };

struct NodeUpdW {           // what node_update needs for one Enc/Dec layer
    int64_t stream;         // warp-level form: 16 KB [32 k][128 n] chunks, W3 (4) | W_in column blocks 0..3 (4 each) | W_out (16) | projections (4 each)
    int32_t nchunks;
    int64_t tc_stream;      // tensor-core images for the tcgen05 form: W3 | (W_in block q, W_out block q) x4 | projections
    int32_t tc_chunks;
    int64_t b3, bin, bout;
    LayerNormW n1, n2;
};

struct EncW {
    int64_t tc_msg;         // [W1b | W2] as tensor-core images (8 x 32 KB)
    int64_t b1, b2;
    int64_t tc_edge;        // [W11b | W12 | W13] as tensor-core images: 12 x ([128 n][32 k] hi | lo), 128B-swizzled
    int64_t b11, b12, b13;
    LayerNormW n3;
    NodeUpdW nu;
};

struct DecW {
    int64_t tc_msg;         // [W1b | W2] images  (W1b = h_E block of the 512-wide W1)
    int64_t b1, b2;
    int64_t Tseq;           // [21][128] = W_s . W1c^T : the h_S[j] term folded per letter
    int64_t qw_kn;          // W1[:, 384:512] as [128 k][128 n] fp32: the h_V[j] projection on its own (unconditional_probs)
    NodeUpdW nu;
};

struct DevLayout {
    int64_t pos_table;      // [66][16]  = embeddings.linear.weight^T + bias
    int64_t tc_feat;        // [edge_embedding (416 in) | W_e] as tensor-core images: 13 + 4 chunks of 32 KB
    int64_t be;             // W_e bias
    LayerNormW ne;          // norm_edges
    int64_t rbf_mu;         // [16]
    EncW enc[3];
    DecW dec[3];
    int64_t Ws;             // [21][128] embedding rows
    int64_t Wout, bout;     // ProteinMPNN W_out [21][128] row-major and its bias (log_probs)
    // default-cfg head (specialised kernels):
    int64_t head_mma;       // small-batch head: per cluster rank the LA / both_out slices it keeps in shared memory (4 x 46080 floats)
    int64_t tc_head;        // the same matrices as tensor-core images for the tcgen05 head (51 chunks)
    int64_t la_b, m1_b, m2_b, m3_b;
    int64_t ddg;            // [2] = scale, shift
    // any other cfg (generic head kernel): nn.Linear weights as stored, [out][in] row-major
    int32_t gen_layers;     // LightAttention (if any) + the both_out Linears
    int32_t gen_in[TMPNN_MAX_HIDDEN + 2], gen_out[TMPNN_MAX_HIDDEN + 2];
    int64_t gen_w[TMPNN_MAX_HIDDEN + 2], gen_b[TMPNN_MAX_HIDDEN + 2];
    int64_t zeros;          // [128] zeros (Q for the first encoder layer)
    int64_t total;
};

}  // namespace tmpnn
