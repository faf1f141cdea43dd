// thermompnn_b200 engine: the C ABI declared in include/thermompnn_b200.h.
//
// Owns the device state of one GPU: the prepared weight block, the per-batch
// activation buffers (sized for the largest batch seen), a pinned staging buffer
// for host inputs and one CUDA stream.  No CPU fallback exists: every entry
// point that needs the GPU fails with TMPNN_E_CUDA when there is none.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/thermompnn_b200.h"
#include "kernels.h"
#include "weights.h"

using namespace tmpnn;

namespace {

thread_local std::string g_create_error;

struct ProfEntry {
    std::string name;
    double total_ms = 0.0;
    int64_t launches = 0;
};

struct PendingEvent {
    int entry;
    cudaEvent_t a, b;
};

uint64_t g_realloc_count = 0;   // (per process; only ever compared for equality)

struct DeviceGuard {   // the C ABI must not change the caller's current device (a torch process may be on another GPU)
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        ++g_realloc_count;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct GraphKey {   // everything a captured kernel chain depends on besides the staged input bytes
    int B = -1, N = 0, lmax = 0, stop = 0;
    uint32_t flags = 0;
    const void* ptr[7] = {};
    uint64_t epoch = 0;
    bool operator==(const GraphKey& o) const {
        return B == o.B && N == o.N && lmax == o.lmax && stop == o.stop && flags == o.flags && epoch == o.epoch &&
               std::memcmp(ptr, o.ptr, sizeof(ptr)) == 0;
    }
};

struct tmpnn_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;

    tmpnn_model_cfg cfg = default_model_cfg();
    std::vector<RawTensor> raw;
    DevLayout lay{};
    float* d_weights = nullptr;
    bool have_weights = false;

    // inputs (device mirror of the staging buffer)
    DevBuf<unsigned char> d_in;
    unsigned char* h_stage = nullptr;
    size_t h_stage_cap = 0;
    cudaEvent_t stage_ev = nullptr;   // completion of the last H2D out of h_stage
    bool stage_busy = false;
    // per-node
    DevBuf<float> atoms, maskf, hV, dec_first, dec_prev, dec_last, P, Q, P2, Q2, table, logp, Spart, cnt_enc, cnt_dec;
    DevBuf<int> S_dev;                // S clamped to 0..20 by the first kernel
    // per-edge
    DevBuf<int> E_idx;
    DevBuf<float> D_nb, hE;
    uint64_t buf_epoch = 0;           // bumped whenever a buffer above is reallocated (invalidates the captured graph)

    int* d_err = nullptr;             // sticky device flag: bit 0 = a letter outside 0..20 was clamped
    int* h_err = nullptr;             // pinned mirror, refreshed at the end of every run
    bool clear_err = false;
    cudaEvent_t run_ev = nullptr;     // completion of the last run (orders runs on different streams, fetch)
    bool run_pending = false;

    int last_N = 0;
    float* last_table = nullptr;
    float* last_logp = nullptr;
    int debug_stop = 0;

    // replay: the kernel chain of the last call shape, captured once it repeats
    bool graphs = true;
    GraphKey last_key, graph_key;
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_launches = 0, graph_kernels = 0;

    long long* d_trace = nullptr;     // -DTMPNN_DEV + TMPNN_TC_TRACE=1: phase stamps of enc layer 1's kernels

    bool profiling = false;
    std::vector<ProfEntry> prof;
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    int64_t launches = 0;
};

namespace {

int fail(tmpnn_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, TMPNN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

// ---------------- weight preparation (host) ---------------------------------
struct Prep {
    const float* raw;
    const std::vector<RawTensor>& table;
    std::vector<float> dev;
    const float* find(const std::string& n, int* rows = nullptr, int* cols = nullptr) const {
        for (auto& t : table)
            if (t.name == n) {
                if (rows) *rows = t.rows;
                if (cols) *cols = t.cols;
                return raw + t.offset;
            }
        return nullptr;
    }
    int64_t alloc(int64_t n) {
        int64_t off = (int64_t)dev.size();
        int64_t padded = (n + 255) / 256 * 256;   // keep every block 1024-byte aligned (bulk copies, swizzled images)
        dev.resize(dev.size() + padded, 0.f);
        return off;
    }
    int64_t vec(const std::string& name) {
        int r, c;
        const float* v = find(name, &r, &c);
        int64_t off = alloc((int64_t)r * c);
        std::copy(v, v + (int64_t)r * c, dev.begin() + off);
        return off;
    }
    LayerNormW ln(const std::string& p) { return {vec(p + ".weight"), vec(p + ".bias")}; }
    // One 16 KB chunk [32 k][128 n] of nn.Linear W[out][in]:  chunk[k][n] = W[out0 + n][in0 + k]  (zero outside the matrix)
    int64_t put_chunk(int64_t off, const std::string& name, int out0, int in0) {
        int out, in;
        const float* W = find(name, &out, &in);
        for (int k = 0; k < 32; ++k)
            for (int n = 0; n < HID; ++n)
                dev[off + (int64_t)k * HID + n] = (out0 + n < out && in0 + k < in) ? W[(int64_t)(out0 + n) * in + in0 + k] : 0.f;
        return 4096;
    }
    // node_update weights in STEP order (four chunks per step, one per thread group q; node_kernels.cu):
    //   W3 (K split: group q = rows 32q..)  |  W_in: step s = rows 32s.. of column block q  |
    //   W_out (K split: group q = rows (4q+s)*32..)  |  projections: step s = rows 32s.. of projection q
    //   (two projections: group q = projection q&1, K half q>>1, two steps)
    NodeUpdW node_upd(const std::string& p, const std::vector<std::pair<std::string, int>>& projs) {
        NodeUpdW w;
        // projections: step s = rows 32s.. of projection q (zero chunks pad q >= number of projections)
        const int pchunks = projs.empty() ? 0 : 16;
        w.nchunks = 4 * (1 + 4 + 4) + pchunks;
        w.stream = alloc((int64_t)w.nchunks * 4096);
        int64_t o = w.stream;
        for (int q = 0; q < 4; ++q) o += put_chunk(o, p + "W3.weight", 0, 32 * q);
        for (int s = 0; s < 4; ++s)
            for (int q = 0; q < 4; ++q) o += put_chunk(o, p + "dense.W_in.weight", HID * q, 32 * s);
        for (int s = 0; s < 4; ++s)
            for (int q = 0; q < 4; ++q) o += put_chunk(o, p + "dense.W_out.weight", 0, (4 * q + s) * 32);
        for (int s = 0; s < (pchunks ? 4 : 0); ++s)
            for (int q = 0; q < 4; ++q) {
                if (q < (int)projs.size()) put_chunk(o, projs[q].first, 0, projs[q].second + 32 * s);
                o += 4096;   // (alloc() zero-fills)
            }
        // the same matrices as tensor-core images, in the order the tcgen05 form consumes them (tc_node_kernels.cu):
        //   W3 | for q = 0..3: W_in rows [128q, +128) , W_out columns [128q, +128) | projections   (4 chunks of 32 KB each)
        w.tc_chunks = 4 + 32 + 4 * (int)projs.size();
        w.tc_stream = alloc((int64_t)w.tc_chunks * 8192);
        {
            int64_t t = w.tc_stream;
            put_tc(t, p + "W3.weight", 0); t += 4 * 8192;
            for (int q = 0; q < 4; ++q) {
                put_tc(t, p + "dense.W_in.weight", 0, 4, HID * q); t += 4 * 8192;
                put_tc(t, p + "dense.W_out.weight", HID * q); t += 4 * 8192;
            }
            for (auto& pr : projs) { put_tc(t, pr.first, pr.second); t += 4 * 8192; }
        }
        w.b3 = vec(p + "W3.bias");
        w.n1 = ln(p + "norm1");
        w.n2 = ln(p + "norm2");
        w.bin = vec(p + "dense.W_in.bias");
        w.bout = vec(p + "dense.W_out.bias");
        return w;
    }
    // Tensor-core image of W[:, col0:col0+128] (nn.Linear [out=128][in]): 4 K-chunks of ([128 n][32 k] hi | lo),
    // rows 128 bytes, 16-byte chunks XOR-swizzled by (n & 7); hi/lo are the two tf32 words of the 3xTF32 split.
    static float rna_tf32(float x) {
        uint32_t u;
        std::memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;
        float r;
        std::memcpy(&r, &u, 4);
        return r;
    }
    // (rows / columns outside the matrix are zero: narrow layers are padded to the 128 x 32 chunk shape)
    void put_tc(int64_t off, const std::string& name, int col0, int nchunks = 4, int row0 = 0) {
        int out, in;
        const float* W = find(name, &out, &in);
        for (int c = 0; c < nchunks; ++c)
            for (int n = 0; n < HID; ++n)
                for (int kk = 0; kk < 32; ++kk) {
                    const int row = row0 + n, colk = col0 + c * 32 + kk;
                    const float x = (row < out && colk < in) ? W[(int64_t)row * in + colk] : 0.f;
                    const float hi = rna_tf32(x), lo = rna_tf32(x - hi);
                    const int64_t idx = off + (int64_t)c * 8192 + n * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
                    dev[idx] = hi;
                    dev[idx + 4096] = lo;
                }
    }
};

DevLayout prepare_weights(Prep& P, const tmpnn_model_cfg& cfg) {
    DevLayout L{};
    const std::string pm = "prot_mpnn.";
    {   // positional table: Linear(one_hot(q)) = W[:, q] + b   (protein_mpnn_utils.py:906-907)
        const float* W = P.find(pm + "features.embeddings.linear.weight");
        const float* b = P.find(pm + "features.embeddings.linear.bias");
        L.pos_table = P.alloc(66 * 16);
        for (int q = 0; q < 66; ++q)
            for (int t = 0; t < 16; ++t) P.dev[L.pos_table + q * 16 + t] = W[t * 66 + q] + b[t];
    }
    L.tc_feat = P.alloc(17 * 8192);
    P.put_tc(L.tc_feat, pm + "features.edge_embedding.weight", 0, 13);
    P.put_tc(L.tc_feat + 13 * 8192, pm + "W_e.weight", 0, 4);
    L.be = P.vec(pm + "W_e.bias");
    L.ne = P.ln(pm + "features.norm_edges");
    {   // torch.linspace(2., 22., 16) in fp32 (protein_mpnn_utils.py:1113-1114)
        L.rbf_mu = P.alloc(16);
        const float start = 2.0f, end = 22.0f;
        const float step = (end - start) / 15.0f;
        for (int i = 0; i < 16; ++i)
            P.dev[L.rbf_mu + i] = (i < 8) ? start + step * (float)i : end - step * (float)(15 - i);
    }
    for (int l = 0; l < 3; ++l) {
        const std::string p = pm + "encoder_layers." + std::to_string(l) + ".";
        EncW& e = L.enc[l];
        e.tc_msg = P.alloc(8 * 8192);
        P.put_tc(e.tc_msg, p + "W1.weight", HID);
        P.put_tc(e.tc_msg + 4 * 8192, p + "W2.weight", 0);
        e.b1 = P.vec(p + "W1.bias");
        e.b2 = P.vec(p + "W2.bias");
        e.tc_edge = P.alloc(12 * 8192);
        P.put_tc(e.tc_edge, p + "W11.weight", HID);
        P.put_tc(e.tc_edge + 4 * 8192, p + "W12.weight", 0);
        P.put_tc(e.tc_edge + 8 * 8192, p + "W13.weight", 0);
        e.b11 = P.vec(p + "W11.bias");
        e.b12 = P.vec(p + "W12.bias");
        e.b13 = P.vec(p + "W13.bias");
        e.n3 = P.ln(p + "norm3");
        {   // projections gathered by this layer's edge update and by the next message kernel
            std::vector<std::pair<std::string, int>> pr = {{p + "W11.weight", 0}, {p + "W11.weight", 2 * HID}};
            if (l < 2) {
                const std::string nx = pm + "encoder_layers." + std::to_string(l + 1) + ".";
                pr.push_back({nx + "W1.weight", 0});
                pr.push_back({nx + "W1.weight", 2 * HID});
            } else {
                const std::string nx = pm + "decoder_layers.0.";
                pr.push_back({nx + "W1.weight", 0});
                pr.push_back({nx + "W1.weight", 3 * HID});
            }
            e.nu = P.node_upd(p, pr);
        }
    }
    const float* Ws = P.find(pm + "W_s.weight");
    for (int l = 0; l < 3; ++l) {
        const std::string p = pm + "decoder_layers." + std::to_string(l) + ".";
        DecW& d = L.dec[l];
        d.tc_msg = P.alloc(8 * 8192);
        P.put_tc(d.tc_msg, p + "W1.weight", HID);
        P.put_tc(d.tc_msg + 4 * 8192, p + "W2.weight", 0);
        d.b1 = P.vec(p + "W1.bias");
        d.b2 = P.vec(p + "W2.bias");
        // h_S[j] term: W1[:, 256:384] . W_s[S_j]  -> one row per letter
        const float* W1 = P.find(p + "W1.weight");
        d.qw_kn = P.alloc(HID * HID);
        for (int k = 0; k < HID; ++k)
            for (int n = 0; n < HID; ++n) P.dev[d.qw_kn + (int64_t)k * HID + n] = W1[(int64_t)n * 4 * HID + 3 * HID + k];
        d.Tseq = P.alloc(21 * HID);
        for (int a = 0; a < 21; ++a)
            for (int n = 0; n < HID; ++n) {
                double s = 0.0;
                for (int k = 0; k < HID; ++k) s += (double)Ws[a * HID + k] * (double)W1[(int64_t)n * 4 * HID + 2 * HID + k];
                P.dev[d.Tseq + a * HID + n] = (float)s;
            }
        {
            std::vector<std::pair<std::string, int>> pr;
            if (l < 2) {
                const std::string nx = pm + "decoder_layers." + std::to_string(l + 1) + ".";
                pr = {{nx + "W1.weight", 0}, {nx + "W1.weight", 3 * HID}};
            }
            d.nu = P.node_upd(p, pr);
        }
    }
    L.Ws = P.vec(pm + "W_s.weight");
    L.Wout = P.vec(pm + "W_out.weight");
    L.bout = P.vec(pm + "W_out.bias");
    if (cfg.has_head && cfg_is_default(cfg)) {
        // Small-batch head (node_mma_kernels.cu: head_mma_kernel): per CTA q of the 4-CTA cluster one contiguous block
        //   LA[:, 96q..96q+96) as [384 k][96 n] | both_out.1[:, 96q..) as [96 k][64 n] | both_out.3 [64 k][32 n] | both_out.5 [32 k][32 n]
        // stored as the kernel keeps it in shared memory: element (k, n) at k * NC + (n ^ ((k & 3) << 3)).
        L.head_mma = P.alloc((int64_t)4 * 46080);
        {
            auto put = [&](int64_t off, const std::string& name, int K, int NC, int out0, int in0) {
                int out, in;
                const float* W = P.find(name, &out, &in);
                for (int k = 0; k < K; ++k)
                    for (int n = 0; n < NC; ++n) {
                        const int o = out0 + n, i = in0 + k;
                        P.dev[off + (int64_t)k * NC + (n ^ ((k & 3) << 3))] = (o < out && i < in) ? W[(int64_t)o * in + i] : 0.f;
                    }
            };
            for (int q = 0; q < 4; ++q) {
                int64_t o = L.head_mma + (int64_t)q * 46080;
                put(o, "light_attention.feature_convolution.weight_center", 384, 96, 96 * q, 0); o += 384 * 96;
                put(o, "both_out.1.weight", 96, 64, 0, 96 * q); o += 96 * 64;
                put(o, "both_out.3.weight", 64, 32, 0, 0); o += 64 * 32;
                put(o, "both_out.5.weight", 32, 32, 0, 0);
            }
        }
        // the same as tensor-core images for the tcgen05 head (tc_head_kernels.cu), in its consumption order:
        //   for j = 0..2: LA[j,0], LA[j,1], LA[j,2] (4 chunks each), both_out.1[:, j] (4) | both_out.3 (2) | both_out.5 (1)
        L.tc_head = P.alloc((int64_t)51 * 8192);
        {
            int64_t t = L.tc_head;
            for (int j = 0; j < 3; ++j) {
                for (int k = 0; k < 3; ++k) { P.put_tc(t, "light_attention.feature_convolution.weight_center", HID * k, 4, HID * j); t += 4 * 8192; }
                P.put_tc(t, "both_out.1.weight", HID * j); t += 4 * 8192;
            }
            P.put_tc(t, "both_out.3.weight", 0, 2); t += 2 * 8192;
            P.put_tc(t, "both_out.5.weight", 0, 1);
        }
        L.la_b = P.vec("light_attention.feature_convolution.bias");
        L.m1_b = P.vec("both_out.1.bias");
        L.m2_b = P.vec("both_out.3.bias");
        L.m3_b = P.vec("both_out.5.bias");
    }
    if (cfg.has_head) {
        // the same matrices as nn.Linear stores them, for the generic head kernel (any cfg)
        int n = 0;
        auto add_layer = [&](const std::string& w, const std::string& b) {
            int r, c;
            P.find(w, &r, &c);
            L.gen_in[n] = c; L.gen_out[n] = r;
            L.gen_w[n] = P.vec(w); L.gen_b[n] = P.vec(b);
            ++n;
        };
        if (cfg.lightattn) add_layer("light_attention.feature_convolution.weight_center", "light_attention.feature_convolution.bias");
        for (int i = 0; i <= cfg.n_hidden; ++i) {
            const std::string q = "both_out." + std::to_string(2 * i + 1);
            add_layer(q + ".weight", q + ".bias");
        }
        L.gen_layers = n;
        L.ddg = P.alloc(2);
        P.dev[L.ddg] = *P.find("ddg_out.weight");
        P.dev[L.ddg + 1] = *P.find("ddg_out.bias");
    }
    L.zeros = P.alloc(HID);
    L.total = (int64_t)P.dev.size();
    return L;
}

// ---------------- profiling helpers ------------------------------------------
int prof_entry(tmpnn_ctx* c, const char* name) {
    for (size_t i = 0; i < c->prof.size(); ++i)
        if (c->prof[i].name == name) return (int)i;
    c->prof.push_back({name, 0.0, 0});
    return (int)c->prof.size() - 1;
}

cudaEvent_t get_event(tmpnn_ctx* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct LaunchScope {   // brackets one kernel launch with events when profiling
    tmpnn_ctx* c;
    cudaStream_t st;
    PendingEvent pe{};
    bool on;
    LaunchScope(tmpnn_ctx* c_, cudaStream_t st_, const char* name) : c(c_), st(st_), on(c_->profiling) {
        c->launches++;
        if (on) {
            pe.entry = prof_entry(c, name);
            pe.a = get_event(c);
            pe.b = get_event(c);
            cudaEventRecord(pe.a, st);
        }
    }
    ~LaunchScope() {
        if (on) {
            cudaEventRecord(pe.b, st);
            c->pending.push_back(pe);
        }
    }
};

void drain_profile(tmpnn_ctx* c) {
    for (auto& pe : c->pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(pe.b) == cudaSuccess && cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
            c->prof[pe.entry].total_ms += ms;
            c->prof[pe.entry].launches += 1;
        }
        c->event_pool.push_back(pe.a);
        c->event_pool.push_back(pe.b);
    }
    c->pending.clear();
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }


const std::vector<RawTensor>& raw_table_default() {
    static const std::vector<RawTensor> t = build_raw_table(default_model_cfg());
    return t;
}
// The table of the cfg last asked about (name pointers handed out by tmpnn_cfg_weight_info stay valid until the next
// call with a different cfg on this thread).
const std::vector<RawTensor>& raw_table_for(const tmpnn_model_cfg& cfg) {
    if (cfg_is_default(cfg)) return raw_table_default();
    static thread_local tmpnn_model_cfg cached{};
    static thread_local std::vector<RawTensor> t;
    if (t.empty() || std::memcmp(&cached, &cfg, sizeof(cfg)) != 0) {
        t = build_raw_table(cfg);
        cached = cfg;
    }
    return t;
}
int64_t table_floats(const std::vector<RawTensor>& t) { return t.back().offset + (int64_t)t.back().rows * t.back().cols; }

void drop_graph(tmpnn_ctx* c) {
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    c->graph_exec = nullptr;
    c->graph_key = GraphKey{};
}

// A letter outside 0..20 reached the device (TMPNN_IN_DEVICE callers): reported once, then the flag is cleared.
int take_range_error(tmpnn_ctx* c) {
    if (c->h_err && *c->h_err) {
        *c->h_err = 0;
        c->clear_err = true;
        return fail(c, TMPNN_E_RANGE, "tmpnn_run_batch: a device-side S value outside 0..20 was clamped to 20 ('X')");
    }
    return TMPNN_OK;
}

// Device addresses of one call's inputs / outputs.
struct RunIO {
    const int* off; const float* X; const int* S; const int* r; const int* c; const uint8_t* m;
    float* table; float* logp;          // device destinations (either may be NULL = not wanted)
    float* host_table; float* host_logp;   // host destinations of the D2H copies, or NULL
    size_t stage_bytes;
    int B, N, lmax;
    bool uncond;                        // TMPNN_UNCONDITIONAL: ProteinMPNN.unconditional_probs
};

// Enqueue the kernel chain of a call and the D2H of its results on `st` (the staged inputs are already on their way).
// Used both directly and under stream capture (so nothing here may synchronise, allocate or record a host-visible event).
int enqueue_pipeline(tmpnn_ctx* ctx, const RunIO& io, cudaStream_t st) {
    const int N = io.N, B = io.B;
    const size_t n = (size_t)N;
    const float* W = ctx->d_weights;
    const DevLayout& L = ctx->lay;
    const int stop = ctx->debug_stop;
    bool done = false;
    auto stage_done = [&](int s) { if (stop == s) done = true; return done; };
    const int* d_S = ctx->S_dev.p;

    // (h_V starts at zero, protein_mpnn_utils.py:1225; cleared up front so that no memset sits between two kernels
    // of the programmatic-dependent-launch chain)
    CU(cudaMemsetAsync(ctx->hV.p, 0, n * HID * sizeof(float), st));
    { LaunchScope ls(ctx, st, "knn");
      launch_knn(io.off, B, io.X, io.m, io.S, N, io.lmax, ctx->atoms.p, ctx->maskf.p, ctx->S_dev.p, ctx->d_err, ctx->E_idx.p,
                 ctx->D_nb.p, ctx->cnt_enc.p, ctx->cnt_dec.p, st); }
    if (!stage_done(1)) {
        {
            LaunchScope ls(ctx, st, "tc_edge_feat");
            launch_tc_edge_feat(ctx->atoms.p, ctx->E_idx.p, ctx->D_nb.p, io.r, io.c, W + L.pos_table, W + L.rbf_mu, W + L.tc_feat,
                                W + L.ne.g, W + L.ne.b, W + L.be, ctx->hE.p, N, ctx->sm_count, st);
        }
    }
    if (!done && !stage_done(2)) {
        for (int l = 0; l < 3 && !done; ++l) {
            const EncW& e = L.enc[l];
            {   // first layer: h_V = 0, so P_i = b1 and Q_j = 0 (one shared row each, stride 0)
                const float* Pp = l == 0 ? W + e.b1 : ctx->P.p;
                const float* Qp = l == 0 ? W + L.zeros : ctx->Q.p;
                const int stride = l == 0 ? 0 : HID;
                {
                    LaunchScope ls(ctx, st, "tc_enc_msg");
                    launch_tc_msg(false, ctx->hE.p, ctx->E_idx.p, Pp, stride, Qp, stride, ctx->maskf.p, W + e.tc_msg, W + e.b2,
                                  ctx->Spart.p, N, ctx->sm_count, st, l == 1 ? ctx->d_trace : nullptr);
                }
            }
            NodeUpdArgs a{};
            a.weights = W; a.w = e.nu; a.hV = ctx->hV.p; a.hV_copy = nullptr;
            a.Spart = ctx->Spart.p; a.cnt_pre = ctx->cnt_enc.p; a.tc_img = e.nu.tc_stream;
            a.maskf = ctx->maskf.p; a.S = d_S; a.N = N;
            a.trace = (l == 1 && ctx->d_trace) ? ctx->d_trace + 1024 : nullptr;
            a.proj[0] = {e.b11, -1, ctx->P2.p};     // edge update of this layer
            a.proj[1] = {-1, -1, ctx->Q2.p};
            if (l < 2) {
                const EncW& nx = L.enc[l + 1];
                a.proj[2] = {nx.b1, -1, ctx->P.p};
                a.proj[3] = {-1, -1, ctx->Q.p};
            } else {
                const DecW& nx = L.dec[0];
                a.proj[2] = {nx.b1, -1, ctx->P.p};
                a.proj[3] = {-1, io.uncond ? (int64_t)-1 : nx.Tseq, ctx->Q.p};   // (no sequence term for unconditional_probs)
            }
            a.nproj = 4;
            { LaunchScope ls(ctx, st, node_update_uses_tcgen05(N) ? "tc_node_update" : "node_update"); launch_node_update(a, ctx->sm_count, st); }
            {
                LaunchScope ls(ctx, st, "tc_edge_upd");
                launch_tc_edge_upd(ctx->hE.p, ctx->E_idx.p, ctx->P2.p, ctx->Q2.p, ctx->maskf.p, W + e.tc_edge, W + e.b12, W + e.b13,
                                   W + e.n3.g, W + e.n3.b, N, ctx->sm_count, st, (l == 1 && ctx->d_trace) ? ctx->d_trace + 512 : nullptr);
            }
            stage_done(3 + l);
        }
    }
    // unconditional_probs (protein_mpnn_utils.py:1557-1587): every decoder layer gathers the neighbours' ENCODER state, so the
    // layers' neighbour projections are computed once, here (P2 / Q2 are free after the encoder): layer 0's is already in Q
    const float* Qdec[3] = {ctx->Q.p, ctx->Q.p, ctx->Q.p};
    if (!done && io.uncond) {
        const int64_t wk[2] = {L.dec[1].qw_kn, L.dec[2].qw_kn};
        float* const outs[2] = {ctx->P2.p, ctx->Q2.p};
        LaunchScope ls(ctx, st, "proj");
        launch_proj(W, ctx->hV.p, 2, wk, outs, N, st);
        Qdec[1] = ctx->P2.p; Qdec[2] = ctx->Q2.p;
    }
    if (!done) {
        for (int l = 0; l < 3 && !done; ++l) {
            const DecW& d = L.dec[l];
            {
                LaunchScope ls(ctx, st, "tc_dec_msg");
                launch_tc_msg(true, ctx->hE.p, ctx->E_idx.p, ctx->P.p, HID, Qdec[l], HID, ctx->maskf.p, W + d.tc_msg, W + d.b2,
                              ctx->Spart.p, N, ctx->sm_count, st);
            }
            NodeUpdArgs a{};
            a.weights = W; a.w = d.nu; a.hV = ctx->hV.p;
            a.Spart = ctx->Spart.p; a.cnt_pre = ctx->cnt_dec.p; a.tc_img = d.nu.tc_stream;
            a.hV_copy = l == 0 ? ctx->dec_first.p : (l == 1 ? ctx->dec_prev.p : ctx->dec_last.p);
            a.maskf = ctx->maskf.p; a.S = d_S; a.N = N;
            a.nproj = 0;
            if (l < 2) {
                const DecW& nx = L.dec[l + 1];
                a.proj[0] = {nx.b1, -1, ctx->P.p};
                a.proj[1] = {-1, nx.Tseq, ctx->Q.p};
                a.nproj = io.uncond ? 1 : 2;            // unconditional: the neighbour projections are fixed (above)
            }
            { LaunchScope ls(ctx, st, node_update_uses_tcgen05(N) ? "tc_node_update" : "node_update"); launch_node_update(a, ctx->sm_count, st); }
            stage_done(6 + l);
        }
    }
    if (!done) {
        if (io.table) {
            if (cfg_is_default(ctx->cfg)) {
                HeadArgs h{};
                h.weights = W; h.Ws = L.Ws; h.la_b = L.la_b; h.m1_b = L.m1_b; h.m2_b = L.m2_b;
                h.m3_b = L.m3_b; h.ddg = L.ddg;
                h.dec_last = ctx->dec_last.p; h.dec_prev = ctx->dec_prev.p; h.S = d_S; h.table = io.table; h.N = N;
                const bool tc = head_uses_tcgen05(N);
                LaunchScope ls(ctx, st, tc ? "tc_head" : "head");
                if (tc) launch_tc_head(h, L.tc_head, ctx->sm_count, st);
                else launch_head_mma(h, L.head_mma, st);
            } else {
                HeadGenArgs h{};
                h.weights = W; h.nlayers = L.gen_layers; h.relu_first = ctx->cfg.lightattn ? 0 : 1;
                for (int i = 0; i < L.gen_layers; ++i) {
                    h.in_dim[i] = L.gen_in[i]; h.out_dim[i] = L.gen_out[i]; h.w[i] = L.gen_w[i]; h.b[i] = L.gen_b[i];
                }
                h.Ws = L.Ws; h.ddg = L.ddg;
                h.dec[0] = ctx->dec_last.p; h.dec[1] = ctx->dec_prev.p; h.dec[2] = ctx->dec_first.p;
                h.num_final = ctx->cfg.num_final_layers; h.S = d_S; h.table = io.table; h.N = N;
                LaunchScope ls(ctx, st, "head_generic");
                launch_head_generic(h, st);
            }
        }
        if (io.logp) {
            LogProbArgs a{};
            a.weights = W; a.Wout = L.Wout; a.bout = L.bout; a.hV = ctx->dec_last.p; a.out = io.logp; a.N = N;
            LaunchScope ls(ctx, st, "log_probs");
            launch_log_probs(a, st);
        }
        if (io.host_table) CU(cudaMemcpyAsync(io.host_table, io.table, n * 21 * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (io.host_logp) CU(cudaMemcpyAsync(io.host_logp, io.logp, n * 21 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaGetLastError());
    CU(tc_take_error());
    return TMPNN_OK;
}

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

int tmpnn_abi_version(void) { return TMPNN_ABI_VERSION; }

void tmpnn_default_cfg(tmpnn_model_cfg* cfg) { if (cfg) *cfg = default_model_cfg(); }

int tmpnn_cfg_weight_count(const tmpnn_model_cfg* cfg) {
    if (!cfg || !cfg_valid(*cfg)) return TMPNN_E_INVALID;
    return (int)raw_table_for(*cfg).size();
}

int tmpnn_cfg_weight_info(const tmpnn_model_cfg* cfg, int index, const char** name, int64_t* offset, int32_t* rows, int32_t* cols) {
    if (!cfg || !cfg_valid(*cfg)) return TMPNN_E_INVALID;
    const auto& t = raw_table_for(*cfg);
    if (index < 0 || index >= (int)t.size()) return TMPNN_E_INVALID;
    if (name) *name = t[index].name.c_str();
    if (offset) *offset = t[index].offset;
    if (rows) *rows = t[index].rows;
    if (cols) *cols = t[index].cols;
    return TMPNN_OK;
}

int64_t tmpnn_cfg_weight_blob_floats(const tmpnn_model_cfg* cfg) {
    if (!cfg || !cfg_valid(*cfg)) return TMPNN_E_INVALID;
    return table_floats(raw_table_for(*cfg));
}

int tmpnn_weight_count(void) { return (int)raw_table_default().size(); }

int tmpnn_weight_info(int index, const char** name, int64_t* offset, int32_t* rows, int32_t* cols) {
    const tmpnn_model_cfg c = default_model_cfg();
    return tmpnn_cfg_weight_info(&c, index, name, offset, rows, cols);
}

int64_t tmpnn_weight_blob_floats(void) { return table_floats(raw_table_default()); }

int tmpnn_create(int device, tmpnn_ctx** out) {
    if (!out) {
        g_create_error = "tmpnn_create: out is NULL";
        return TMPNN_E_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("tmpnn_create: no CUDA device (") + cudaGetErrorString(e) +
                         "); thermompnn_b200 has no CPU path";
        return TMPNN_E_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "tmpnn_create: device index out of range";
        return TMPNN_E_INVALID;
    }
    DeviceGuard guard(device);
    tmpnn_ctx* ctx = new tmpnn_ctx();
    ctx->device = device;
    ctx->raw = raw_table_default();
    auto bail = [&](const std::string& m) {
        g_create_error = m;
        tmpnn_destroy(ctx);
        return TMPNN_E_CUDA;
    };
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return bail(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return bail("tmpnn_create: device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                    ", this library is built for sm_100a (B200) only");
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&ctx->stage_ev, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->run_ev, cudaEventDisableTiming)) != cudaSuccess)
        return bail(std::string("cudaEventCreate: ") + cudaGetErrorString(e));
    if ((e = cudaMalloc(&ctx->d_err, sizeof(int))) != cudaSuccess || (e = cudaMemset(ctx->d_err, 0, sizeof(int))) != cudaSuccess ||
        (e = cudaMallocHost(&ctx->h_err, sizeof(int))) != cudaSuccess)
        return bail(std::string("error flag allocation: ") + cudaGetErrorString(e));
    *ctx->h_err = 0;
    if ((e = kernels_init()) != cudaSuccess) return bail(std::string("kernels_init: ") + cudaGetErrorString(e));
    if ((e = node_kernels_init()) != cudaSuccess) return bail(std::string("node_kernels_init: ") + cudaGetErrorString(e));
    if ((e = tc_kernels_init()) != cudaSuccess) return bail(std::string("tc_kernels_init: ") + cudaGetErrorString(e));
    if ((e = head_kernels_init()) != cudaSuccess) return bail(std::string("head_kernels_init: ") + cudaGetErrorString(e));
    if ((e = tc_node_kernels_init()) != cudaSuccess) return bail(std::string("tc_node_kernels_init: ") + cudaGetErrorString(e));
    if ((e = tc_head_kernels_init()) != cudaSuccess) return bail(std::string("tc_head_kernels_init: ") + cudaGetErrorString(e));
#ifdef TMPNN_DEV
    if (const char* env = std::getenv("TMPNN_TC_TRACE"))
        if (std::atoi(env)) { cudaMalloc(&ctx->d_trace, 3 * 512 * sizeof(long long)); cudaMemset(ctx->d_trace, 0, 3 * 512 * sizeof(long long)); }
#endif
    if (const char* env = std::getenv("TMPNN_GRAPHS")) ctx->graphs = std::atoi(env) != 0;
    *out = ctx;
    return TMPNN_OK;
}

void tmpnn_destroy(tmpnn_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->run_pending && ctx->run_ev) cudaEventSynchronize(ctx->run_ev);
    drain_profile(ctx);
    drop_graph(ctx);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->stage_ev) cudaEventDestroy(ctx->stage_ev);
    if (ctx->run_ev) cudaEventDestroy(ctx->run_ev);
    if (ctx->d_weights) cudaFree(ctx->d_weights);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_err) cudaFree(ctx->d_err);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->d_trace) cudaFree(ctx->d_trace);
    ctx->d_in.release();
    ctx->atoms.release(); ctx->maskf.release(); ctx->hV.release(); ctx->dec_first.release(); ctx->dec_prev.release();
    ctx->dec_last.release(); ctx->P.release(); ctx->Q.release(); ctx->P2.release(); ctx->Q2.release();
    ctx->Spart.release(); ctx->S_dev.release(); ctx->logp.release(); ctx->cnt_enc.release(); ctx->cnt_dec.release();
    ctx->table.release(); ctx->E_idx.release(); ctx->D_nb.release(); ctx->hE.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* tmpnn_last_error(const tmpnn_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tmpnn_load_weights_cfg(tmpnn_ctx* ctx, const tmpnn_model_cfg* cfg, const float* host_blob, int64_t n_floats) {
    if (!ctx) return TMPNN_E_INVALID;
    if (!cfg || !cfg_valid(*cfg)) return fail(ctx, TMPNN_E_INVALID, "tmpnn_load_weights: invalid model configuration");
    if (!host_blob) return fail(ctx, TMPNN_E_INVALID, "tmpnn_load_weights: blob is NULL");
    const std::vector<RawTensor> table = build_raw_table(*cfg);
    if (n_floats != table_floats(table))
        return fail(ctx, TMPNN_E_RANGE, "tmpnn_load_weights: blob has " + std::to_string(n_floats) + " floats, expected " +
                                            std::to_string(table_floats(table)));
    for (int64_t i = 0; i < n_floats; ++i)
        if (!std::isfinite(host_blob[i])) return fail(ctx, TMPNN_E_INVALID, "tmpnn_load_weights: non-finite weight");
    Prep P{host_blob, table, {}};
    const DevLayout lay = prepare_weights(P, *cfg);
    DeviceGuard guard(ctx->device);
    if (ctx->run_pending) { CU(cudaEventSynchronize(ctx->run_ev)); ctx->run_pending = false; }
    drop_graph(ctx);
    if (ctx->d_weights) cudaFree(ctx->d_weights);
    ctx->d_weights = nullptr;
    ctx->have_weights = false;
    CU(cudaMalloc(&ctx->d_weights, P.dev.size() * sizeof(float)));
    ++g_realloc_count;
    CU(cudaMemcpyAsync(ctx->d_weights, P.dev.data(), P.dev.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->cfg = *cfg;
    ctx->raw = table;
    ctx->lay = lay;
    ctx->have_weights = true;
    return TMPNN_OK;
}

int tmpnn_load_weights(tmpnn_ctx* ctx, const float* host_blob, int64_t n_floats) {
    const tmpnn_model_cfg c = default_model_cfg();
    return tmpnn_load_weights_cfg(ctx, &c, host_blob, n_floats);
}

int tmpnn_run_batch_ex(tmpnn_ctx* ctx, int32_t B, const int32_t* offsets, const float* X, const int32_t* S,
                       const int32_t* residue_idx, const int32_t* chain_enc, const uint8_t* mask, float* out_table,
                       float* out_log_probs, uint32_t flags, void* stream_) {
    if (!ctx) return TMPNN_E_INVALID;
    if (!ctx->have_weights) return fail(ctx, TMPNN_E_NOWEIGHTS, "tmpnn_run_batch: weights not loaded");
    if (B <= 0 || !offsets || !X || !S || !residue_idx || !chain_enc || !mask)
        return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: NULL pointer or B <= 0");
    if (!out_table && !out_log_probs) return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: no output requested");
    if ((flags & TMPNN_UNCONDITIONAL) && (out_table || !out_log_probs))
        return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: TMPNN_UNCONDITIONAL produces log-probabilities only (out_table must be NULL)");
    if (out_table && !ctx->cfg.has_head)
        return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: the loaded model has no ddG head (has_head = 0); ask for log_probs only");
    if (offsets[0] != 0) return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: offsets[0] must be 0");
    int lmax = 0;
    for (int b = 0; b < B; ++b) {
        int len = offsets[b + 1] - offsets[b];
        if (len <= 0) return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: empty protein or unsorted offsets at " + std::to_string(b));
        lmax = std::max(lmax, len);
    }
    if (lmax > knn_max_len())
        return fail(ctx, TMPNN_E_RANGE, "tmpnn_run_batch: protein longer than " + std::to_string(knn_max_len()) + " residues");
    const int N = offsets[B];
    const bool in_dev = flags & TMPNN_IN_DEVICE, out_dev = flags & TMPNN_OUT_DEVICE;
    if (!in_dev) {
        for (int i = 0; i < N; ++i)
            if (S[i] < 0 || S[i] > 20) return fail(ctx, TMPNN_E_RANGE, "tmpnn_run_batch: S[" + std::to_string(i) + "] outside 0..20");
    }
    DeviceGuard guard(ctx->device);
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    // a previous TMPNN_NO_SYNC call may have left a range error behind
    if (ctx->run_pending && cudaEventQuery(ctx->run_ev) == cudaSuccess) {
        ctx->run_pending = false;
        if (int rc = take_range_error(ctx)) return rc;
    }

    // ---- inputs: one packed staging buffer -> one H2D -------------------------------
    const size_t o_off = 0;
    const size_t o_X = align16(o_off + (size_t)(B + 1) * 4);
    const size_t o_S = align16(o_X + (size_t)N * 48);
    const size_t o_r = align16(o_S + (size_t)N * 4);
    const size_t o_c = align16(o_r + (size_t)N * 4);
    const size_t o_m = align16(o_c + (size_t)N * 4);
    const size_t in_bytes = align16(o_m + (size_t)N);
    const size_t used = in_dev ? align16((size_t)(B + 1) * 4) : in_bytes;
    if (ctx->stage_busy) {   // a previous TMPNN_NO_SYNC call may still be reading the staging buffer
        CU(cudaEventSynchronize(ctx->stage_ev));
        ctx->stage_busy = false;
    }
    const uint64_t realloc0 = g_realloc_count;
    CU(ctx->d_in.reserve(used));
    if (ctx->h_stage_cap < used) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_cap = 0;
        CU(cudaMallocHost(&ctx->h_stage, used + used / 4));
        ctx->h_stage_cap = used + used / 4;
        ++g_realloc_count;
    }
    std::memcpy(ctx->h_stage + o_off, offsets, (size_t)(B + 1) * 4);
    RunIO io{};
    io.B = B; io.N = N; io.lmax = lmax; io.stage_bytes = used; io.uncond = (flags & TMPNN_UNCONDITIONAL) != 0;
    io.off = reinterpret_cast<const int*>(ctx->d_in.p + o_off);
    io.X = X; io.S = S; io.r = residue_idx; io.c = chain_enc; io.m = mask;
    if (!in_dev) {
        std::memcpy(ctx->h_stage + o_X, X, (size_t)N * 48);
        std::memcpy(ctx->h_stage + o_S, S, (size_t)N * 4);
        std::memcpy(ctx->h_stage + o_r, residue_idx, (size_t)N * 4);
        std::memcpy(ctx->h_stage + o_c, chain_enc, (size_t)N * 4);
        std::memcpy(ctx->h_stage + o_m, mask, (size_t)N);
        io.X = reinterpret_cast<const float*>(ctx->d_in.p + o_X);
        io.S = reinterpret_cast<const int*>(ctx->d_in.p + o_S);
        io.r = reinterpret_cast<const int*>(ctx->d_in.p + o_r);
        io.c = reinterpret_cast<const int*>(ctx->d_in.p + o_c);
        io.m = ctx->d_in.p + o_m;
    }

    // ---- buffers --------------------------------------------------------------
    const size_t n = (size_t)N, ne = n * KNB;
    CU(ctx->atoms.reserve(n * 16)); CU(ctx->maskf.reserve(n)); CU(ctx->S_dev.reserve(n));
    CU(ctx->cnt_enc.reserve(n)); CU(ctx->cnt_dec.reserve(n));
    CU(ctx->hV.reserve(n * HID)); CU(ctx->dec_first.reserve(n * HID)); CU(ctx->dec_prev.reserve(n * HID));
    CU(ctx->dec_last.reserve(n * HID)); CU(ctx->Spart.reserve(n * 3 * HID));
    CU(ctx->P.reserve(n * HID)); CU(ctx->Q.reserve(n * HID)); CU(ctx->P2.reserve(n * HID)); CU(ctx->Q2.reserve(n * HID));
    CU(ctx->E_idx.reserve(ne)); CU(ctx->D_nb.reserve(ne)); CU(ctx->hE.reserve(ne * HID));
    io.table = out_table; io.logp = out_log_probs;
    if (!out_dev) {
        if (out_table) { CU(ctx->table.reserve(n * 21)); io.table = ctx->table.p; io.host_table = out_table; }
        if (out_log_probs) { CU(ctx->logp.reserve(n * 21)); io.logp = ctx->logp.p; io.host_logp = out_log_probs; }
    }
    if (g_realloc_count != realloc0) ctx->buf_epoch = g_realloc_count;
    ctx->last_N = N;
    ctx->last_table = io.table;
    ctx->last_logp = io.logp;

    // runs on one context are ordered even across streams (they share the buffers above)
    if (ctx->run_pending) CU(cudaStreamWaitEvent(st, ctx->run_ev, 0));
    if (ctx->clear_err) { CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), st)); ctx->clear_err = false; }
    CU(cudaMemcpyAsync(ctx->d_in.p, ctx->h_stage, io.stage_bytes, cudaMemcpyHostToDevice, st));   // ONE copy for all inputs
    CU(cudaEventRecord(ctx->stage_ev, st));
    ctx->stage_busy = true;

    // ---- the kernel chain: launched directly, or replayed from the graph captured for this call shape ----
    GraphKey key;
    key.B = B; key.N = N; key.lmax = lmax; key.stop = ctx->debug_stop; key.flags = flags; key.epoch = ctx->buf_epoch;
    key.ptr[0] = in_dev ? X : nullptr; key.ptr[1] = in_dev ? S : nullptr; key.ptr[2] = in_dev ? residue_idx : nullptr;
    key.ptr[3] = in_dev ? chain_enc : nullptr; key.ptr[4] = in_dev ? mask : nullptr;
    key.ptr[5] = out_table; key.ptr[6] = out_log_probs;
    const bool graph_ok = ctx->graphs && !ctx->profiling && !ctx->d_trace;
    bool launched = false;
    if (graph_ok && ctx->graph_exec && ctx->graph_key == key) {
        CU(cudaGraphLaunch(ctx->graph_exec, st));
        ctx->launches += ctx->graph_kernels;
        ctx->graph_launches++;
        launched = true;
    } else if (graph_ok && ctx->last_key == key) {
        // the same call shape twice in a row: capture the chain (kernels launched under capture do not run)
        drop_graph(ctx);
        const int64_t l0 = ctx->launches;
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (ce == cudaSuccess) {
            const int rc = enqueue_pipeline(ctx, io, st);
            ce = cudaStreamEndCapture(st, &g);
            if (rc == TMPNN_OK && ce == cudaSuccess && g) {
                ce = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
                if (ce == cudaSuccess) {
                    ctx->graph_key = key;
                    ctx->graph_kernels = ctx->launches - l0;
                    ctx->launches = l0;
                }
            }
            if (g) cudaGraphDestroy(g);
        }
        if (ctx->graph_exec) {
            CU(cudaGraphLaunch(ctx->graph_exec, st));
            ctx->launches += ctx->graph_kernels;
            ctx->graph_launches++;
            launched = true;
        } else {
            (void)cudaGetLastError();       // capture is an optimisation: fall back to direct launches
            ctx->launches = l0;
            ctx->graphs = false;
        }
    }
    ctx->last_key = key;
    if (!launched) {
        if (int rc = enqueue_pipeline(ctx, io, st)) return rc;
    }
    CU(cudaEventRecord(ctx->run_ev, st));
    ctx->run_pending = true;
    if (!(flags & TMPNN_NO_SYNC) || !out_dev) {
        CU(cudaStreamSynchronize(st));
        ctx->stage_busy = false;
        ctx->run_pending = false;
        if (ctx->profiling) drain_profile(ctx);
        if (int rc = take_range_error(ctx)) return rc;
    }
    return TMPNN_OK;
}

int tmpnn_run_batch(tmpnn_ctx* ctx, int32_t B, const int32_t* offsets, const float* X, const int32_t* S,
                    const int32_t* residue_idx, const int32_t* chain_enc, const uint8_t* mask, float* out_table,
                    uint32_t flags, void* stream_) {
    if (ctx && !out_table) return fail(ctx, TMPNN_E_INVALID, "tmpnn_run_batch: NULL pointer or B <= 0");
    return tmpnn_run_batch_ex(ctx, B, offsets, X, S, residue_idx, chain_enc, mask, out_table, nullptr, flags, stream_);
}

int tmpnn_set_graphs(tmpnn_ctx* ctx, int enabled) {
    if (!ctx) return TMPNN_E_INVALID;
    ctx->graphs = enabled != 0;
    if (!ctx->graphs) { DeviceGuard guard(ctx->device); drop_graph(ctx); }
    return TMPNN_OK;
}

int tmpnn_fetch(tmpnn_ctx* ctx, const char* name, void* host_dst, int64_t bytes) {
    if (!ctx || !name || !host_dst) return TMPNN_E_INVALID;
    const size_t n = (size_t)ctx->last_N;
    if (n == 0) return fail(ctx, TMPNN_E_INVALID, "tmpnn_fetch: no batch has been run");
    const void* src = nullptr;
    size_t sz = 0;
    const std::string s = name;
    if (s == "E_idx") { src = ctx->E_idx.p; sz = n * KNB * 4; }
    else if (s == "D_nb") { src = ctx->D_nb.p; sz = n * KNB * 4; }
    else if (s == "h_E") { src = ctx->hE.p; sz = n * KNB * HID * 4; }
    else if (s == "h_V") { src = ctx->hV.p; sz = n * HID * 4; }
    else if (s == "dec0") { src = ctx->dec_first.p; sz = n * HID * 4; }
    else if (s == "dec1") { src = ctx->dec_prev.p; sz = n * HID * 4; }
    else if (s == "dec2") { src = ctx->dec_last.p; sz = n * HID * 4; }
    else if (s == "table" && ctx->last_table) { src = ctx->last_table; sz = n * 21 * 4; }
    else if (s == "log_probs" && ctx->last_logp) { src = ctx->last_logp; sz = n * 21 * 4; }
    else if (s == "tc_trace" && ctx->d_trace) { src = ctx->d_trace; sz = 2 * 512 * sizeof(long long); }
    else if (s == "node_trace" && ctx->d_trace) { src = ctx->d_trace + 1024; sz = 64 * sizeof(long long); }
    else return fail(ctx, TMPNN_E_NOTFOUND, "tmpnn_fetch: unknown tensor '" + s + "'");
    if ((size_t)bytes != sz)
        return fail(ctx, TMPNN_E_RANGE, "tmpnn_fetch: '" + s + "' is " + std::to_string(sz) + " bytes, caller gave " + std::to_string(bytes));
    DeviceGuard guard(ctx->device);
    if (ctx->run_pending) { CU(cudaEventSynchronize(ctx->run_ev)); ctx->run_pending = false; }   // whichever stream the last run used
    CU(cudaMemcpy(host_dst, src, sz, cudaMemcpyDeviceToHost));
    return TMPNN_OK;
}

int tmpnn_set_debug_stop(tmpnn_ctx* ctx, int stage) {
    if (!ctx || stage < 0 || stage > 8) return TMPNN_E_INVALID;
    ctx->debug_stop = stage;
    return TMPNN_OK;
}

int tmpnn_set_profiling(tmpnn_ctx* ctx, int enabled) {
    if (!ctx) return TMPNN_E_INVALID;
    ctx->profiling = enabled != 0;
    return TMPNN_OK;
}

int tmpnn_profile_count(tmpnn_ctx* ctx) {
    if (!ctx) return TMPNN_E_INVALID;
    DeviceGuard guard(ctx->device);
    drain_profile(ctx);   // blocks until the recorded events have completed
    return (int)ctx->prof.size();
}

int tmpnn_profile_get(tmpnn_ctx* ctx, int index, const char** kernel_name, double* total_ms, int64_t* launches) {
    if (!ctx || index < 0 || index >= (int)ctx->prof.size()) return TMPNN_E_INVALID;
    if (kernel_name) *kernel_name = ctx->prof[index].name.c_str();
    if (total_ms) *total_ms = ctx->prof[index].total_ms;
    if (launches) *launches = ctx->prof[index].launches;
    return TMPNN_OK;
}

int tmpnn_profile_reset(tmpnn_ctx* ctx) {
    if (!ctx) return TMPNN_E_INVALID;
    DeviceGuard guard(ctx->device);
    drain_profile(ctx);
    ctx->prof.clear();
    ctx->launches = 0;
    ctx->graph_launches = 0;
    return TMPNN_OK;
}

int64_t tmpnn_launch_count(const tmpnn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t tmpnn_graph_launch_count(const tmpnn_ctx* ctx) { return ctx ? ctx->graph_launches : 0; }

}  // extern "C"
