// C ABI of libgenvc_b200.so (include/genvc_b200.h): context, weight-blob layout, and the host-side
// orchestration of the path — perceiver, prefix embedding, prefill, decode (fused persistent kernel
// or per-op kernels), teacher-forced latent pass.  No torch types; the caller owns device memory.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/genvc_b200.h"
#include "layout.h"
#include "mega.cuh"
#include "ops.cuh"
#include "stream_layout.h"

using namespace gv;

namespace {

constexpr size_t kSplitKFloats = size_t(4) << 20;  // deterministic split-K partials (ops.cu)

struct Workspace {
    size_t bytes = 0;
    size_t take(size_t n, size_t align = 256) {
        bytes = (bytes + align - 1) / align * align;
        size_t o = bytes;
        bytes += n;
        return o;
    }
};

}  // namespace

struct genvc_ctx {
    genvc_config cfg;
    int device = 0;
    int n_sm = 0;
    int grid = 0;  // CTAs of the fused decode kernel
    bool mega_ok = false;
    bool batch_ok = false;  // batched fused decode kernel (decode_batch.cu) supports this shape
    int fused_rows = 1;     // rows the exchange buffers are sized for (1, or GV_BATCH_ROWS when max_batch > 1)
    int gs_cur = 0;         // which copy of the (double-buffered) generation state is current
    Layout layout;
    StreamDims sdims;
    mutable std::string err;
    unsigned long long nlaunch = 0;

    const float* blob = nullptr;
    float* stream = nullptr;
    bool stream_packed = false;
    float* kv = nullptr;
    float* vw = nullptr;  // projected-value cache of the single-row fused kernel (optional: genvc_bind_vw)
    char* ws = nullptr;

    // workspace offsets (bytes)
    size_t o_state, o_seen, state_stride = 0, seen_stride = 0, o_tokx, o_flags, o_status_scratch, o_pend_logits, o_pend_latent, o_splitk, o_X, o_A, o_QKV, o_U;
    // exchange buffers of the fused decode kernel ({value, tag} pairs; one contiguous region)
    size_t o_xchg, xchg_bytes, o_xq, o_matt_o, o_matt_ml, o_x1, o_pp, o_x2, o_lg, o_hops, o_acc, o_ao, o_attcnt, o_sbuf, o_tcptr, o_gbar;
    size_t hops_bytes = 0;
    size_t acc_bytes = 0;
    uint32_t tag_next = 1;
    size_t o_pc_melT, o_pc_ctx, o_pc_kv, o_pc_lat, o_pc_q, o_pc_o, o_pc_h, o_pc_g;
    size_t ws_bytes = 0;
    int Vpad = 0;

    // debug timeline of the fused decode kernel (genvc_debug_trace)
    unsigned long long* trace = nullptr;
    int trace_slots = 0, trace_step = 0;
    int window = 2, dbg_nosync = 0, l2_ahead = 0, hop_settle = 0, hop_hold = 0, hop_near = -1, hop_near_ao = -1;

    // prefill as a CUDA graph: the ~270 launches of the 30 blocks + head are captured once per (batch, prefix length) and
    // replayed with one launch (the per-op prefill was bound by the host's launch rate: 2.4 ms on one box, 3.5 ms on another)
    struct PrefillGraph {
        cudaGraphExec_t exec = nullptr;
        unsigned long long launches = 0;
    };
    std::map<std::pair<int, int>, PrefillGraph> prefill_graphs;
    std::map<std::pair<int, int>, PrefillGraph> perceiver_graphs;  // the perceiver's ~45 launches, per (batch, mel frames)
    void drop_graphs() {  // captured launches bake workspace / weight pointers
        for (auto* m : {&prefill_graphs, &perceiver_graphs}) {
            for (auto& kv : *m)
                if (kv.second.exec) (void)cudaGraphExecDestroy(kv.second.exec);
            m->clear();
        }
    }
    bool use_graphs = true;
    bool pc_attention_tc = true;  // perceiver cross-attention on tcgen05 (GENVC_PC_TC=0: CUDA-core attention kernel)
    // persistent fused prefill (gemm_tc.cu): device table of the packed tensor-core weights of the blocks
    std::vector<const float*> tc_table;  // [L][4] filled by genvc_pack_tc
    bool tc_table_uploaded = false;
    // Off by default: measured 2.8 ms against 2.4 ms for the per-op prefill replayed as a CUDA graph (48 rows).  Its GEMM
    // phases run at ~1 us per 32 KB stage even with MMAs, activation loads and weight copies disabled (mbarrier hand-offs
    // between loader warps and the MMA thread) and every grid barrier costs 2.4-3.6 us (two gpu-scope fences under the
    // weight stream): profiles/r02b/prefill_fused_phases.txt.  GENVC_FUSED_PREFILL=1 enables it (parity-green).
    bool use_fused_prefill = false;

    // host mirror of the generation state
    int B = 0, P = 0;
    bool prefilled = false, pending = false;
    bool vw_filled = false;  // the projected-value cache covers the cached positions (false after a prefill until the first fused forward)
    // The projected-value variant pays once per sequence (its first forward projects the cached V rows: ~0.5 ms at a 48-row
    // prefix) and saves ~11 us per token: it is used when at least this many tokens may still be generated.
    int vw_min_tokens = 96;
    bool seg_pvw = false;    // decided at the first fused launch after a prefill, kept for the sequence
    int n_host = 0;

    int D() const { return cfg.d_model; }
    int H() const { return cfg.n_head; }
    int hd() const { return cfg.d_model / cfg.n_head; }
    int V() const { return cfg.n_audio_vocab; }
    size_t rows_cap() const { return (size_t)cfg.max_batch * cfg.max_seq; }
    size_t kv_plane() const { return (size_t)cfg.max_batch * cfg.d_model * cfg.max_seq; }  // floats per (layer, k|v)
    size_t kv_batch_stride() const { return (size_t)cfg.d_model * cfg.max_seq; }

    template <class T>
    T* at(size_t off) const {
        return reinterpret_cast<T*>(ws + off);
    }
    GenState* gstate(int which) const { return at<GenState>(o_state + (size_t)which * state_stride); }
    unsigned char* gseen(int which) const { return at<unsigned char>(o_seen + (size_t)which * seen_stride); }
    const float* w(uint64_t off) const { return blob + off; }

    int fail(int code, const char* fmt, ...) const {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

// current-device guard of the entry points: the library works on ctx->device and restores the caller's device
struct DevGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit DevGuard(int d) : dev(d) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DevGuard() {
        if (prev >= 0 && prev != dev) (void)cudaSetDevice(prev);
    }
};

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return ctx->fail(GENVC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                             __LINE__);                                                                       \
    } while (0)

static bool valid_hd(int hd) { return hd == 32 || hd == 64 || hd == 128 || hd == 256; }

static void plan_workspace(genvc_ctx* c) {
    const genvc_config& g = c->cfg;
    const size_t D = g.d_model, V = g.n_audio_vocab, MB = g.max_batch, F = sizeof(float);
    c->Vpad = (int)((V + 15) / 16 * 16);
    Workspace w;
    // generation state + repetition-penalty sets, two copies each: a fused launch reads one and writes the other
    c->state_stride = (sizeof(GenState) + 64 + 255) / 256 * 256;
    c->o_state = w.take(2 * c->state_stride);
    c->seen_stride = (MB * c->Vpad + 255) / 256 * 256;
    c->o_seen = w.take(2 * c->seen_stride);
    c->o_status_scratch = w.take(64);
    c->o_flags = w.take(64);  // int[0]: an embedding kernel clamped an out-of-range id (decode path); int[1]: same, latent pass
    c->o_pend_logits = w.take(MB * V * F);
    c->o_pend_latent = w.take(MB * D * F);
    // exchange buffers of the fused decode kernel: {value, tag} pairs (2 floats per element)
    // (single-row kernel: H * att_nsplit() <= H * 8 items; batched kernel: rows * H * nsplit <= grid items)
    const size_t FR = c->fused_rows = (MB > 1 && c->batch_ok) ? GV_BATCH_ROWS : 1;
    const size_t items = std::max((size_t)g.n_head * 8, FR > 1 ? (size_t)c->grid : (size_t)0);
    c->o_xchg = c->o_xq = w.take(2 * FR * 3 * D * F);
    c->o_matt_o = w.take(2 * items * (D / g.n_head) * F);
    c->o_matt_ml = w.take(2 * items * 2 * F);
    c->o_x1 = w.take(2 * FR * D * F);
    c->o_pp = w.take(2 * (size_t)std::max(c->grid, 1) * FR * D * F);
    c->o_x2 = w.take(2 * FR * D * F);
    c->o_lg = w.take(2 * FR * (size_t)c->Vpad * F);
    c->o_tokx = w.take(2 * GV_BATCH_ROWS * F);
    c->o_ao = w.take(2 * FR * D * F);
    c->o_sbuf = w.take(2 * (size_t)g.n_head * g.max_seq * F);  // scaled attention scores of a step (single-row kernel, PVW)
    // arrival counters (zeroed before every fused launch): the hop counters, then one item counter per (row, head)
    c->o_hops = w.take(HC_COUNT * GV_HOP_STRIDE * sizeof(unsigned));
    c->o_attcnt = w.take((size_t)std::max(c->grid, 1) * GV_ATTCNT_STRIDE * sizeof(unsigned));
    c->hops_bytes = w.take(0) - c->o_hops;
    c->acc_bytes = 2 * (size_t)g.n_layer * D * sizeof(unsigned long long);
    c->o_acc = w.take(c->acc_bytes);
    c->xchg_bytes = w.take(0) - c->o_xchg;
    c->o_tcptr = w.take((size_t)g.n_layer * 4 * sizeof(const float*));
    c->o_gbar = w.take(256);
    c->o_splitk = w.take(kSplitKFloats * F);
    const size_t R = c->rows_cap();
    c->o_X = w.take(R * D * F);
    c->o_A = w.take(R * D * F);
    c->o_QKV = w.take(R * 3 * D * F);
    c->o_U = w.take(R * 4 * D * F);
    // perceiver
    const size_t S = g.max_mel_frames, NL = g.pc_latents, inner = (size_t)g.pc_dim_head * g.pc_heads;
    const size_t Rc = MB * (NL + S);
    c->o_pc_melT = w.take(MB * S * c->layout.pc_ctx_pad * F);
    c->o_pc_ctx = w.take(Rc * D * F);
    c->o_pc_kv = w.take(Rc * 2 * inner * F);
    c->o_pc_lat = w.take(MB * NL * D * F);
    c->o_pc_q = w.take(MB * NL * inner * F);
    c->o_pc_o = w.take(MB * NL * inner * F);
    c->o_pc_h = w.take(MB * NL * 2 * g.pc_ff_inner * F);
    c->o_pc_g = w.take(MB * NL * c->layout.pc_ff_inner_pad * F);
    c->ws_bytes = (w.bytes + 255) / 256 * 256;
}

extern "C" {

int genvc_create(const genvc_config* cfg, int device, genvc_ctx** out) {
    if (!cfg || !out) return GENVC_E_INVALID;
    *out = nullptr;
    genvc_ctx* ctx = new (std::nothrow) genvc_ctx();
    if (!ctx) return GENVC_E_INVALID;
    ctx->cfg = *cfg;
    ctx->device = device;
    const genvc_config& g = ctx->cfg;
    auto bad = [&](const char* why) {
        // the context is returned so the caller can read the message, then destroy it
        *out = ctx;
        return ctx->fail(GENVC_E_INVALID, "genvc_create: %s", why);
    };
    if (g.n_layer <= 0 || g.d_model <= 0 || g.n_head <= 0 || g.d_model % g.n_head) return bad("bad n_layer/d_model/n_head");
    if (g.d_model % 128 || g.d_model > 1024) return bad("d_model must be a multiple of 128 and <= 1024");
    if (!valid_hd(g.d_model / g.n_head)) return bad("head_dim must be 32, 64, 128 or 256");
    if (g.n_audio_vocab <= 0 || g.n_audio_vocab > 2048) return bad("n_audio_vocab must be in (0, 2048]");
    if (g.max_batch <= 0 || g.max_batch > GV_MAX_BATCH) return bad("max_batch must be in [1, 64]");
    if (g.max_seq <= 0 || g.max_mel_frames <= 0) return bad("max_seq / max_mel_frames must be positive");
    if (!valid_hd(g.pc_dim_head) || g.pc_dim_context % 4 || g.pc_latents <= 0 || g.pc_depth <= 0 || g.pc_ff_inner <= 0)
        return bad("unsupported perceiver shape");
    if (g.start_audio < 0 || g.start_audio >= g.n_audio_vocab || g.stop_audio < 0 || g.stop_audio >= g.n_audio_vocab)
        return bad("start/stop audio token outside the vocabulary");
    ctx->layout.build(g);
    if (const char* e = getenv("GENVC_GRAPH")) ctx->use_graphs = e[0] != '0';
    if (const char* e = getenv("GENVC_PC_TC")) ctx->pc_attention_tc = e[0] != '0';
    if (const char* e = getenv("GENVC_VW_MIN_TOKENS")) ctx->vw_min_tokens = atoi(e);
    if (const char* e = getenv("GENVC_FUSED_PREFILL")) ctx->use_fused_prefill = e[0] != '0';
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e == cudaSuccess && device >= 0 && device < ndev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->n_sm = prop.multiProcessorCount;
    } else {
        (void)cudaGetLastError();
    }
    // Without a device (CPU-only layout queries) the grid defaults to a B200's 148 SMs.
    ctx->grid = ctx->n_sm > 0 ? ctx->n_sm : 148;
    ctx->sdims = StreamDims{g.n_layer, g.d_model, g.n_audio_vocab, ctx->grid};
    // limits of the fused kernel (decode_mega.cu): D = 128 * {1,2,4,8}; at most 32 units of a phase per CTA
    // (two per consumer warp); attention items (n_head * 8) and the partial-sum gather fit the grid / scratch
    auto ceil_div = [](int a, int b) { return (a + b - 1) / b; };
    const int nxv = g.d_model / 128;
    ctx->mega_ok = (nxv == 1 || nxv == 2 || nxv == 4 || nxv == 8) && ceil_div(4 * g.d_model, ctx->grid) <= 32 &&
                   ceil_div(g.n_audio_vocab, ctx->grid) <= 32 && g.n_audio_vocab <= 2048 && g.n_head * 8 <= ctx->grid &&
                   g.d_model / 8 <= ctx->grid && ctx->grid >= 128 && ctx->grid <= 304 &&
                   mega_smem_bytes(g.d_model, (g.n_audio_vocab + 15) / 16 * 16) <= 232448 &&
                   (uint64_t)g.max_gen_mel_tokens * ((uint64_t)GV_TAGS_PER_LAYER * g.n_layer + 1ull) < 0x7FFFFFFFull;
    // batched fused kernel (decode_batch.cu): rows * H attention items must fit the grid (checked per call), the
    // reducer CTAs (D / 8) and the 8-column residual stash per row must fit
    ctx->batch_ok = ctx->mega_ok && ceil_div(g.d_model, ctx->grid) <= 8 &&
                    batch_smem_bytes(g.d_model, (g.n_audio_vocab + 15) / 16 * 16) <= 232448 &&
                    (uint64_t)g.max_gen_mel_tokens * ((uint64_t)GV_TAGS_PER_LAYER * g.n_layer + 1ull + GV_BATCH_TAGS_EXTRA) < 0x7FFFFFFFull;
    plan_workspace(ctx);
    *out = ctx;
    return GENVC_OK;
}

void genvc_destroy(genvc_ctx* ctx) {
    if (!ctx) return;
    if (ctx->blob) gemm_tc_forget(ctx->blob, ctx->blob + ctx->layout.total);
    ctx->drop_graphs();
    delete ctx;
}

const char* genvc_last_error(const genvc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int genvc_decode_grid(const genvc_ctx* ctx) { return ctx ? ctx->grid : 0; }

int genvc_fused_max_rows(const genvc_ctx* ctx) {
    if (!ctx || !ctx->mega_ok) return 0;
    if (!ctx->batch_ok || ctx->fused_rows <= 1) return 1;
    return std::min({(int)GV_BATCH_ROWS, ctx->cfg.max_batch, ctx->grid / ctx->cfg.n_head});
}

uint64_t genvc_blob_floats(const genvc_ctx* ctx) { return ctx ? ctx->layout.total : 0; }
int genvc_num_tensors(const genvc_ctx* ctx) { return ctx ? (int)ctx->layout.tensors.size() : 0; }

int genvc_tensor_name(const genvc_ctx* ctx, int index, char* buf, size_t buf_len) {
    if (!ctx || !buf || index < 0 || index >= (int)ctx->layout.tensors.size()) return GENVC_E_INVALID;
    const std::string& n = ctx->layout.tensors[index].name;
    if (n.size() + 1 > buf_len) return GENVC_E_INVALID;
    memcpy(buf, n.c_str(), n.size() + 1);
    return GENVC_OK;
}

int genvc_tensor_info(const genvc_ctx* ctx, const char* key, uint64_t* offset, uint64_t* rows, uint64_t* cols,
                      uint64_t* row_stride) {
    if (!ctx || !key) return GENVC_E_INVALID;
    auto it = ctx->layout.index.find(key);
    if (it == ctx->layout.index.end()) return ctx->fail(GENVC_E_INVALID, "unknown tensor '%s'", key);
    const TensorEntry& t = ctx->layout.tensors[it->second];
    if (offset) *offset = t.off;
    if (rows) *rows = t.rows;
    if (cols) *cols = t.cols;
    if (row_stride) *row_stride = t.stride;
    return GENVC_OK;
}

// the dense matrices the batched (prefill / latent pass / perceiver) GEMMs stream: {offset, N, K, ldw, w_nk}
struct TcMat {
    uint64_t off;
    int N, K, ldw, w_nk;
};
static std::vector<TcMat> tc_matrices(const genvc_ctx* c) {
    const genvc_config& g = c->cfg;
    const int D = g.d_model, inner = g.pc_dim_head * g.pc_heads, ffi = g.pc_ff_inner;
    std::vector<TcMat> v;
    for (const LayerOff& o : c->layout.layers) {
        v.push_back({o.attn_w, 3 * D, D, 3 * D, 0});
        v.push_back({o.proj_w, D, D, D, 0});
        v.push_back({o.fc_w, 4 * D, D, 4 * D, 0});
        v.push_back({o.proj2_w, D, 4 * D, D, 0});
    }
    const int ffp = (int)c->layout.pc_ff_inner_pad, cp = (int)c->layout.pc_ctx_pad;
    v.push_back({c->layout.pc_proj_w, D, cp, cp, 1});  // proj_context: K = 80 zero-padded to 96
    for (const PcLayerOff& o : c->layout.pc_layers) {
        v.push_back({o.to_q, inner, D, D, 1});
        v.push_back({o.to_kv, 2 * inner, D, D, 1});
        v.push_back({o.to_out, D, inner, inner, 1});
        v.push_back({o.ff0_w, 2 * ffi, D, D, 1});
        v.push_back({o.ff2_w, D, ffp, ffp, 1});        // FF 2: K = 2730 zero-padded to 2752
    }
    std::vector<TcMat> ok;
    for (const TcMat& m : v)
        if (gemm_tc_packed_floats(m.N, m.K) > 0) ok.push_back(m);
    return ok;
}

uint64_t genvc_tc_floats(const genvc_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t t = 0;
    for (const TcMat& m : tc_matrices(ctx)) t += gemm_tc_packed_floats(m.N, m.K);
    return t;
}

int genvc_pack_tc(genvc_ctx* ctx, float* tc_dev, uint64_t n_floats, void* stream) {
    if (!ctx) return GENVC_E_INVALID;
    if (!ctx->blob) return ctx->fail(GENVC_E_STATE, "bind weights first");
    if (!tc_dev || n_floats < genvc_tc_floats(ctx) || reinterpret_cast<uintptr_t>(tc_dev) % 128)
        return ctx->fail(GENVC_E_INVALID, "tensor-core weight buffer too small or misaligned");
    DevGuard guard(ctx->device);
    CK(guard.err);
    uint64_t o = 0;
    ctx->tc_table.clear();
    ctx->tc_table_uploaded = false;
    const size_t n_block_mats = (size_t)ctx->cfg.n_layer * 4;
    const std::vector<TcMat> mats = tc_matrices(ctx);
    // (tc_matrices lists the four matrices of every block first; the table is only valid if none of them was skipped)
    const bool blocks_complete = mats.size() >= n_block_mats && mats[0].off == ctx->layout.layers[0].attn_w &&
                                 mats[n_block_mats - 1].off == ctx->layout.layers[ctx->cfg.n_layer - 1].proj2_w;
    for (size_t i = 0; i < mats.size(); ++i) {
        const TcMat& m = mats[i];
        CK(gemm_tc_pack(ctx->w(m.off), m.N, m.K, m.ldw, m.w_nk, tc_dev + o, (cudaStream_t)stream));
        ctx->nlaunch += 1;
        if (blocks_complete && i < n_block_mats) ctx->tc_table.push_back(tc_dev + o);
        o += gemm_tc_packed_floats(m.N, m.K);
    }
    return GENVC_OK;
}

int genvc_bind_weights(genvc_ctx* ctx, const float* blob_dev, uint64_t n_floats) {
    if (!ctx) return GENVC_E_INVALID;
    if (!blob_dev || n_floats < ctx->layout.total) return ctx->fail(GENVC_E_INVALID, "weight blob too small");
    if (reinterpret_cast<uintptr_t>(blob_dev) % 128) return ctx->fail(GENVC_E_INVALID, "weight blob must be 128-byte aligned");
    if (ctx->blob) gemm_tc_forget(ctx->blob, ctx->blob + ctx->layout.total);
    gemm_tc_forget(blob_dev, blob_dev + ctx->layout.total);
    ctx->blob = blob_dev;
    ctx->stream_packed = false;
    ctx->tc_table.clear();
    ctx->tc_table_uploaded = false;
    ctx->drop_graphs();
    return GENVC_OK;
}

uint64_t genvc_stream_floats(const genvc_ctx* ctx) {
    if (!ctx || !ctx->mega_ok) return 0;
    return (uint64_t)stream_total_floats(ctx->sdims);
}

int genvc_pack_stream(genvc_ctx* ctx, float* stream_dev, uint64_t n_floats, void* stream) {
    if (!ctx) return GENVC_E_INVALID;
    if (!ctx->mega_ok) return ctx->fail(GENVC_E_UNSUPPORTED, "fused decode kernel does not support this shape");
    if (!ctx->blob) return ctx->fail(GENVC_E_STATE, "bind weights first");
    if (!stream_dev || n_floats < genvc_stream_floats(ctx) || reinterpret_cast<uintptr_t>(stream_dev) % 128)
        return ctx->fail(GENVC_E_INVALID, "decode stream buffer too small or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    DevGuard guard(ctx->device);
    CK(guard.err);
    const Layout& L = ctx->layout;
    for (int l = 0; l < ctx->cfg.n_layer; ++l) {
        const LayerOff& o = L.layers[l];
        // ln_1 / ln_2 are folded into the matrices that consume them (decode_mega.cu: pack_stream_kernel)
        CK(launch_pack_stream(ctx->sdims, l, PH_QKV, ctx->w(o.attn_w), ctx->w(o.attn_b), 0, ctx->w(o.ln1_w), ctx->w(o.ln1_b),
                              stream_dev, st));
        CK(launch_pack_stream(ctx->sdims, l, PH_PROJ, ctx->w(o.proj_w), ctx->w(o.proj_b), 0, nullptr, nullptr, stream_dev, st));
        CK(launch_pack_stream(ctx->sdims, l, PH_FC, ctx->w(o.fc_w), ctx->w(o.fc_b), 0, ctx->w(o.ln2_w), ctx->w(o.ln2_b), stream_dev,
                              st));
        // mlp.c_proj [4D, D] is split along K: unit k = row k (its bias is added by the reducer CTAs)
        CK(launch_pack_stream(ctx->sdims, l, PH_P2, ctx->w(o.proj2_w), nullptr, 1, nullptr, nullptr, stream_dev, st));
        ctx->nlaunch += 4;
    }
    CK(launch_pack_stream(ctx->sdims, 0, PH_HEAD, ctx->w(L.mel_head_w), ctx->w(L.mel_head_b), 1, nullptr, nullptr, stream_dev, st));
    ctx->nlaunch += 1;
    ctx->stream = stream_dev;
    ctx->stream_packed = true;
    return GENVC_OK;
}

uint64_t genvc_vw_floats(const genvc_ctx* ctx) {
    if (!ctx || !ctx->mega_ok || ctx->cfg.n_head > 32 || (ctx->cfg.d_model + ctx->grid - 1) / ctx->grid > 8) return 0;
    return (uint64_t)ctx->cfg.n_layer * ctx->grid * ctx->cfg.max_seq * ctx->cfg.n_head * 8;
}

int genvc_bind_vw(genvc_ctx* ctx, float* vw_dev, uint64_t n_floats) {
    if (!ctx) return GENVC_E_INVALID;
    if (vw_dev == nullptr) {  // unbind: the fused kernel goes back to K / V attention items
        ctx->vw = nullptr;
    } else {
        const uint64_t need = genvc_vw_floats(ctx);
        if (need == 0) return ctx->fail(GENVC_E_UNSUPPORTED, "no projected-value cache for this shape");
        if (n_floats < need || reinterpret_cast<uintptr_t>(vw_dev) % 128)
            return ctx->fail(GENVC_E_INVALID, "projected-value cache too small or misaligned");
        ctx->vw = vw_dev;
    }
    ctx->prefilled = false;
    ctx->drop_graphs();
    return GENVC_OK;
}

uint64_t genvc_kv_floats(const genvc_ctx* ctx) { return ctx ? (uint64_t)ctx->cfg.n_layer * 2 * ctx->kv_plane() : 0; }
uint64_t genvc_workspace_bytes(const genvc_ctx* ctx) { return ctx ? ctx->ws_bytes : 0; }

int genvc_bind_buffers(genvc_ctx* ctx, float* kv_dev, uint64_t kv_floats, void* workspace_dev, uint64_t workspace_bytes) {
    if (!ctx) return GENVC_E_INVALID;
    if (!kv_dev || kv_floats < genvc_kv_floats(ctx)) return ctx->fail(GENVC_E_INVALID, "KV cache buffer too small");
    if (!workspace_dev || workspace_bytes < ctx->ws_bytes) return ctx->fail(GENVC_E_INVALID, "workspace too small");
    if (reinterpret_cast<uintptr_t>(kv_dev) % 128 || reinterpret_cast<uintptr_t>(workspace_dev) % 256)
        return ctx->fail(GENVC_E_INVALID, "KV cache / workspace misaligned");
    ctx->kv = kv_dev;
    ctx->ws = static_cast<char*>(workspace_dev);
    ctx->prefilled = false;
    ctx->drop_graphs();
    // exchange tags start at 1 over zeroed buffers
    DevGuard guard(ctx->device);
    CK(guard.err);
    CK(cudaMemset(ctx->ws + ctx->o_xchg, 0, ctx->xchg_bytes));
    CK(cudaMemset(ctx->ws + ctx->o_flags, 0, 64));
    CK(cudaDeviceSynchronize());
    ctx->tag_next = 1;
    return GENVC_OK;
}

uint64_t genvc_launch_count(const genvc_ctx* ctx) { return ctx ? ctx->nlaunch : 0; }

int genvc_debug_tune(genvc_ctx* ctx, int window, int nosync, int l2_ahead_tiles, int hop_settle_ns, int hop_hold) {
    if (!ctx) return GENVC_E_INVALID;
    if (window > 0) ctx->window = std::min(window, (int)GV_MEGA_NSLOT);
    if (l2_ahead_tiles >= 0) ctx->l2_ahead = l2_ahead_tiles;
    if (hop_settle_ns >= 0) ctx->hop_settle = hop_settle_ns;
    if (hop_hold >= 1000) {  // 1000 + near + 100 * near_ao: explicit early-release margins (1000 = none)
        ctx->hop_near = (hop_hold - 1000) % 100;
        ctx->hop_near_ao = (hop_hold - 1000) / 100;
    } else if (hop_hold >= 0) {
        ctx->hop_hold = hop_hold;
    }
    ctx->dbg_nosync = nosync ? 1 : 0;
    return GENVC_OK;
}

int genvc_debug_layout(const genvc_ctx* ctx, uint64_t* out, int n) {
    if (!ctx || !out || n < 12) return GENVC_E_INVALID;
    const uint64_t v[12] = {ctx->o_xq, ctx->o_matt_o, ctx->o_matt_ml, ctx->o_x1, ctx->o_pp, ctx->o_x2, ctx->o_lg, ctx->o_hops,
                            ctx->o_sbuf, (uint64_t)ctx->grid, (uint64_t)GV_HOP_STRIDE, (uint64_t)HC_COUNT};
    for (int i = 0; i < 12; ++i) out[i] = v[i];
    return GENVC_OK;
}

int genvc_debug_trace(genvc_ctx* ctx, uint64_t* trace_dev, int slots_per_cta, int step) {
    if (!ctx) return GENVC_E_INVALID;
    if (trace_dev && (slots_per_cta <= 0 || step < 0)) return ctx->fail(GENVC_E_INVALID, "bad trace geometry");
    ctx->trace = reinterpret_cast<unsigned long long*>(trace_dev);
    ctx->trace_slots = trace_dev ? slots_per_cta : 0;
    ctx->trace_step = step;
    return GENVC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// helpers shared by prefill / latent pass / per-op decode
// ---------------------------------------------------------------------------------------------
static int check_ready(genvc_ctx* ctx) {
    if (!ctx) return GENVC_E_INVALID;
    if (!ctx->blob) return ctx->fail(GENVC_E_STATE, "weights not bound");
    if (!ctx->kv || !ctx->ws) return ctx->fail(GENVC_E_STATE, "buffers not bound");
    return GENVC_OK;
}

static GemmArgs gemm(const float* A, int lda, const float* W, int ldw, int w_nk, const float* bias, const float* res, int ldr,
                     float* C, int ldc, int M, int N, int K, int act) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.w_nk = w_nk; a.bias = bias; a.residual = res; a.ldr = ldr;
    a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.act = act;
    return a;
}

// The 30 pre-LN blocks over rows X [B, M, D] (contiguous, R = B*M rows).
//   pos0 >= 0: rows are positions pos0..pos0+M-1 of the KV cache: K/V are appended and attention runs
//              against the cache (decode, M == 1) or over the rows themselves (prefill, pos0 == 0);
//   pos0 <  0: uncached causal pass (teacher-forced latent pass).
static int run_blocks(genvc_ctx* ctx, int B, int M, int pos0, const int* skip, cudaStream_t st) {
    const genvc_config& g = ctx->cfg;
    const int D = g.d_model, H = g.n_head, hd = D / H, R = B * M;
    float* X = ctx->at<float>(ctx->o_X);
    float* A = ctx->at<float>(ctx->o_A);
    float* QKV = ctx->at<float>(ctx->o_QKV);
    float* U = ctx->at<float>(ctx->o_U);
    float* sk = ctx->at<float>(ctx->o_splitk);
    unsigned long long* nl = &ctx->nlaunch;
    const float scale = 1.0f / sqrtf((float)hd);
    for (int l = 0; l < g.n_layer; ++l) {
        const LayerOff& o = ctx->layout.layers[l];
        float* kc = ctx->kv + ((size_t)l * 2 + 0) * ctx->kv_plane();
        float* vc = ctx->kv + ((size_t)l * 2 + 1) * ctx->kv_plane();
        // ln_1: first layer only; afterwards it rides in the epilogue of the previous block's mlp.c_proj GEMM
        if (l == 0)
            CK(launch_layernorm(X, D, 0, A, D, 0, R, R, D, ctx->w(o.ln1_w), ctx->w(o.ln1_b), nullptr, nullptr, skip, st, nl));
        {   // [q|k|v] projection; the K/V cache append rides in its epilogue
            GemmArgs q = gemm(A, D, ctx->w(o.attn_w), 3 * D, 0, ctx->w(o.attn_b), nullptr, 0, QKV, 3 * D, R, 3 * D, D, ACT_NONE);
            if (pos0 >= 0) {
                q.kv_k = kc; q.kv_v = vc; q.kv_bs = (long)ctx->kv_batch_stride(); q.kv_rows = M; q.kv_H = H; q.kv_S_max = g.max_seq;
                q.kv_pos0 = pos0;
            }
            CK(launch_gemm(q, sk, kSplitKFloats, skip, st, nl));
        }
        if (pos0 >= 0 && M == 1) {
            // single-token decode against the cache (positions 0..pos0)
            CK(launch_kv_attention(QKV, 3L * D, kc, vc, (long)ctx->kv_batch_stride(), B, H, hd, pos0 + 1, g.max_seq, A, D, skip,
                                   st, nl));
        } else {
            AttnArgs a;
            a.Q = QKV;         a.q_bs = (long)M * 3 * D; a.q_rs = 3 * D; a.q_hs = hd;
            a.K = QKV + D;     a.k_bs = (long)M * 3 * D; a.k_rs = 3 * D; a.k_hs = hd;
            a.V = QKV + 2 * D; a.v_bs = (long)M * 3 * D; a.v_rs = 3 * D; a.v_hs = hd;
            a.O = A;           a.o_bs = (long)M * D;     a.o_rs = D;     a.o_hs = hd;
            a.B = B; a.H = H; a.M = M; a.hd = hd; a.n_keys = M; a.causal = 1; a.pos0 = 0; a.scale = scale;
            CK(launch_attention(a, skip, st, nl));
        }
        {   // attention output projection + residual; ln_2 of the result rides in the epilogue
            GemmArgs pj = gemm(A, D, ctx->w(o.proj_w), D, 0, ctx->w(o.proj_b), X, D, X, D, R, D, D, ACT_NONE);
            pj.ln_w = ctx->w(o.ln2_w); pj.ln_b = ctx->w(o.ln2_b); pj.ln_out = A; pj.ld_ln = D;
            CK(launch_gemm(pj, sk, kSplitKFloats, skip, st, nl));
        }
        CK(launch_gemm(gemm(A, D, ctx->w(o.fc_w), 4 * D, 0, ctx->w(o.fc_b), nullptr, 0, U, 4 * D, R, 4 * D, D, ACT_GELU_NEW), sk,
                       kSplitKFloats, skip, st, nl));
        {   // mlp.c_proj + residual; the next block's ln_1 rides in the epilogue
            GemmArgs p2 = gemm(U, 4 * D, ctx->w(o.proj2_w), D, 0, ctx->w(o.proj2_b), X, D, X, D, R, D, 4 * D, ACT_NONE);
            if (l + 1 < g.n_layer) {
                const LayerOff& nx = ctx->layout.layers[l + 1];
                p2.ln_w = ctx->w(nx.ln1_w); p2.ln_b = ctx->w(nx.ln1_b); p2.ln_out = A; p2.ld_ln = D;
            }
            CK(launch_gemm(p2, sk, kSplitKFloats, skip, st, nl));
        }
    }
    return GENVC_OK;
}

// ln_f -> final_norm of row `row` of each batch element -> pending latent; mel_head -> pending logits
static int run_head(genvc_ctx* ctx, int B, int M, int row, const int* skip, cudaStream_t st) {
    const genvc_config& g = ctx->cfg;
    const int D = g.d_model, V = g.n_audio_vocab;
    const Layout& L = ctx->layout;
    float* X = ctx->at<float>(ctx->o_X);
    float* lat = ctx->at<float>(ctx->o_pend_latent);
    float* lg = ctx->at<float>(ctx->o_pend_logits);
    CK(launch_layernorm(X + (size_t)row * D, D, (long)M * D, lat, D, D, B, 1, D, ctx->w(L.lnf_w), ctx->w(L.lnf_b), ctx->w(L.fn_w),
                        ctx->w(L.fn_b), skip, st, &ctx->nlaunch));
    CK(launch_gemm(gemm(lat, D, ctx->w(L.mel_head_w), D, 1, ctx->w(L.mel_head_b), nullptr, 0, lg, V, B, V, D, ACT_NONE),
                   ctx->at<float>(ctx->o_splitk), kSplitKFloats, skip, st, &ctx->nlaunch));
    return GENVC_OK;
}

// Runs `body` (a sequence of launches on `st` that only touches context-owned buffers) -- eagerly the first time a key is
// seen (sets kernel attributes, warms caches), captured into a CUDA graph on a second pass, replayed from then on.
template <class Body>
static int run_or_replay(genvc_ctx* ctx, std::map<std::pair<int, int>, genvc_ctx::PrefillGraph>& cache, std::pair<int, int> key,
                         cudaStream_t st, Body body) {
    if (!ctx->use_graphs) return body();
    auto it = cache.find(key);
    if (it != cache.end()) {
        if (it->second.exec == nullptr) return body();  // capture failed once for this shape: per-op launches
        CK(cudaGraphLaunch(it->second.exec, st));
        ctx->nlaunch += it->second.launches;
        return GENVC_OK;
    }
    if (int rc = body()) return rc;
    genvc_ctx::PrefillGraph pg;
    const unsigned long long l0 = ctx->nlaunch;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        const int rc = body();
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        ok = rc == GENVC_OK && e == cudaSuccess && graph != nullptr;
    }
    pg.launches = ctx->nlaunch - l0;
    ctx->nlaunch = l0;  // nothing ran during the capture
    if (ok) ok = cudaGraphInstantiate(&pg.exec, graph, 0) == cudaSuccess;
    if (graph) (void)cudaGraphDestroy(graph);
    if (!ok) {
        (void)cudaGetLastError();
        pg.exec = nullptr;
    }
    if (cache.size() >= 64) {  // bounded cache
        for (auto& kv : cache)
            if (kv.second.exec) (void)cudaGraphExecDestroy(kv.second.exec);
        cache.clear();
    }
    cache[key] = pg;
    return GENVC_OK;  // the eager pass above did the work
}

extern "C" {

// ---------------------------------------------------------------------------------------------
// perceiver
// ---------------------------------------------------------------------------------------------
int genvc_perceiver(genvc_ctx* ctx, const float* mel_dev, int B, int S_mel, float* latents_out_dev, void* stream) {
    if (int rc = check_ready(ctx)) return rc;
    DevGuard guard(ctx->device);
    CK(guard.err);
    const genvc_config& g = ctx->cfg;
    if (!mel_dev || !latents_out_dev) return ctx->fail(GENVC_E_INVALID, "null pointer");
    if (B <= 0 || B > g.max_batch) return ctx->fail(GENVC_E_INVALID, "batch %d outside [1, %d]", B, g.max_batch);
    if (S_mel <= 0 || S_mel > g.max_mel_frames)
        return ctx->fail(GENVC_E_INVALID, "mel frames %d outside [1, %d]", S_mel, g.max_mel_frames);
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = ctx->layout;
    const int D = g.d_model, C = g.pc_dim_context, NL = g.pc_latents, inner = g.pc_dim_head * g.pc_heads;
    const int ffi = g.pc_ff_inner, ffp = (int)L.pc_ff_inner_pad;
    const int RC = NL + S_mel;  // context rows per batch element
    float* melT = ctx->at<float>(ctx->o_pc_melT);
    float* cx = ctx->at<float>(ctx->o_pc_ctx);
    float* kvb = ctx->at<float>(ctx->o_pc_kv);
    float* lat = ctx->at<float>(ctx->o_pc_lat);
    float* q = ctx->at<float>(ctx->o_pc_q);
    float* ob = ctx->at<float>(ctx->o_pc_o);
    float* hb = ctx->at<float>(ctx->o_pc_h);
    float* gb = ctx->at<float>(ctx->o_pc_g);
    float* sk = ctx->at<float>(ctx->o_splitk);
    unsigned long long* nl = &ctx->nlaunch;

    const int Cp = (int)L.pc_ctx_pad;  // mel channels zero-padded to the k granularity of the tensor-core GEMM
    CK(launch_transpose_mel(mel_dev, B, C, Cp, S_mel, melT, st, nl));  // (caller's pointer: outside the replayed part)
    auto body = [&]() -> int {
    for (int b = 0; b < B; ++b)  // proj_context into rows NL.. of this element's context block
        CK(launch_gemm(gemm(melT + (size_t)b * S_mel * Cp, Cp, ctx->w(L.pc_proj_w), Cp, 1, ctx->w(L.pc_proj_b), nullptr, 0,
                            cx + ((size_t)b * RC + NL) * D, D, S_mel, D, Cp, ACT_NONE),
                       sk, kSplitKFloats, nullptr, st, nl));
    CK(launch_copy_rows(ctx->w(L.pc_latents), 0, lat, (long)NL * D, B, (long)NL * D, st, nl));  // repeat(latents)
    for (int i = 0; i < g.pc_depth; ++i) {
        const PcLayerOff& o = L.pc_layers[i];
        CK(launch_copy_rows(lat, (long)NL * D, cx, (long)RC * D, B, (long)NL * D, st, nl));  // ctx = [latents ; x]
        CK(launch_gemm(gemm(lat, D, ctx->w(o.to_q), D, 1, nullptr, nullptr, 0, q, inner, B * NL, inner, D, ACT_NONE), sk,
                       kSplitKFloats, nullptr, st, nl));
        CK(launch_gemm(gemm(cx, D, ctx->w(o.to_kv), D, 1, nullptr, nullptr, 0, kvb, 2 * inner, B * RC, 2 * inner, D, ACT_NONE), sk,
                       kSplitKFloats, nullptr, st, nl));
        if (ctx->pc_attention_tc && pc_attention_tc_supported(NL, g.pc_dim_head, RC)) {
            // cross-attention on the tensor cores: scores, softmax and P.V of one (element, head) stay on one SM
            CK(launch_pc_attention_tc(q, kvb, ob, B, g.pc_heads, RC, st));
            ctx->nlaunch += 1;
        } else {
            AttnArgs a;
            a.Q = q;           a.q_bs = (long)NL * inner;     a.q_rs = inner;     a.q_hs = g.pc_dim_head;
            a.K = kvb;         a.k_bs = (long)RC * 2 * inner; a.k_rs = 2 * inner; a.k_hs = g.pc_dim_head;
            a.V = kvb + inner; a.v_bs = (long)RC * 2 * inner; a.v_rs = 2 * inner; a.v_hs = g.pc_dim_head;
            a.O = ob;          a.o_bs = (long)NL * inner;     a.o_rs = inner;     a.o_hs = g.pc_dim_head;
            a.B = B; a.H = g.pc_heads; a.M = NL; a.hd = g.pc_dim_head; a.n_keys = RC; a.causal = 0; a.pos0 = 0;
            a.scale = 1.0f / sqrtf((float)g.pc_dim_head);
            CK(launch_attention(a, nullptr, st, nl));
        }
        CK(launch_gemm(gemm(ob, inner, ctx->w(o.to_out), inner, 1, nullptr, lat, D, lat, D, B * NL, D, inner, ACT_NONE), sk,
                       kSplitKFloats, nullptr, st, nl));
        CK(launch_gemm(gemm(lat, D, ctx->w(o.ff0_w), D, 1, ctx->w(o.ff0_b), nullptr, 0, hb, 2 * ffi, B * NL, 2 * ffi, D, ACT_NONE),
                       sk, kSplitKFloats, nullptr, st, nl));
        CK(launch_geglu(hb, B * NL, ffi, ffp, gb, st, nl));
        CK(launch_gemm(gemm(gb, ffp, ctx->w(o.ff2_w), ffp, 1, ctx->w(o.ff2_b), lat, D, lat, D, B * NL, D, ffp, ACT_NONE), sk,
                       kSplitKFloats, nullptr, st, nl));
    }
    return GENVC_OK;
    };
    if (int rc = run_or_replay(ctx, ctx->perceiver_graphs, std::make_pair(B, S_mel), st, body)) return rc;
    CK(launch_rmsnorm(lat, latents_out_dev, B * NL, D, ctx->w(L.pc_gamma), st, nl));
    return GENVC_OK;
}

// ---------------------------------------------------------------------------------------------
// prefix embedding + prefill
// ---------------------------------------------------------------------------------------------
int genvc_embed_prefix(genvc_ctx* ctx, const float* cond_dev, const int64_t* text_ids_dev, int B, int T, float* prefix_out_dev,
                       void* stream) {
    if (int rc = check_ready(ctx)) return rc;
    DevGuard guard(ctx->device);
    CK(guard.err);
    const genvc_config& g = ctx->cfg;
    if (!cond_dev || !text_ids_dev || !prefix_out_dev) return ctx->fail(GENVC_E_INVALID, "null pointer");
    if (B <= 0 || T < 0 || T + 2 > g.n_text_pos)
        return ctx->fail(GENVC_E_INVALID, "text length %d exceeds the %d text positions", T, g.n_text_pos - 2);
    const Layout& L = ctx->layout;
    const int P = g.pc_latents + T + 2;
    CK(launch_embed_prefix(cond_dev, reinterpret_cast<const long long*>(text_ids_dev), B, T, g.pc_latents, g.d_model,
                           ctx->w(L.text_emb), ctx->w(L.text_pos), g.start_text, g.stop_text, prefix_out_dev, (long)P * g.d_model,
                           g.n_text_vocab, ctx->at<int>(ctx->o_flags), (cudaStream_t)stream, &ctx->nlaunch));
    return GENVC_OK;
}

int genvc_prefill(genvc_ctx* ctx, const float* prefix_dev, int B, int P, void* stream) {
    if (int rc = check_ready(ctx)) return rc;
    DevGuard guard(ctx->device);
    CK(guard.err);
    const genvc_config& g = ctx->cfg;
    if (!prefix_dev) return ctx->fail(GENVC_E_INVALID, "null pointer");
    if (B <= 0 || B > g.max_batch) return ctx->fail(GENVC_E_INVALID, "batch %d outside [1, %d]", B, g.max_batch);
    if (P <= 0 || P + 1 + g.max_gen_mel_tokens > g.max_seq)
        return ctx->fail(GENVC_E_INVALID, "prefix %d + %d generated tokens exceed the KV cache (%d positions)", P,
                         g.max_gen_mel_tokens, g.max_seq);
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = ctx->layout;
    const int D = g.d_model, M = P + 1;
    float* X = ctx->at<float>(ctx->o_X);
    // rows = [prefix (P) ; mel_embedding[start_audio] + mel_pos[0]]   (layers/gpt_inference.py:81-91)
    CK(launch_copy_rows(prefix_dev, (long)P * D, X, (long)M * D, B, (long)P * D, st, &ctx->nlaunch));
    CK(launch_embed_mel_rows(nullptr, B, 1, 0, g.start_audio, g.stop_audio, 0, D, ctx->w(L.mel_emb), ctx->w(L.mel_pos),
                             X + (size_t)P * D, (long)M * D, g.n_audio_vocab, ctx->at<int>(ctx->o_flags), st, &ctx->nlaunch));
    // the generation state lives in copy 0 after a prefill (the fused decode launches flip between the two copies)
    ctx->gs_cur = 0;
    // ---- persistent fused prefill (batch 1, <= 128 rows, tensor-core weights packed): one cooperative launch for the
    // 30 blocks instead of nine launches per block
    if (ctx->use_fused_prefill && B == 1 && ctx->tc_table.size() == (size_t)g.n_layer * 4 && ctx->n_sm == ctx->grid &&
        prefill_fused_supported(D, g.n_head, M, ctx->grid, kSplitKFloats)) {
        if (!ctx->tc_table_uploaded) {
            CK(cudaMemcpyAsync(ctx->ws + ctx->o_tcptr, ctx->tc_table.data(), ctx->tc_table.size() * sizeof(const float*),
                               cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));  // (pageable source; once per weight binding)
            ctx->tc_table_uploaded = true;
        }
        const LayerOff& o0 = L.layers[0];
        // ln_1 of the first block (the later ones ride in the row-wise reductions of the fused kernel)
        CK(launch_layernorm(X, D, 0, ctx->at<float>(ctx->o_A), D, 0, M, M, D, ctx->w(o0.ln1_w), ctx->w(o0.ln1_b), nullptr, nullptr,
                            nullptr, st, &ctx->nlaunch));
        PrefillFusedArgs fa;
        fa.L = g.n_layer; fa.D = D; fa.H = g.n_head; fa.M = M; fa.S_max = g.max_seq;
        fa.blob = ctx->blob;
        fa.layer_stride = g.n_layer > 1 ? (long long)(L.layers[1].ln1_w - o0.ln1_w) : 0;
        fa.ln1_w = (long long)o0.ln1_w; fa.ln1_b = (long long)o0.ln1_b; fa.attn_b = (long long)o0.attn_b;
        fa.proj_b = (long long)o0.proj_b; fa.ln2_w = (long long)o0.ln2_w; fa.ln2_b = (long long)o0.ln2_b;
        fa.fc_b = (long long)o0.fc_b; fa.proj2_b = (long long)o0.proj2_b;
        fa.tcw = ctx->at<const float*>(ctx->o_tcptr);
        fa.X = X; fa.A = ctx->at<float>(ctx->o_A); fa.QKV = ctx->at<float>(ctx->o_QKV); fa.U = ctx->at<float>(ctx->o_U);
        fa.ws = ctx->at<float>(ctx->o_splitk);
        fa.kv = ctx->kv; fa.kv_layer_stride = (long long)ctx->kv_plane();
        fa.gbar = ctx->at<unsigned>(ctx->o_gbar);
        fa.prof = (ctx->trace != nullptr && ctx->trace_slots >= 16) ? ctx->trace : nullptr;  // debug hook shared with the decode kernels
        if (const char* e = getenv("GENVC_PREFILL_DBG")) fa.dbg = atoi(e);
        CK(cudaMemsetAsync(fa.gbar, 0, sizeof(unsigned), st));
        CK(launch_prefill_fused(fa, ctx->grid, st));
        ctx->nlaunch += 1;
        if (int rc = run_head(ctx, B, M, P, nullptr, st)) return rc;
        CK(launch_init_state(ctx->gstate(0), ctx->gseen(0), B, P, g.n_audio_vocab, ctx->Vpad, g.start_audio, st, &ctx->nlaunch));
        ctx->B = B;
        ctx->P = P;
        ctx->prefilled = true;
        ctx->pending = true;
        ctx->vw_filled = false;
        ctx->n_host = 0;
        return GENVC_OK;
    }
    auto body = [&]() -> int {
        if (int rc = run_blocks(ctx, B, M, 0, nullptr, st)) return rc;
        if (int rc = run_head(ctx, B, M, P, nullptr, st)) return rc;
        CK(launch_init_state(ctx->gstate(0), ctx->gseen(0), B, P, g.n_audio_vocab, ctx->Vpad, g.start_audio, st, &ctx->nlaunch));
        return GENVC_OK;
    };
    bool done = false;
    // (the tcgen05 GEMM path must be registered before the first capture: launch_gemm_tc sets function attributes once)
    if (ctx->use_graphs) {
        const auto key = std::make_pair(B, P);
        auto it = ctx->prefill_graphs.find(key);
        if (it == ctx->prefill_graphs.end()) {
            // first use of this shape: run it eagerly once (sets kernel attributes, warms caches), then capture a second pass
            if (int rc = body()) return rc;
            genvc_ctx::PrefillGraph pg;
            const unsigned long long l0 = ctx->nlaunch;
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                const int rc = body();
                const cudaError_t e = cudaStreamEndCapture(st, &graph);
                ok = rc == GENVC_OK && e == cudaSuccess && graph != nullptr;
            }
            pg.launches = ctx->nlaunch - l0;
            ctx->nlaunch = l0;  // nothing ran during the capture
            if (ok) ok = cudaGraphInstantiate(&pg.exec, graph, 0) == cudaSuccess;
            if (graph) (void)cudaGraphDestroy(graph);
            if (!ok) {
                (void)cudaGetLastError();
                pg.exec = nullptr;  // remembered as "do not try again": this shape keeps the per-op launches
            }
            if (ctx->prefill_graphs.size() >= 64) {  // bounded cache
                for (auto& kv : ctx->prefill_graphs)
                    if (kv.second.exec) (void)cudaGraphExecDestroy(kv.second.exec);
                ctx->prefill_graphs.clear();
            }
            ctx->prefill_graphs[key] = pg;
            done = true;  // the eager pass above did the work
        } else if (it->second.exec != nullptr) {
            CK(cudaGraphLaunch(it->second.exec, st));
            ctx->nlaunch += it->second.launches;
            done = true;
        }
    }
    if (!done)
        if (int rc = body()) return rc;
    ctx->B = B;
    ctx->P = P;
    ctx->prefilled = true;
    ctx->pending = true;
    ctx->vw_filled = false;
    ctx->n_host = 0;
    return GENVC_OK;
}

// ---------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------
int genvc_decode(genvc_ctx* ctx, int n_steps, const genvc_sampling* sp, const float* exp_noise_dev, const int64_t* forced_ids_dev,
                 int64_t* ids_out_dev, float* latents_out_dev, float* logits_out_dev, int32_t* status_dev, int mode, void* stream) {
    if (int rc = check_ready(ctx)) return rc;
    DevGuard guard(ctx->device);
    CK(guard.err);
    if (!ctx->prefilled) return ctx->fail(GENVC_E_STATE, "genvc_decode before genvc_prefill");
    if (!sp || !ids_out_dev || !latents_out_dev || !status_dev) return ctx->fail(GENVC_E_INVALID, "null pointer");
    if (n_steps <= 0) return ctx->fail(GENVC_E_INVALID, "n_steps must be positive");
    if (mode < 0 || mode > 2) return ctx->fail(GENVC_E_INVALID, "mode must be 0, 1 or 2");
    if (!(sp->temperature > 0.0f) || !(sp->repetition_penalty > 0.0f) || sp->top_k < 0)
        return ctx->fail(GENVC_E_INVALID, "temperature / repetition_penalty must be > 0 and top_k >= 0");
    const genvc_config& g = ctx->cfg;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = ctx->B, D = g.d_model, V = g.n_audio_vocab;
    int max_total = g.max_gen_mel_tokens;
    if (sp->max_new_tokens > 0) max_total = std::min(max_total, (int)sp->max_new_tokens);
    const bool fused_dev = ctx->stream_packed && ctx->n_sm == ctx->grid;
    const bool can_batch = fused_dev && ctx->batch_ok && B > 1 && B <= GV_BATCH_ROWS && ctx->fused_rows >= B &&
                           B * g.n_head <= ctx->grid;
    const bool can_mega = (fused_dev && ctx->mega_ok && B == 1) || can_batch;
    if (mode == 2 && !can_mega)
        return ctx->fail(GENVC_E_UNSUPPORTED, "fused decode needs batch <= %d (batch * heads <= %d), a packed decode stream and a "
                         "supported shape", GV_BATCH_ROWS, ctx->grid);
    const bool mega = mode == 2 || (mode == 0 && can_mega);
    GenState* gs = ctx->gstate(ctx->gs_cur);
    unsigned char* seen = ctx->gseen(ctx->gs_cur);
    CK(cudaMemsetAsync(status_dev, 0, 4 * sizeof(int32_t), st));

    if (mega) {
        const Layout& L = ctx->layout;
        MegaParams p;
        memset(&p, 0, sizeof p);
        p.L = g.n_layer; p.D = D; p.H = g.n_head; p.V = V; p.Vpad = ctx->Vpad; p.S_max = g.max_seq;
        p.P = ctx->P; p.n_steps = n_steps;
        p.stream = ctx->stream; p.blob = ctx->blob;
        p.proj2_b_off = (long long)L.layers[0].proj2_b;
        p.layer_stride = g.n_layer > 1 ? (long long)(L.layers[1].ln1_w - L.layers[0].ln1_w) : 0;
        p.lnf_off = (long long)L.lnf_w; p.mel_emb_off = (long long)L.mel_emb; p.mel_pos_off = (long long)L.mel_pos;
        p.kv = ctx->kv; p.kv_layer_stride = (long long)ctx->kv_plane();
        p.xq = ctx->at<float>(ctx->o_xq); p.att_o = ctx->at<float>(ctx->o_matt_o); p.att_ml = ctx->at<float>(ctx->o_matt_ml);
        p.x1 = ctx->at<float>(ctx->o_x1); p.pp = ctx->at<float>(ctx->o_pp); p.x2 = ctx->at<float>(ctx->o_x2);
        p.lg = ctx->at<float>(ctx->o_lg);
        p.hops = ctx->at<unsigned>(ctx->o_hops);
        CK(cudaMemsetAsync(p.hops, 0, ctx->hops_bytes, st));
        p.ao = ctx->at<float>(ctx->o_ao); p.att_cnt = ctx->at<unsigned>(ctx->o_attcnt);
        if (ctx->pending || !ctx->vw_filled)  // first fused launch of the sequence (or after per-op steps): choose the variant
            ctx->seg_pvw = B == 1 && ctx->vw != nullptr && max_total - ctx->n_host >= ctx->vw_min_tokens;
        p.vw = ctx->seg_pvw ? ctx->vw : nullptr; p.sbuf = ctx->at<float>(ctx->o_sbuf);
        // the per-op prefill fills K / V only: the first fused forward after it computes the projected values of the cached
        // positions from the V cache (the CTA's attn c_proj columns are in shared memory then anyway)
        const bool runs_forward = n_steps - (ctx->pending ? 1 : 0) >= 1;
        p.vw_fill = (p.vw != nullptr && !ctx->vw_filled && runs_forward) ? 1 : 0;
        if (p.vw_fill) ctx->vw_filled = true;
        p.acc = ctx->at<unsigned long long>(ctx->o_acc);
        CK(cudaMemsetAsync(p.acc, 0, ctx->acc_bytes, st));
        {   // exchange tags: unique per (launch, step, layer, buffer); restart over zeroed buffers before a wrap
            const uint64_t need = (uint64_t)n_steps * ((uint64_t)GV_TAGS_PER_LAYER * g.n_layer + 1ull + GV_BATCH_TAGS_EXTRA);
            if ((uint64_t)ctx->tag_next + need >= 0xFFFFFFF0ull) {
                CK(cudaMemsetAsync(ctx->ws + ctx->o_xchg, 0, ctx->xchg_bytes, st));
                ctx->tag_next = 1;
            }
            p.tag0 = ctx->tag_next;
            ctx->tag_next += (uint32_t)need;
        }
        p.pend_logits = ctx->at<float>(ctx->o_pend_logits); p.pend_latent = ctx->at<float>(ctx->o_pend_latent);
        // the launch reads the current copy of the state and writes the other one
        p.st = gs; p.seen = seen;
        p.st_out = ctx->gstate(ctx->gs_cur ^ 1); p.seen_out = ctx->gseen(ctx->gs_cur ^ 1);
        p.B = B; p.tokx = ctx->at<float>(ctx->o_tokx);
        p.top_k = sp->top_k; p.top_p = sp->top_p; p.top_p_threshold = sp->top_p_threshold; p.temperature = sp->temperature;
        p.rep_penalty = sp->repetition_penalty; p.ignore_eos = sp->ignore_eos; p.stop_token = g.stop_audio;
        p.max_total = max_total; p.seed = sp->seed;
        p.noise = exp_noise_dev; p.forced = reinterpret_cast<const long long*>(forced_ids_dev);
        p.ids_out = reinterpret_cast<long long*>(ids_out_dev); p.latents_out = latents_out_dev; p.logits_out = logits_out_dev;
        p.status = status_dev; p.bad_ids = ctx->at<int>(ctx->o_flags);
        p.trace = ctx->trace; p.trace_slots = ctx->trace_slots; p.trace_step = ctx->trace_step;
        p.window = ctx->window; p.dbg_nosync = ctx->dbg_nosync; p.l2_ahead_tiles = ctx->l2_ahead; p.hop_settle_ns = ctx->hop_settle; p.hop_hold = ctx->hop_hold; p.hop_near = ctx->hop_near; p.hop_near_ao = ctx->hop_near_ao;
        if (B == 1) CK(launch_decode_mega(p, ctx->grid, st));
        else CK(launch_decode_batch(p, ctx->grid, st));
        ctx->gs_cur ^= 1;
        ctx->nlaunch += 1;
        ctx->n_host = std::min(max_total, ctx->n_host + n_steps);
        ctx->pending = false;
        return GENVC_OK;
    }

    // per-op path (any batch): one forward + one sample kernel per step; steps enqueued after the
    // device-side loop has finished are skipped on the device (GenState::done)
    const Layout& L = ctx->layout;
    const int* skip = &gs->done;
    ctx->vw_filled = false;  // per-op steps append K / V only: a later fused launch recomputes the projected values
    for (int i = 0; i < n_steps; ++i) {
        const int n = ctx->n_host;
        if (n >= max_total) break;
        if (!ctx->pending) {
            // forward of the last token at mel position n, cache row P + n   (layers/gpt_inference.py:92-96)
            CK(launch_embed_last_token(gs, B, n, D, ctx->w(L.mel_emb), ctx->w(L.mel_pos), ctx->at<float>(ctx->o_X), st, &ctx->nlaunch));
            if (int rc = run_blocks(ctx, B, 1, ctx->P + n, skip, st)) return rc;
            if (int rc = run_head(ctx, B, 1, 0, skip, st)) return rc;
        }
        SampleArgs a;
        memset(&a, 0, sizeof a);
        a.st = gs; a.logits = ctx->at<float>(ctx->o_pend_logits); a.latent = ctx->at<float>(ctx->o_pend_latent); a.seen = seen;
        a.V = V; a.Vpad = ctx->Vpad; a.D = D;
        a.top_k = sp->top_k; a.top_p = sp->top_p; a.top_p_threshold = sp->top_p_threshold; a.temperature = sp->temperature;
        a.rep_penalty = sp->repetition_penalty; a.ignore_eos = sp->ignore_eos; a.stop_token = g.stop_audio; a.max_total = max_total;
        a.seed = sp->seed;
        a.noise = exp_noise_dev ? exp_noise_dev + (size_t)i * B * V : nullptr;
        a.forced = forced_ids_dev ? reinterpret_cast<const long long*>(forced_ids_dev) + (size_t)i * B : nullptr;
        a.ids_out = reinterpret_cast<long long*>(ids_out_dev) + (size_t)i * B;
        a.latents_out = latents_out_dev + (size_t)i * B * D;
        a.logits_out = logits_out_dev ? logits_out_dev + (size_t)i * B * V : nullptr;
        a.status = status_dev; a.bad_ids = ctx->at<int>(ctx->o_flags); a.step_in_call = i;
        CK(launch_sample(a, B, st, &ctx->nlaunch));
        ctx->n_host = n + 1;
        ctx->pending = false;
    }
    return GENVC_OK;
}

// ---------------------------------------------------------------------------------------------
// teacher-forced latent pass
// ---------------------------------------------------------------------------------------------
int genvc_forward_latents(genvc_ctx* ctx, const float* cond_dev, const int64_t* text_ids_dev, int T, const int64_t* codes_dev,
                          int M, int B, float* latents_out_dev, void* stream) {
    if (int rc = check_ready(ctx)) return rc;
    DevGuard guard(ctx->device);
    CK(guard.err);
    const genvc_config& g = ctx->cfg;
    if (!cond_dev || !text_ids_dev || !codes_dev || !latents_out_dev) return ctx->fail(GENVC_E_INVALID, "null pointer");
    const int NLp = g.pc_latents, D = g.d_model;
    const int RT = T + 2, RM = M + 5, R = NLp + RT + RM;
    if (B <= 0 || M <= 0 || T < 0) return ctx->fail(GENVC_E_INVALID, "bad B/T/M");
    if (RT > g.n_text_pos || RM > g.n_mel_pos) return ctx->fail(GENVC_E_INVALID, "sequence exceeds the position tables");
    if ((size_t)B * R > ctx->rows_cap()) return ctx->fail(GENVC_E_INVALID, "B*rows = %d exceeds the workspace (%zu rows)", B * R, ctx->rows_cap());
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = ctx->layout;
    float* X = ctx->at<float>(ctx->o_X);
    // rows = [cond (32) ; text (T+2) ; start, codes, stop x4 (M+5)]
    CK(launch_embed_prefix(cond_dev, reinterpret_cast<const long long*>(text_ids_dev), B, T, NLp, D, ctx->w(L.text_emb),
                           ctx->w(L.text_pos), g.start_text, g.stop_text, X, (long)R * D, g.n_text_vocab, ctx->at<int>(ctx->o_flags) + 1, st, &ctx->nlaunch));
    CK(launch_embed_mel_rows(reinterpret_cast<const long long*>(codes_dev), B, RM, M, g.start_audio, g.stop_audio, 0, D,
                             ctx->w(L.mel_emb), ctx->w(L.mel_pos), X + (size_t)(NLp + RT) * D, (long)R * D, g.n_audio_vocab,
                             ctx->at<int>(ctx->o_flags) + 1, st, &ctx->nlaunch));
    if (int rc = run_blocks(ctx, B, R, -1, nullptr, st)) return rc;
    // final_norm(ln_f(h)) of the first M mel rows of each element
    CK(launch_layernorm(X + (size_t)(NLp + RT) * D, D, (long)R * D, latents_out_dev, D, (long)M * D, B * M, M, D, ctx->w(L.lnf_w),
                        ctx->w(L.lnf_b), ctx->w(L.fn_w), ctx->w(L.fn_b), nullptr, st, &ctx->nlaunch));
    // this pass reuses the row buffers but not the KV cache or the generation state
    return GENVC_OK;
}

// ---------------------------------------------------------------------------------------------
// KV-cache attention microbenchmark
// ---------------------------------------------------------------------------------------------
int genvc_kv_attention(const float* q_dev, const float* k_dev, const float* v_dev, int N, int H, int hd, int S, int S_max,
                       float* out_dev, void* stream) {
    if (!q_dev || !k_dev || !v_dev || !out_dev || N <= 0 || H <= 0 || !valid_hd(hd) || S <= 0 || S > S_max) return GENVC_E_INVALID;
    cudaError_t e = launch_kv_attention(q_dev, (long)H * hd, k_dev, v_dev, (long)H * S_max * hd, N, H, hd, S, S_max, out_dev,
                                        (long)H * hd, nullptr, (cudaStream_t)stream, nullptr);
    return e == cudaSuccess ? GENVC_OK : GENVC_E_CUDA;
}

}  // extern "C"
