// Host side of libttb200: weight store, workspace, forward orchestration and the C ABI
// (include/ttb200.h).  One engine = one model on one GPU; calls on an engine are not re-entrant.
#include "kernels.cuh"
#include "../../include/ttb200.h"

#include <sched.h>
#include <cmath>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace ttb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

void launch_f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t s);

int ensure_dyn_smem(const void* kernel, int bytes) {
    if (bytes <= 48 * 1024) return 0;
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> granted;
    int dev = 0;
    TTB_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int& have = granted[{dev, kernel}];
    if (bytes <= have) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_last_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize, ") + std::to_string(bytes) + ") failed on device " +
                       std::to_string(dev) + ": " + cudaGetErrorString(e));
        return 1;
    }
    have = bytes;
    return 0;
}

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}

bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("TTB_NO_PDL"); return !(v && v[0] == '1'); }();
    return on;
}


struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        TTB_CUDA_OK(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Param {
    int64_t numel = 0;
    float* dev = nullptr;
    __nv_bfloat16* devh = nullptr;  // bf16 copy (GEMM weights, precision = bf16)
    bool is_gemm_weight = false;
    bool set = false;
};

struct Lin {
    const float* w = nullptr;
    const float* b = nullptr;
    const __nv_bfloat16* wh = nullptr;
    int N = 0, K = 0;
    Lin rows(int r0, int n) const {
        Lin l = *this;
        l.w = w + (long long)r0 * K;
        l.wh = wh ? wh + (long long)r0 * K : nullptr;
        l.b = b ? b + r0 : nullptr;
        l.N = n;
        return l;
    }
};

struct Norm { const float* g = nullptr; const float* b = nullptr; };

// Kernel classes for the optional per-class CUDA-event timing (ttb_engine_set_profiling).
enum KC : int {
    KC_EMBED = 0, KC_GEMM_QKV, KC_SELF_ATTN, KC_GEMM_SELF_OUT, KC_LAYERNORM, KC_GEMM_CROSS_Q, KC_CROSS_ATTN,
    KC_GEMM_CROSS_OUT, KC_GEMM_FFN1, KC_GEMM_FFN2, KC_GEMM_CLASSIFIER, KC_ARGMAX, KC_ACCEPT, KC_CACHE_APPEND,
    KC_ENCODER, KC_MISC, KC_COUNT
};
static const char* kKcNames[KC_COUNT] = {
    "embed", "gemm_qkv", "self_attn", "gemm_self_out", "add_layernorm", "gemm_cross_q", "cross_attn",
    "gemm_cross_out", "gemm_ffn1", "gemm_ffn2", "gemm_classifier", "argmax", "accept", "cache_append",
    "encoder", "misc"};

struct Prof {
    uint32_t mask = 0;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int kc; size_t e0, e1; };
    std::vector<Rec> recs;
    double ms[KC_COUNT] = {};
    long long n[KC_COUNT] = {};
    int override_kc = -1;  // when >= 0 every launch is attributed to this class (encoder)
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t ev;
            cudaEventCreate(&ev);
            pool.push_back(ev);
        }
        return pool[used++];
    }
};

struct EncLayer { Lin in_proj, out_proj, ff1, ff2; Norm n1, n2; };
struct DecLayer { Lin self_in, self_out, cross_in, cross_out, ff1, ff2; Norm n1, n2, n3; };

}  // namespace ttb

using namespace ttb;

struct ttb_engine {
    ttb_model_desc d{};
    int device = 0;
    bool finalized = false;
    std::map<std::string, Param> params;
    std::vector<EncLayer> enc;
    std::vector<DecLayer> dec;
    Norm enc_norm, dec_norm;
    Lin classifier;
    const float* src_emb = nullptr;
    const float* tgt_emb = nullptr;
    const float* pe = nullptr;

    // forward workspace (typed by the precision's activation type at use)
    DevBuf x, xh, y, qkv, att, q2, hid, logits, tok32, keytok32, pred;
    // encoder products / decoding state
    DevBuf src32, srclen, desc, memory, memh, crosskv, crossvt, kcache, vcache, kcache2, vcache2, drafts, gen, front, active, ctrl, sel, out64;
    int* h_ctrl = nullptr;  // pinned snapshots of ctrl for lagged polling
    cudaEvent_t poll_ev[4]{};
    cudaEvent_t t0{}, t1{};
    long long launches = 0;
    Prof prof;
    // decoding loop: private stream + one captured CUDA graph per (shape, buffers) configuration
    cudaStream_t stream = nullptr;
    cudaEvent_t join_ev{};
    static constexpr int kBuckets = 10;
    cudaGraphExec_t graph_exec[kBuckets][2] = {};   // greedy iteration graphs per live-query bucket (grids sized for the bucket) x {long, short}
    long long graph_key[12] = {};
    long long graph_launches = 0;
    static constexpr int kBeamBuckets = 8;
    cudaGraphExec_t beam_graph[2][kBeamBuckets] = {};   // steady-state iteration of the speculative beam search, per ping-pong parity and live-candidate bucket
    long long beam_graph_key[20] = {};
    long long beam_graph_launches = 0;
    DevBuf beam;                 // arena of the beam-search state
    DevBuf hist;                 // per-iteration live-query count of the last generate()
    std::vector<int> h_hist;

    int E() const { return d.embedding_dim; }
    int HD() const { return d.embedding_dim / d.num_heads; }

    // Fingerprint of this engine's workspace (address and capacity of every buffer): part of the key of the captured
    // graph, so a buffer that moved invalidates the graph of THIS engine only (several engines decode concurrently from
    // different host threads, translation_transformer_b200/pipeline.py)
    long long alloc_signature() const {
        const DevBuf* bufs[] = {&x, &xh, &y, &qkv, &att, &q2, &hid, &logits, &tok32, &keytok32, &pred, &src32, &srclen, &desc, &memory, &memh,
                                &crosskv, &crossvt, &kcache, &vcache, &kcache2, &vcache2, &drafts, &gen, &front, &active, &ctrl, &sel, &out64, &beam, &hist};
        unsigned long long h = 1469598103934665603ull;
        for (const DevBuf* b : bufs) {
            h = (h ^ (unsigned long long)reinterpret_cast<uintptr_t>(b->p)) * 1099511628211ull;
            h = (h ^ (unsigned long long)b->cap) * 1099511628211ull;
        }
        return (long long)(h >> 1);
    }
};

namespace ttb {

static int add_param(ttb_engine* e, const std::string& name, int64_t numel, bool gemm_w) {
    Param p;
    p.numel = numel;
    p.is_gemm_weight = gemm_w;
    TTB_CUDA_OK(cudaMalloc(&p.dev, (size_t)numel * sizeof(float)));
    e->params[name] = p;
    return 0;
}

static int register_params(ttb_engine* e) {
    const ttb_model_desc& d = e->d;
    const int64_t E = d.embedding_dim, F = d.feedforward_dim;
    int rc = 0;
    rc |= add_param(e, "src_token_featurizer.embedding.weight", (int64_t)d.src_vocab_size * E, false);
    rc |= add_param(e, "tgt_token_featurizer.embedding.weight", (int64_t)d.tgt_vocab_size * E, false);
    rc |= add_param(e, "positional_encoding.pe", (int64_t)(d.max_positions + 1) * E, false);
    auto attn = [&](const std::string& p) {
        rc |= add_param(e, p + ".in_proj_weight", 3 * E * E, true);
        rc |= add_param(e, p + ".in_proj_bias", 3 * E, false);
        rc |= add_param(e, p + ".out_proj.weight", E * E, true);
        rc |= add_param(e, p + ".out_proj.bias", E, false);
    };
    auto ffn = [&](const std::string& p, int norms) {
        rc |= add_param(e, p + ".linear1.weight", F * E, true);
        rc |= add_param(e, p + ".linear1.bias", F, false);
        rc |= add_param(e, p + ".linear2.weight", E * F, true);
        rc |= add_param(e, p + ".linear2.bias", E, false);
        for (int j = 1; j <= norms; ++j) {
            rc |= add_param(e, p + ".norm" + std::to_string(j) + ".weight", E, false);
            rc |= add_param(e, p + ".norm" + std::to_string(j) + ".bias", E, false);
        }
    };
    for (int i = 0; i < d.num_encoder_layers; ++i) {
        std::string p = "transformer.encoder.layers." + std::to_string(i);
        attn(p + ".self_attn");
        ffn(p, 2);
    }
    rc |= add_param(e, "transformer.encoder.norm.weight", E, false);
    rc |= add_param(e, "transformer.encoder.norm.bias", E, false);
    for (int i = 0; i < d.num_decoder_layers; ++i) {
        std::string p = "transformer.decoder.layers." + std::to_string(i);
        attn(p + ".self_attn");
        attn(p + ".multihead_attn");
        ffn(p, 3);
    }
    rc |= add_param(e, "transformer.decoder.norm.weight", E, false);
    rc |= add_param(e, "transformer.decoder.norm.bias", E, false);
    rc |= add_param(e, "next_token_classifier.weight", (int64_t)d.tgt_vocab_size * E, true);
    rc |= add_param(e, "next_token_classifier.bias", d.tgt_vocab_size, false);
    return rc;
}

static Lin make_lin(ttb_engine* e, const std::string& w, const std::string& b, int N, int K) {
    Lin l;
    l.w = e->params[w].dev;
    l.wh = e->params[w].devh;
    l.b = e->params[b].dev;
    l.N = N;
    l.K = K;
    return l;
}
static Norm make_norm(ttb_engine* e, const std::string& p) {
    Norm n;
    n.g = e->params[p + ".weight"].dev;
    n.b = e->params[p + ".bias"].dev;
    return n;
}

// RAII scope around one kernel launch: counts it and, when the class is being profiled, brackets it
// with CUDA events on the launching stream.
struct Scope {
    ttb_engine* e;
    cudaStream_t s;
    int kc;
    bool on;
    size_t e0 = 0;
    Scope(ttb_engine* e_, int kc_, cudaStream_t s_) : e(e_), s(s_), kc(e_->prof.override_kc >= 0 ? e_->prof.override_kc : kc_) {
        e->launches++;
        on = (e->prof.mask >> kc) & 1u;
        if (on) {
            e0 = e->prof.used;
            cudaEventRecord(e->prof.get(), s);
        }
    }
    ~Scope() {
        if (on) {
            size_t e1 = e->prof.used;
            cudaEventRecord(e->prof.get(), s);
            e->prof.recs.push_back({kc, e0, e1});
        }
    }
};
static void prof_collect(ttb_engine* e) {  // call after the stream has been synchronised
    for (auto& r : e->prof.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e->prof.pool[r.e0], e->prof.pool[r.e1]) == cudaSuccess) {
            e->prof.ms[r.kc] += ms;
            e->prof.n[r.kc] += 1;
        }
    }
    e->prof.recs.clear();
    e->prof.used = 0;
}

// ---- typed forward helpers ---------------------------------------------------------------------
template <typename ActT> struct Prec;
template <> struct Prec<float> { static constexpr bool lowp = false; };
template <> struct Prec<__nv_bfloat16> { static constexpr bool lowp = true; };

template <typename OutT>
static int linear(ttb_engine* e, int kc, const float* A, int lda, const Lin& L, OutT* C, int ldc, RowCount rows, bool relu, cudaStream_t s) {
    Scope sc(e, kc, s);
    launch_gemm_f32<OutT>(A, lda, L.w, L.b, C, ldc, rows, L.N, L.K, relu, s);
    return 0;
}
template <typename OutT>
static int linear(ttb_engine* e, int kc, const __nv_bfloat16* A, int lda, const Lin& L, OutT* C, int ldc, RowCount rows, bool relu, cudaStream_t s) {
    Scope sc(e, kc, s);
    return launch_gemm_bf16_tc<OutT>(A, lda, L.wh, L.b, C, ldc, rows, L.N, L.K, relu, s);
}

// Sub-layer tail  x <- LN2?(LN1(x + A W^T + b)):  one fused tcgen05 kernel on the bf16 path with E = 256,
// GEMM (fp32 out) + add_layernorm otherwise.  TTB_NO_FUSED_LN=1 forces the unfused kernels (A/B comparisons).
static bool fused_ln_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TTB_NO_FUSED_LN"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
static int linear_resid_ln(ttb_engine* e, int kc, const float* A, int lda, const Lin& L, const Norm& n1, const Norm* n2,
                           float* x, float* /*xh*/, float* y, float* dst, float* dsth, RowCount rows, cudaStream_t s,
                           const Lin* = nullptr, float* = nullptr, bool* chained = nullptr) {
    if (chained) *chained = false;
    if (linear<float>(e, kc, A, lda, L, y, L.N, rows, false, s)) return 1;
    Scope sc(e, KC_LAYERNORM, s);
    launch_add_layernorm<float>(x, y, n1.g, n1.b, n2 ? n2->g : nullptr, n2 ? n2->b : nullptr, dst, dsth, rows, L.N, s);
    return 0;
}
// `chain` (optional): a projection of the normalised result that the same kernel computes on the way out
// (the cross-attention query); *chained tells the caller whether it was taken.
static bool chain_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TTB_NO_CHAIN"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
static int linear_resid_ln(ttb_engine* e, int kc, const __nv_bfloat16* A, int lda, const Lin& L, const Norm& n1, const Norm* n2,
                           float* x, __nv_bfloat16* xh, float* y, float* dst, __nv_bfloat16* dsth, RowCount rows, cudaStream_t s,
                           const Lin* chain = nullptr, __nv_bfloat16* q2 = nullptr, bool* chained = nullptr) {
    if (chained) *chained = false;
    if (L.N == 256 && dst == x && dsth == xh && fused_ln_enabled()) {
        Scope sc(e, kc, s);
        const bool do_chain = chain && q2 && L.K == 256 && chain->N == 256 && chain->K == 256 && chain_enabled();
        if (chained) *chained = do_chain;
        return launch_gemm_resid_ln(A, lda, L.wh, L.b, x, xh, n1.g, n1.b, n2 ? n2->g : nullptr, n2 ? n2->b : nullptr, rows, L.K, s,
                                    do_chain ? chain->wh : nullptr, do_chain ? chain->b : nullptr, do_chain ? q2 : nullptr);
    }
    if (linear<float>(e, kc, A, lda, L, y, L.N, rows, false, s)) return 1;
    Scope sc(e, KC_LAYERNORM, s);
    launch_add_layernorm<__nv_bfloat16>(x, y, n1.g, n1.b, n2 ? n2->g : nullptr, n2 ? n2->b : nullptr, dst, dsth, rows, L.N, s);
    return 0;
}

// Feed-forward sub-layer  x <- LN2?(LN1(x + relu(x W1^T + b1) W2^T + b2)):  one fused kernel (hidden activations stay
// on the SM) on the bf16 path with E = 256, FFN1 GEMM + (FFN2 GEMM + LN) otherwise.  TTB_NO_FUSED_FFN=1 disables it.
static bool fused_ffn_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TTB_NO_FUSED_FFN"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
static int ffn_block(ttb_engine* e, int kc1, int kc2, const Lin& L1, const Lin& L2, const Norm& n1, const Norm* n2, float* x, float* xh,
                     float* hid, float* y, float* dst, float* dsth, RowCount rows, cudaStream_t s) {
    if (linear<float>(e, kc1, x, L1.K, L1, hid, L1.N, rows, true, s)) return 1;
    return linear_resid_ln(e, kc2, hid, L1.N, L2, n1, n2, x, xh, y, dst, dsth, rows, s);
}
static int ffn_block(ttb_engine* e, int kc1, int kc2, const Lin& L1, const Lin& L2, const Norm& n1, const Norm* n2, float* x,
                     __nv_bfloat16* xh, __nv_bfloat16* hid, float* y, float* dst, __nv_bfloat16* dsth, RowCount rows, cudaStream_t s) {
    if (L2.N == 256 && L1.K == 256 && L1.N % 256 == 0 && L1.N <= 4096 && dst == x && dsth == xh && fused_ffn_enabled() && fused_ln_enabled()) {
        Scope sc(e, kc1, s);
        return launch_ffn_fused(xh, L1.wh, L1.b, L2.wh, L2.b, x, n1.g, n1.b, n2 ? n2->g : nullptr, n2 ? n2->b : nullptr, rows, L1.N, s);
    }
    if (linear<__nv_bfloat16>(e, kc1, xh, L1.K, L1, hid, L1.N, rows, true, s)) return 1;
    return linear_resid_ln(e, kc2, hid, L1.N, L2, n1, n2, x, xh, y, dst, dsth, rows, s);
}

// Out-projection + LayerNorm of the preceding sub-layer chained into the fused feed-forward launch (bf16, E = 256, CTA-pair
// kernel available).  Returns 0 when taken, -1 when not applicable (caller launches the two kernels), > 0 on error.
static int ffn_block_chained(ttb_engine*, int, const float*, const Lin&, const Norm&, const Lin&, const Lin&, const Norm&, const Norm*, float*, float*,
                             RowCount, cudaStream_t) {
    return -1;
}
static int ffn_block_chained(ttb_engine* e, int kc, const __nv_bfloat16* att, const Lin& Lo, const Norm& n0, const Lin& L1, const Lin& L2,
                             const Norm& n1, const Norm* n2, float* x, __nv_bfloat16* xh, RowCount rows, cudaStream_t s) {
    if (!(Lo.N == 256 && Lo.K == 256 && L2.N == 256 && L1.K == 256 && L1.N % 256 == 0 && L1.N <= 4096 && fused_ffn_enabled() && fused_ln_enabled() &&
          chain_enabled() && ffn_pair_available(rows.max_rows)))
        return -1;
    static const bool off = [] { const char* v = getenv("TTB_NO_FFN_CHAIN"); return v && v[0] == '1'; }();
    if (off) return -1;
    Scope sc(e, kc, s);
    return launch_ffn_fused(xh, L1.wh, L1.b, L2.wh, L2.b, x, n1.g, n1.b, n2 ? n2->g : nullptr, n2 ? n2->b : nullptr, rows, L1.N, s, att, Lo.wh, Lo.b,
                            n0.g, n0.b) ? 1 : 0;
}

// Attention dispatch: fp32 path -> SIMT kernels (exact), bf16 path -> tensor-core kernels
// (TTB_ATTN_SIMT=1 forces the SIMT kernels for A/B comparisons).
static bool attn_simt_forced() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TTB_ATTN_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
static void attn(const float* q, int q_ld, const float* k, const float* v, int kv_ld, float* out, int out_ld, int ng, const int* ng_dev,
                 int Lq, int Lk, long long kvs, const int* kvmap, const int* key_tok, int kts, int pad, bool causal, int H, int HD, cudaStream_t s,
                 const int* lk_dev = nullptr, const int* = nullptr, const int4* = nullptr) {
    launch_attention<float>(q, q_ld, k, v, kv_ld, out, out_ld, ng, ng_dev, Lq, Lk, kvs, kvmap, key_tok, kts, pad, causal, H, HD, s, lk_dev);
}
static void attn(const __nv_bfloat16* q, int q_ld, const __nv_bfloat16* k, const __nv_bfloat16* v, int kv_ld, __nv_bfloat16* out, int out_ld,
                 int ng, const int* ng_dev, int Lq, int Lk, long long kvs, const int* kvmap, const int* key_tok, int kts, int pad,
                 bool causal, int H, int HD, cudaStream_t s, const int* lk_dev = nullptr, const int* lk_group = nullptr,
                 const int4* desc = nullptr) {
    if (attn_simt_forced())
        launch_attention<__nv_bfloat16>(q, q_ld, k, v, kv_ld, out, out_ld, ng, ng_dev, Lq, Lk, kvs, kvmap, key_tok, kts, pad, causal, H, HD, s, lk_dev);
    else
        launch_attention_mma(q, q_ld, k, v, kv_ld, out, out_ld, ng, ng_dev, Lq, Lk, kvs, kvmap, key_tok, kts, pad, causal, H, HD, s, lk_dev, lk_group, desc);
}
static void spec_attn(const float* qkv, int ld, const float* kc, const float* vc, long long cqs, int cld, float* out, int old, int B,
                      const int* na, const int* active, const int* front, const int* gen, int gen_ld, int pad, int N, int D, int H,
                      int HD, int P, cudaStream_t s, const int4* = nullptr) {
    launch_spec_self_attention<float>(qkv, ld, kc, vc, cqs, cld, out, old, B, na, active, front, gen, gen_ld, pad, N, D, H, HD, P, s);
}
static void spec_attn(const __nv_bfloat16* qkv, int ld, const __nv_bfloat16* kc, const __nv_bfloat16* vc, long long cqs, int cld,
                      __nv_bfloat16* out, int old, int B, const int* na, const int* active, const int* front, const int* gen,
                      int gen_ld, int pad, int N, int D, int H, int HD, int P, cudaStream_t s, const int4* desc = nullptr) {
    if (attn_simt_forced())
        launch_spec_self_attention<__nv_bfloat16>(qkv, ld, kc, vc, cqs, cld, out, old, B, na, active, front, gen, gen_ld, pad, N, D, H, HD, P, s);
    else
        launch_spec_self_attention_mma(qkv, ld, kc, vc, cqs, cld, out, old, B, na, active, front, gen, gen_ld, pad, N, D, H, HD, s, desc);
}

// GEMM A-operand view of the residual stream: fp32 path reads x itself, bf16 path its bf16 copy
template <typename ActT> static const ActT* a_view(const float* x, const ActT* xh);
template <> const float* a_view<float>(const float* x, const float*) { return x; }
template <> const __nv_bfloat16* a_view<__nv_bfloat16>(const float*, const __nv_bfloat16* xh) { return xh; }

template <typename ActT>
static int ensure_work(ttb_engine* e, long long T, int qkv_layers) {
    const long long E = e->E(), F = e->d.feedforward_dim, V = e->d.tgt_vocab_size;
    int rc = 0;
    rc |= e->x.ensure(T * E * sizeof(float));
    if (Prec<ActT>::lowp) rc |= e->xh.ensure(T * E * sizeof(ActT));
    rc |= e->y.ensure(T * E * sizeof(float));
    rc |= e->qkv.ensure(T * 3 * E * sizeof(ActT) * qkv_layers);
    rc |= e->att.ensure(T * E * sizeof(ActT));
    rc |= e->q2.ensure(T * E * sizeof(ActT));
    rc |= e->hid.ensure(T * F * sizeof(ActT));
    rc |= e->logits.ensure(T * V * sizeof(float));
    rc |= e->tok32.ensure(T * sizeof(int));
    rc |= e->pred.ensure(T * sizeof(int));
    return rc;
}

// Encoder stack on tokens src32 (B, Ls); writes memory (fp32) and, on the bf16 path, memh.
template <typename ActT>
static int encode_impl(ttb_engine* e, const int* src32, const int* key_tok, int B, int Ls, float* mem_out, ActT* memh_out, cudaStream_t s) {
    const int E = e->E(), T = B * Ls, H = e->d.num_heads, HD = e->HD();
    if (ensure_work<ActT>(e, T, 1)) return 1;
    float* x = e->x.as<float>();
    ActT* xh = Prec<ActT>::lowp ? e->xh.as<ActT>() : nullptr;
    float* y = e->y.as<float>();
    ActT* qkv = e->qkv.as<ActT>();
    ActT* att = e->att.as<ActT>();
    ActT* hid = e->hid.as<ActT>();
    RowCount rows(T);
    e->prof.override_kc = KC_ENCODER;
    struct Reset { ttb_engine* e; ~Reset() { e->prof.override_kc = -1; } } reset{e};
    { Scope sc(e, KC_ENCODER, s); launch_embed_seq<ActT>(src32, T, Ls, e->src_emb, e->pe, E, x, xh, s); }
    const int n_layers = (int)e->enc.size();
    for (int l = 0; l < n_layers; ++l) {
        const EncLayer& L = e->enc[l];
        const bool last = l + 1 == n_layers;
        if (linear<ActT>(e, KC_ENCODER, a_view<ActT>(x, xh), E, L.in_proj, qkv, 3 * E, rows, false, s)) return 1;
        {
            Scope sc(e, KC_ENCODER, s);
            attn(qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, B, nullptr, Ls, Ls, Ls, nullptr,
                                   key_tok, Ls, e->d.src_pad_token_idx, false, H, HD, s);
        }
        if (linear_resid_ln(e, KC_ENCODER, att, E, L.out_proj, L.n1, nullptr, x, xh, y, x, xh, rows, s)) return 1;
        float* dst = last ? mem_out : x;
        ActT* dsth = last ? memh_out : xh;
        if (ffn_block(e, KC_ENCODER, KC_ENCODER, L.ff1, L.ff2, L.n2, last ? &e->enc_norm : nullptr, x, xh, hid, y, dst, dsth, rows, s)) return 1;
    }
    return 0;
}

// Cross-attention K/V of every decoder layer from the encoder memory: crosskv[l] = (B*Ls, 2E)
template <typename ActT>
static int cross_kv_impl(ttb_engine* e, const float* mem, const ActT* memh, int rows_n, ActT* crosskv, cudaStream_t s,
                         long long layer_stride = 0) {
    const int E = e->E();
    RowCount rows(rows_n);
    if (layer_stride == 0) layer_stride = (long long)rows_n * 2 * E;
    for (size_t l = 0; l < e->dec.size(); ++l) {
        Lin kv = e->dec[l].cross_in.rows(E, 2 * E);
        if (linear<ActT>(e, KC_ENCODER, a_view<ActT>(mem, memh), E, kv, crosskv + (long long)l * layer_stride, 2 * E, rows, false, s)) return 1;
    }
    return 0;
}

// Decoder layers on the token matrix in e->x / e->xh (`rows` live rows).  `self_attn` / `cross_attn`
// enqueue the layer's attention kernels given the freshly projected q/k/v of that layer.
template <typename ActT, typename SelfAttn, typename CrossAttn>
static int decoder_stack(ttb_engine* e, RowCount rows, int qkv_layers, long long qkv_layer_stride,
                         SelfAttn self_attn, CrossAttn cross_attn, cudaStream_t s) {
    const int E = e->E();
    float* x = e->x.as<float>();
    ActT* xh = Prec<ActT>::lowp ? e->xh.as<ActT>() : nullptr;
    float* y = e->y.as<float>();
    ActT* att = e->att.as<ActT>();
    ActT* q2 = e->q2.as<ActT>();
    ActT* hid = e->hid.as<ActT>();
    const int n_layers = (int)e->dec.size();
    for (int l = 0; l < n_layers; ++l) {
        const DecLayer& L = e->dec[l];
        const bool last = l + 1 == n_layers;
        ActT* qkv = e->qkv.as<ActT>() + (qkv_layers > 1 ? (long long)l * qkv_layer_stride : 0);
        if (linear<ActT>(e, KC_GEMM_QKV, a_view<ActT>(x, xh), E, L.self_in, qkv, 3 * E, rows, false, s)) return 1;
        { Scope sc(e, KC_SELF_ATTN, s); self_attn(l, qkv, att); }
        const Lin cross_q = L.cross_in.rows(0, E);
        bool chained = false;
        if (linear_resid_ln(e, KC_GEMM_SELF_OUT, att, E, L.self_out, L.n1, nullptr, x, xh, y, x, xh, rows, s, &cross_q, q2, &chained)) return 1;
        if (!chained && linear<ActT>(e, KC_GEMM_CROSS_Q, a_view<ActT>(x, xh), E, cross_q, q2, E, rows, false, s)) return 1;
        { Scope sc(e, KC_CROSS_ATTN, s); cross_attn(l, q2, att); }
        // cross-attention out-projection + LayerNorm2 + feed-forward block + LayerNorm3 in one launch where the CTA-pair
        // feed-forward kernel runs, two launches otherwise
        int fused = ffn_block_chained(e, KC_GEMM_FFN1, att, L.cross_out, L.n2, L.ff1, L.ff2, L.n3, last ? &e->dec_norm : nullptr, x, xh, rows, s);
        if (fused > 0) return 1;
        if (fused == 0) continue;
        if (linear_resid_ln(e, KC_GEMM_CROSS_OUT, att, E, L.cross_out, L.n2, nullptr, x, xh, y, x, xh, rows, s)) return 1;
        if (ffn_block(e, KC_GEMM_FFN1, KC_GEMM_FFN2, L.ff1, L.ff2, L.n3, last ? &e->dec_norm : nullptr, x, xh, hid, y, x, xh, rows, s)) return 1;
    }
    return 0;
}

template <typename ActT>
static int encode_api(ttb_engine* e, const int64_t* src_dev, const uint8_t* mask_dev, int B, int Ls, float* mem_out, cudaStream_t s) {
    const long long T = (long long)B * Ls;
    if (e->src32.ensure(T * sizeof(int))) return 1;
    int* src32 = e->src32.as<int>();
    { Scope sc(e, KC_MISC, s); launch_i64_to_i32(reinterpret_cast<const long long*>(src_dev), src32, T, s); }
    const int* key_src = src32;
    if (mask_dev) {  // explicit mask: express it as a token-like array for the attention kernels
        if (e->keytok32.ensure(T * sizeof(int))) return 1;
        Scope sc(e, KC_MISC, s);
        launch_mask_to_tokens(mask_dev, e->keytok32.as<int>(), T, e->d.src_pad_token_idx, s);
        key_src = e->keytok32.as<int>();
    }
    if (Prec<ActT>::lowp && e->memh.ensure(T * e->E() * sizeof(ActT))) return 1;
    int rc = encode_impl<ActT>(e, src32, key_src, B, Ls, mem_out, Prec<ActT>::lowp ? e->memh.as<ActT>() : nullptr, s);
    if (rc) return rc;
    TTB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename ActT>
static int decode_api(ttb_engine* e, const int64_t* tgt_dev, int B, int Lt, const float* memory_dev,
                      const uint8_t* mem_mask_dev, int Ls, float* logits_out, cudaStream_t s) {
    const int E = e->E(), H = e->d.num_heads, HD = e->HD(), V = e->d.tgt_vocab_size;
    const long long T = (long long)B * Lt, TS = (long long)B * Ls;
    if (ensure_work<ActT>(e, T, 1)) return 1;
    if (e->keytok32.ensure(TS * sizeof(int))) return 1;
    if (e->crosskv.ensure(TS * 2 * E * sizeof(ActT) * e->dec.size())) return 1;
    int* tgt32 = e->tok32.as<int>();
    int* memtok = e->keytok32.as<int>();
    { Scope sc(e, KC_MISC, s); launch_i64_to_i32(reinterpret_cast<const long long*>(tgt_dev), tgt32, T, s); }
    { Scope sc(e, KC_MISC, s); launch_mask_to_tokens(mem_mask_dev, memtok, TS, e->d.src_pad_token_idx, s); }
    const ActT* memh = nullptr;
    if (Prec<ActT>::lowp) {
        if (e->memh.ensure(TS * E * sizeof(ActT))) return 1;
        Scope sc(e, KC_MISC, s);
        launch_f32_to_bf16(memory_dev, reinterpret_cast<__nv_bfloat16*>(e->memh.p), TS * E, s);
        memh = e->memh.as<ActT>();
    }
    ActT* crosskv = e->crosskv.as<ActT>();
    if (cross_kv_impl<ActT>(e, memory_dev, memh, (int)TS, crosskv, s)) return 1;
    float* x = e->x.as<float>();
    ActT* xh = Prec<ActT>::lowp ? e->xh.as<ActT>() : nullptr;
    { Scope sc(e, KC_EMBED, s); launch_embed_seq<ActT>(tgt32, (int)T, Lt, e->tgt_emb, e->pe, E, x, xh, s); }
    RowCount rows((int)T);
    auto self_attn = [&](int, ActT* qkv, ActT* att) {
        attn(qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, B, nullptr, Lt, Lt, Lt, nullptr,
                               tgt32, Lt, e->d.tgt_pad_token_idx, true, H, HD, s);
    };
    auto cross_attn = [&](int l, ActT* q2, ActT* att) {
        const ActT* kv = crosskv + (long long)l * TS * 2 * E;
        attn(q2, E, kv, kv + E, 2 * E, att, E, B, nullptr, Lt, Ls, Ls, nullptr,
                               memtok, Ls, e->d.src_pad_token_idx, false, H, HD, s);
    };
    if (decoder_stack<ActT>(e, rows, 1, 0, self_attn, cross_attn, s)) return 1;
    if (linear<float>(e, KC_GEMM_CLASSIFIER, a_view<ActT>(x, xh), E, e->classifier, logits_out, V, rows, false, s)) return 1;
    TTB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename ActT>
static int greedy_api(ttb_engine* e, const int64_t* src_dev, int B, int Ls, int max_len, int draft_len, int N,
                      int pad, int bos, int eos, int replace, int tie_break, int64_t* out_dev, int32_t* trace_dev,
                      ttb_generate_stats* stats, cudaStream_t user_stream, bool standard = false) {
    const int E = e->E(), H = e->d.num_heads, HD = e->HD(), V = e->d.tgt_vocab_size;
    const int n_dec = (int)e->dec.size();
    // the loop runs on the engine's own stream (graph capture is not allowed on the legacy default
    // stream); it is ordered after the caller's stream and fully synchronised before returning
    cudaStream_t s = e->stream;
    TTB_CUDA_OK(cudaEventRecord(e->join_ev, user_stream));
    TTB_CUDA_OK(cudaStreamWaitEvent(s, e->join_ev, 0));
    // speculative: make_drafts(min_draft_len=1, max_draft_len=max_len); standard greedy decoding (standard_decoding.py)
    // is the same KV-cached loop with one row per query and no draft tokens
    const int D = standard ? 0 : std::min(std::max(1, draft_len), max_len);
    if (standard) N = 1;
    const long long TS = (long long)B * Ls;
    // buffers that depend on the source length are sized for a rounded-up capacity so that batches of
    // different length neither reallocate nor invalidate the captured graph
    const int Ls_cap = std::max(256, (Ls + 63) / 64 * 64);
    const long long TS_cap = (long long)B * Ls_cap;
    const int per_q = N * (D + 1);
    const long long T = (long long)B * per_q;
    const int gen_ld = max_len + D + 2;
    const int P = gen_ld;  // cache positions per query
    const long long launches0 = e->launches;
    const int max_iters = max_len + 1;

    // encoder + cross-attention K/V (computed once per query, not once per draft row and iteration)
    if (e->src32.ensure(TS_cap * sizeof(int)) || e->memory.ensure(TS_cap * E * sizeof(float)) || e->srclen.ensure((size_t)B * sizeof(int))) return 1;
    if (Prec<ActT>::lowp && e->memh.ensure(TS_cap * E * sizeof(ActT))) return 1;
    if (e->crosskv.ensure(TS_cap * 2 * E * sizeof(ActT) * n_dec)) return 1;
    // tcgen05 attention inside the loop (bf16, head_dim 32, key counts within its tensor-memory budget): the value caches
    // are kept transposed ([dim][position], pitch VT_PITCH) so that both products read K-major operands
    constexpr int VT_PITCH = 256;
    const bool tc_attn = Prec<ActT>::lowp && !attn_simt_forced() && attention_tc_supported(HD, per_q, std::max(P, Ls), D + 1);
    const size_t vcache_bytes = tc_attn ? (size_t)n_dec * B * E * VT_PITCH * sizeof(ActT) : (size_t)n_dec * B * P * E * sizeof(ActT);
    if (e->kcache.ensure((size_t)n_dec * B * P * E * sizeof(ActT)) || e->vcache.ensure(vcache_bytes)) return 1;
    if (tc_attn && e->crossvt.ensure((size_t)n_dec * B * E * VT_PITCH * sizeof(ActT))) return 1;
    if (e->drafts.ensure((size_t)B * N * D * sizeof(int)) || e->gen.ensure((size_t)B * gen_ld * sizeof(int))) return 1;
    if (e->front.ensure(B * sizeof(int)) || e->active.ensure(B * sizeof(int)) || e->ctrl.ensure(CTRL_COUNT * sizeof(int))) return 1;
    if (e->sel.ensure((size_t)B * 4 * sizeof(int)) || e->out64.ensure((size_t)B * max_len * sizeof(long long))) return 1;
    if (e->desc.ensure((size_t)B * sizeof(int4))) return 1;
    if (e->hist.ensure((size_t)(max_iters + 1) * sizeof(int))) return 1;
    if (ensure_work<ActT>(e, std::max(T, TS_cap), n_dec)) return 1;

    TTB_CUDA_OK(cudaEventRecord(e->t0, s));
    int* src32 = e->src32.as<int>();
    { Scope sc(e, KC_MISC, s); launch_i64_to_i32(reinterpret_cast<const long long*>(src_dev), src32, TS, s); }
    // per-query source length: the cross-attention kernels skip the padding behind it
    { Scope sc(e, KC_MISC, s); launch_row_lengths(src32, B, Ls, e->d.src_pad_token_idx, e->srclen.as<int>(), s); }
    float* mem = e->memory.as<float>();
    ActT* memh = Prec<ActT>::lowp ? e->memh.as<ActT>() : nullptr;
    if (encode_impl<ActT>(e, src32, src32, B, Ls, mem, memh, s)) return 1;
    ActT* crosskv = e->crosskv.as<ActT>();
    if (cross_kv_impl<ActT>(e, mem, memh, (int)TS, crosskv, s, TS_cap * 2 * E)) return 1;
    const long long vt_q_stride = (long long)E * VT_PITCH, vt_l_stride = (long long)B * E * VT_PITCH;
    if constexpr (Prec<ActT>::lowp) {
        if (tc_attn) {
            // every position a product can touch must hold a finite number (masked keys get probability 0, and 0 x NaN = NaN)
            TTB_CUDA_OK(cudaMemsetAsync(e->vcache.p, 0, vcache_bytes, s));
            TTB_CUDA_OK(cudaMemsetAsync(e->crossvt.p, 0, (size_t)n_dec * B * E * VT_PITCH * sizeof(ActT), s));
            for (int l = 0; l < n_dec; ++l) {
                Scope sc(e, KC_ENCODER, s);
                launch_transpose_v(crosskv + (long long)l * TS_cap * 2 * E + E, 2 * E, B, Ls, E, e->crossvt.as<ActT>() + (long long)l * vt_l_stride, VT_PITCH, s);
            }
        }
    }
    // drafts from the source without its BOS column (speculative_decoding.py:64-73)
    if (!standard) { Scope sc(e, KC_MISC, s); launch_make_drafts(src32 + 1, Ls, B, Ls - 1, D, N, eos, pad, replace, e->drafts.as<int>(), s); }

    GreedyState st{};
    st.B = B; st.N = N; st.D = D; st.max_len = max_len; st.gen_ld = gen_ld; st.pad = pad; st.bos = bos; st.eos = eos; st.Ls = Ls;
    st.gen = e->gen.as<int>(); st.front = e->front.as<int>(); st.active = e->active.as<int>(); st.ctrl = e->ctrl.as<int>();
    st.drafts = e->drafts.as<int>(); st.pred = e->pred.as<int>(); st.out = e->out64.as<long long>();
    st.sel = e->sel.as<int>(); st.trace = trace_dev; st.tie_break = tie_break; st.hist = e->hist.as<int>();
    st.desc = e->desc.as<int4>(); st.src_len = e->srclen.as<int>();
    { Scope sc(e, KC_MISC, s); launch_greedy_init(st, s); }

    float* x = e->x.as<float>();
    ActT* xh = Prec<ActT>::lowp ? e->xh.as<ActT>() : nullptr;
    ActT* kc = e->kcache.as<ActT>();
    ActT* vc = e->vcache.as<ActT>();
    const long long cache_q_stride = (long long)P * E, cache_l_stride = (long long)B * P * E;
    const long long qkv_l_stride = T * 3 * E;
    const int* n_active = st.ctrl + CTRL_N_ACTIVE;
    // Grids of an iteration are sized for `Bq` queries: the whole batch, or -- once the host has seen queries retire -- one of
    // a few smaller buckets (live-query count only shrinks, so the count the host read last is an upper bound for every
    // iteration it launches afterwards).  CTAs beyond the live range exit at once either way, but each of them still has to
    // find an SM with ~200 KB of shared memory free just to do so, which serialises the kernels of the OTHER batches in
    // flight behind it; with batches that thin out (trained-like weights) most CTAs of a full-size grid are such no-ops.
    int Bq = B;

    auto self_attn = [&](int l, ActT* qkv, ActT* att) {
        if constexpr (Prec<ActT>::lowp) {
            if (tc_attn) {
                launch_spec_self_attention_tc(qkv, 3 * E, kc + l * cache_l_stride, cache_q_stride, E, vc + l * vt_l_stride, vt_q_stride, VT_PITCH, att, E,
                                              Bq, n_active, st.gen, gen_ld, e->d.tgt_pad_token_idx, N, D, H, st.desc, s);
                return;
            }
        }
        spec_attn(qkv, 3 * E, kc + l * cache_l_stride, vc + l * cache_l_stride, cache_q_stride, E,
                                         att, E, Bq, n_active, st.active, st.front, st.gen, gen_ld, e->d.tgt_pad_token_idx,
                                         N, D, H, HD, P, s, st.desc);
    };
    auto cross_attn = [&](int l, ActT* q2, ActT* att) {
        const ActT* kv = crosskv + (long long)l * TS_cap * 2 * E;
        if constexpr (Prec<ActT>::lowp) {
            if (tc_attn) {
                // the K/V rows of a query start at row query * Ls of the layer's block (the batch's own source length, read on the
                // device by the kernel: the row stride of a group is Ls * 2E elements)
                launch_cross_attention_tc(q2, E, kv, 2 * E, (long long)Ls * 2 * E, e->crossvt.as<ActT>() + (long long)l * vt_l_stride, vt_q_stride, VT_PITCH,
                                          att, E, Bq, n_active, per_q, src32, Ls, e->d.src_pad_token_idx, st.ctrl + CTRL_LS, H, st.desc, s);
                return;
            }
        }
        attn(q2, E, kv, kv + E, 2 * E, att, E, Bq, n_active, per_q, Ls, Ls, st.active,
             src32, Ls, e->d.src_pad_token_idx, false, H, HD, s, st.ctrl + CTRL_LS, e->srclen.as<int>(), st.desc);
    };

    auto enqueue_iteration = [&]() -> int {
        RowCount rows(Bq * per_q, n_active, per_q);
        {   // KV-cache append of the previous iteration's accepted tokens + embedding of this iteration's step tokens
            Scope sc(e, KC_EMBED, s);
            GreedyState stq = st;
            stq.B = Bq;               // grid sizing only: the kernel indexes the live slots, all below the bucket
            launch_greedy_advance<ActT>(stq, e->tgt_emb, e->pe, E, x, xh, e->qkv.as<ActT>(), qkv_l_stride, n_dec, 3 * E, kc, vc,
                                        cache_l_stride, cache_q_stride, E, s, tc_attn ? vt_l_stride : 0, tc_attn ? vt_q_stride : 0, tc_attn ? VT_PITCH : 0);
        }
        if (decoder_stack<ActT>(e, rows, n_dec, qkv_l_stride, self_attn, cross_attn, s)) return 1;
        bool fused_cls = false;
        if constexpr (Prec<ActT>::lowp) {
            static const bool off = [] { const char* v = getenv("TTB_NO_FUSED_ARGMAX"); return v && v[0] == '1'; }();
            if (!off) {
                Scope sc(e, KC_GEMM_CLASSIFIER, s);
                const int rc = launch_classifier_argmax(xh, E, e->classifier.wh, e->classifier.b, st.pred, rows, V, E, s);
                if (rc > 0) return 1;
                fused_cls = rc == 0;
                if (!fused_cls) e->launches--;
            }
        }
        if (!fused_cls) {
            if (linear<float>(e, KC_GEMM_CLASSIFIER, a_view<ActT>(x, xh), E, e->classifier, e->logits.as<float>(), V, rows, false, s)) return 1;
            { Scope sc(e, KC_ARGMAX, s); launch_argmax_rows(e->logits.as<float>(), V, V, st.pred, rows, s); }
        }
        { Scope sc(e, KC_ACCEPT, s); if (standard) launch_greedy_std_step(st, s); else launch_greedy_accept(st, s); }
        return 0;
    };

    // One decoding iteration is a fixed kernel sequence on fixed buffers (all sizes that change while
    // decoding live in device memory), so it is captured once into a CUDA graph and replayed.
    static const bool no_graph = [] { const char* v = getenv("TTB_NO_GRAPH"); return v && v[0] == '1'; }();
    const bool use_graph = !no_graph && !trace_dev && e->prof.mask == 0;
    // Several iterations per graph: back-to-back graph launches of one stream leave the GPU idle for ~12 us (launch latency
    // plus the control-word copy in between), a graph edge costs ~2 us.  Iterations behind the end of the loop are no-ops
    // (every kernel reads a live count of zero), so a longer graph only adds a few of those per batch.
    static const int graph_iters = [] { const char* v = getenv("TTB_GRAPH_ITERS"); const int k = v ? atoi(v) : 4; return k < 1 ? 1 : (k > 16 ? 16 : k); }();
    const int K_it = use_graph ? graph_iters : 1;
    // buckets of live queries a graph is captured for (largest first; lazily, the first time the host picks one)
    int bucket_q[ttb_engine::kBuckets];
    int n_buckets = 0;
    {
        static const bool no_buckets = [] { const char* v = getenv("TTB_NO_BUCKETS"); return v && v[0] == '1'; }();
        int cand[ttb_engine::kBuckets];
        for (int i = 0; i < 8; ++i) cand[i] = (B * (8 - i) + 7) / 8;      // B, 7B/8, ... B/8
        cand[8] = (B + 15) / 16;
        cand[9] = 1;
        for (int i = 0; i < ttb_engine::kBuckets; ++i) {
            if (i > 0 && (no_buckets || standard)) break;
            if (n_buckets == 0 || (cand[i] >= 1 && cand[i] < bucket_q[n_buckets - 1])) bucket_q[n_buckets++] = cand[i];
        }
    }
    if (use_graph) {
        const long long key[12] = {B, N, D, standard ? 1 : 0, max_len, pad, bos, eos, tie_break, e->alloc_signature(),
                                   (long long)sizeof(ActT) + 16 * K_it + (tc_attn ? 1024 : 0), replace};
        if (memcmp(key, e->graph_key, sizeof(key)) != 0) {
            for (auto& gb : e->graph_exec) for (auto& g : gb) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
            memcpy(e->graph_key, key, sizeof(key));
        }
    }
    // Once queries retire the graphs get SHORT (half the iterations): the host then follows the shrinking live count more
    // closely (smaller buckets sooner) and fewer no-op iterations run behind the end of the batch; while the whole batch is
    // alive (always, with random-init weights) the long graph saves the ~12 us boundary between launches.
    const int K_short = std::max(1, K_it / 2);
    auto graph_for = [&](int bi, int sh) -> int {      // capture on first use
        if (e->graph_exec[bi][sh]) return 0;
        const long long l0 = e->launches;
        cudaGraph_t graph = nullptr;
        Bq = bucket_q[bi];
        const int n_it = sh ? K_short : K_it;
        TTB_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        int rc = 0;
        for (int k = 0; k < n_it && !rc; ++k) rc = enqueue_iteration();
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (!rc) set_last_error(std::string("CUDA graph capture failed: ") + cudaGetErrorString(ce));
            return 1;
        }
        TTB_CUDA_OK(cudaGraphInstantiate(&e->graph_exec[bi][sh], graph, 0));
        cudaGraphDestroy(graph);
        e->graph_launches = (e->launches - l0) / n_it;   // per iteration
        e->launches = l0;  // capturing did not launch anything
        return 0;
    };

    // Lagged polling: the host looks at the control words of iteration (it - LAG) while iterations up
    // to `it` are already queued, so the GPU never waits for it.
    constexpr int RING = 4;
    const int LAG = K_it > 1 ? 1 : 2;   // launches (of K_it iterations each) the host stays ahead of the control words it reads
    int it = 0, iters_enq = 0;
    bool done = false;
    int known_live = B;                 // live queries in the newest control words read: an upper bound from then on
    while (!done && iters_enq < max_iters) {
        int bi = 0;
        while (bi + 1 < n_buckets && bucket_q[bi + 1] >= known_live) ++bi;
        Bq = bucket_q[bi];
        if (use_graph) {
            const int sh = (known_live < B && K_short < K_it) ? 1 : 0;
            if (graph_for(bi, sh)) return 1;
            TTB_CUDA_OK(cudaGraphLaunch(e->graph_exec[bi][sh], s));
            e->launches += e->graph_launches * (sh ? K_short : K_it);
            iters_enq += sh ? K_short : K_it;
        } else if (enqueue_iteration()) {
            return 1;
        } else {
            ++iters_enq;
        }
        TTB_CUDA_OK(cudaMemcpyAsync(e->h_ctrl + (it % RING) * CTRL_COUNT, st.ctrl, CTRL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
        TTB_CUDA_OK(cudaEventRecord(e->poll_ev[it % RING], s));
        if (it >= LAG) {
            const int j = (it - LAG) % RING;
            TTB_CUDA_OK(cudaEventSynchronize(e->poll_ev[j]));
            if (e->h_ctrl[j * CTRL_COUNT + CTRL_DONE]) done = true;
            else known_live = std::min(known_live, std::max(1, e->h_ctrl[j * CTRL_COUNT + CTRL_N_ACTIVE]));
        }
        ++it;
    }
    TTB_CUDA_OK(cudaMemcpyAsync(out_dev, st.out, (size_t)B * max_len * sizeof(long long), cudaMemcpyDeviceToDevice, s));
    TTB_CUDA_OK(cudaMemcpyAsync(e->h_ctrl, st.ctrl, CTRL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
    TTB_CUDA_OK(cudaEventRecord(e->t1, s));
    TTB_CUDA_OK(cudaStreamSynchronize(s));
    TTB_CUDA_OK(cudaGetLastError());
    prof_collect(e);
    const int* c = e->h_ctrl;
    TTB_CHECK(c[CTRL_DONE] == 1, "greedy decoding loop did not terminate within max_len + 1 iterations");
    e->h_hist.assign((size_t)c[CTRL_ITERS], 0);
    if (c[CTRL_ITERS] > 0)
        TTB_CUDA_OK(cudaMemcpy(e->h_hist.data(), st.hist, (size_t)c[CTRL_ITERS] * sizeof(int), cudaMemcpyDeviceToHost));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->t0, e->t1);
    if (stats) {
        stats->model_calls = c[CTRL_ITERS];
        stats->accepted_tokens = c[CTRL_ACCEPTED];
        stats->produced_tokens = c[CTRL_TOKENS];
        stats->unfinished = c[CTRL_N_LEFT];
        stats->error = c[CTRL_ERROR] == 1 ? TTB_ERR_REF_INDEX : (c[CTRL_ERROR] == 2 ? TTB_ERR_REF_SHAPE : 0);
        stats->gpu_launches = (int32_t)(e->launches - launches0);
        stats->gpu_ms = ms;
    }
    if (c[CTRL_ERROR] == 1) {
        set_last_error("index out of bounds while splicing drafts into the token matrix (reference scatter, speculative_decoding.py:111)");
        return TTB_ERR_REF_INDEX;
    }
    if (c[CTRL_ERROR] == 2) {
        set_last_error("shape mismatch: a finished row is wider than max_len (reference speculative_decoding.py:158)");
        return TTB_ERR_REF_SHAPE;
    }
    return 0;
}


// ---- speculative beam search (speculative_decoding.py:428-598), host loop + device kernels ---------------
// The per-iteration shapes (token-matrix width, draft length, live rows) depend on the hypotheses, so the
// host reads a handful of control words back once per iteration; tokens, scores and logits never leave
// the device.
template <typename ActT>
static int beam_api(ttb_engine* e, const int64_t* src_dev, int B, int Ls, int max_len, int K, int draft_len, int N,
                    int pad, int bos, int eos, int c_token, int tie_break, int64_t* out_dev, int32_t* out_width,
                    int32_t* trace_nacc, int32_t* trace_pick, ttb_generate_stats* stats, cudaStream_t user_stream, bool smart) {
    const int E = e->E(), H = e->d.num_heads, HD = e->HD(), V = e->d.tgt_vocab_size;
    const int n_dec = (int)e->dec.size();
    cudaStream_t s = e->stream;
    TTB_CUDA_OK(cudaEventRecord(e->join_ev, user_stream));
    TTB_CUDA_OK(cudaStreamWaitEvent(s, e->join_ev, 0));
    const long long launches0 = e->launches;
    const int D0 = std::min(std::max(5, draft_len), 200);       // speculative_decoding.py:278-284
    // smart_drafts_mode (:600-615): library of Ls - 5 windows of draft_len + 1 tokens (BOS column included), first token = key
    const int D_lib = std::min(std::max(5, D0 + 1), 200);
    const int n_lib = Ls - 5;
    if (smart) TTB_CHECK(n_lib > 0, "The number of drafts must be greater than 0");
    const long long TS = (long long)B * Ls;
    // buffers that depend on the source length are sized for a rounded-up capacity: batches of different length must not
    // reallocate (cudaFree synchronises the whole device, i.e. every other engine decoding on this GPU as well)
    const int Ls_cap = std::max(256, (Ls + 63) / 64 * 64);
    const long long TS_cap = (long long)B * Ls_cap;
    const int Cmax = B * K, Rmax = Cmax * N;
    const int ldw = max_len + D0 + 4;
    // KV-cached pass (default): only the dl+1 scored positions of every (candidate, draft) row go through the decoder,
    // the prefix of a candidate is served from its self-attention cache, which follows the hypotheses through the
    // re-parenting of every step.  TTB_BEAM_NO_CACHE=1 keeps the full-prefix recomputation (A/B comparisons).
    static const bool no_cache = [] { const char* v = getenv("TTB_BEAM_NO_CACHE"); return v && v[0] == '1'; }();
    const bool cached = !no_cache;
    TTB_CHECK(cached || !smart, "smart_drafts_mode needs the KV-cached decoder pass (unset TTB_BEAM_NO_CACHE)");
    const long long Tc = (long long)Rmax * (D0 + 1);
    const long long Tmax = cached ? Tc : (long long)Rmax * ldw;

    if (e->src32.ensure(TS_cap * sizeof(int)) || e->memory.ensure(TS_cap * E * sizeof(float))) return 1;
    if (Prec<ActT>::lowp && e->memh.ensure(TS_cap * E * sizeof(ActT))) return 1;
    if (e->crosskv.ensure(TS_cap * 2 * E * sizeof(ActT) * n_dec)) return 1;
    if (e->drafts.ensure(smart ? (size_t)B * std::max(n_lib, Ls_cap) * D_lib * sizeof(int) : (size_t)B * N * D0 * sizeof(int))) return 1;
    if (ensure_work<ActT>(e, std::max(Tmax, TS_cap), cached ? n_dec : 1)) return 1;
    const long long cache_c_stride = (long long)ldw * E, cache_l_stride = (long long)Cmax * ldw * E;
    if (cached) {
        const size_t cbytes = (size_t)n_dec * cache_l_stride * sizeof(ActT);
        if (e->kcache.ensure(cbytes) || e->vcache.ensure(cbytes) || e->kcache2.ensure(cbytes) || e->vcache2.ensure(cbytes)) return 1;
    }
    DevBuf& bb = e->beam;
    // one arena for the small beam buffers
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_cand0 = take((size_t)Cmax * ldw * 4), o_cand1 = take((size_t)Cmax * ldw * 4);
    const size_t o_lp0 = take(Cmax * 4), o_lp1 = take(Cmax * 4);
    const size_t o_slot = take(Cmax * 4), o_fin = take(Cmax * 4), o_base = take(Cmax * 4), o_nacc = take((size_t)Cmax * N * 4);
    const size_t o_pick = take(Cmax * 4), o_acc = take(Cmax * 4), o_ctrl = take(32 * 4);   // control words + loop-control accumulators
    const size_t o_rows = take((size_t)Rmax * ldw * 4), o_rc = take(Rmax * 4), o_rq = take(Rmax * 4), o_rs = take(Rmax * 4);
    const size_t n_rp = (size_t)Rmax * (D0 + 1);
    const size_t o_topv = take(n_rp * K * 4), o_topi = take(n_rp * K * 4), o_keep = take(n_rp * 4), o_max = take(n_rp * 4), o_sum = take(n_rp * 4);
    const size_t o_xg = take(n_rp * E * 4), o_xgh = take(n_rp * E * 2), o_lg = take(n_rp * V * 4), o_rt = take(n_rp * 4), o_tv = take(n_rp * 4);
    const size_t o_lc = take(Rmax * 4), o_lq = take(Rmax * 4), o_cf = take(Cmax * 4), o_np = take(Cmax * 4), o_nk = take(Cmax * 4), o_nr = take(Cmax * 4);
    const size_t o_tc = take(smart ? (size_t)B * V * 4 : 4), o_tl = take(smart ? (size_t)B * V * N * 4 : 4), o_cc = take(Cmax * 4),
                 o_cl = take(Cmax * 4), o_rd = take(Rmax * 4), o_ds = take(Rmax * 16), o_dc = take(Rmax * 16);
    if (bb.ensure(off)) return 1;
    char* base = bb.as<char>();

    TTB_CUDA_OK(cudaEventRecord(e->t0, s));
    int* src32 = e->src32.as<int>();
    { Scope sc(e, KC_MISC, s); launch_i64_to_i32(reinterpret_cast<const long long*>(src_dev), src32, TS, s); }
    float* mem = e->memory.as<float>();
    ActT* memh = Prec<ActT>::lowp ? e->memh.as<ActT>() : nullptr;
    if (encode_impl<ActT>(e, src32, src32, B, Ls, mem, memh, s)) return 1;
    ActT* crosskv = e->crosskv.as<ActT>();
    const long long ckv_l_stride = TS_cap * 2 * E;
    if (cross_kv_impl<ActT>(e, mem, memh, (int)TS, crosskv, s, ckv_l_stride)) return 1;
    if (smart) { Scope sc(e, KC_MISC, s); launch_make_drafts(src32, Ls, B, Ls, D_lib, n_lib, eos, pad, c_token, e->drafts.as<int>(), s); }
    else { Scope sc(e, KC_MISC, s); launch_make_drafts(src32 + 1, Ls, B, Ls - 1, D0, N, eos, pad, c_token, e->drafts.as<int>(), s); }

    BeamState st{};
    st.B = B; st.K = K; st.N = N; st.dl0 = smart ? D_lib : D0; st.V = V; st.smart = smart ? 1 : 0; st.n_lib = n_lib; st.pad = pad; st.bos = bos; st.eos = eos; st.ldw = ldw; st.tie_break = tie_break;
    st.max_len = max_len;
    {   // width of the token matrix in the first iteration (the loop below applies the same rule)
        const int dl_first = std::min(max_len - 2, smart ? D_lib - 1 : D0);
        st.w0 = 1 + std::max(dl_first + 1, 0);
    }
    st.cand_cur = (int*)(base + o_cand0); st.cand_next = (int*)(base + o_cand1);
    st.logp_cur = (float*)(base + o_lp0); st.logp_next = (float*)(base + o_lp1);
    st.drafts = e->drafts.as<int>();
    st.c_slot0 = (int*)(base + o_slot); st.c_fin = (int*)(base + o_fin); st.c_rowbase = (int*)(base + o_base);
    st.c_nacc = (int*)(base + o_nacc); st.c_pick = (int*)(base + o_pick); st.acc_stat = (int*)(base + o_acc); st.ctrl = (int*)(base + o_ctrl);
    st.rows_tok = (int*)(base + o_rows); st.row_cand = (int*)(base + o_rc); st.row_query = (int*)(base + o_rq); st.row_slot0 = (int*)(base + o_rs);
    st.topv = (float*)(base + o_topv); st.topi = (int*)(base + o_topi); st.nkeep = (int*)(base + o_keep);
    st.lmax = (float*)(base + o_max); st.lsum = (float*)(base + o_sum);
    st.trace_nacc = trace_nacc; st.trace_pick = trace_pick;
    {
        void* dp = nullptr;
        for (int i = 16; i < 24; ++i) e->h_ctrl[i] = 0;   // ring of four 64-bit words the expand kernel posts into
        st.host_ctrl = cudaHostGetDevicePointer(&dp, e->h_ctrl + 16, 0) == cudaSuccess ? static_cast<int*>(dp) : nullptr;
        (void)cudaGetLastError();
    }
    if (cached) { st.row_tok = (int*)(base + o_rt); st.tokv = (float*)(base + o_tv); }
    st.live_cand = (int*)(base + o_lc); st.live_query = (int*)(base + o_lq); st.c_front = (int*)(base + o_cf);
    st.n_parent = (int*)(base + o_np); st.n_keep = (int*)(base + o_nk); st.n_row = (int*)(base + o_nr);
    st.tok_cnt = (int*)(base + o_tc); st.tok_list = (int*)(base + o_tl); st.c_cnt = (int*)(base + o_cc); st.c_last = (int*)(base + o_cl);
    st.row_draft = (int*)(base + o_rd);
    static const bool no_desc = [] { const char* v = getenv("TTB_BEAM_NO_DESC"); return v && v[0] == '1'; }();
    if (cached && !no_desc) {
        // per-query source length: the cross-attention kernels skip the padding behind it
        if (e->srclen.ensure((size_t)B * sizeof(int))) return 1;
        { Scope sc(e, KC_MISC, s); launch_row_lengths(src32, B, Ls, e->d.src_pad_token_idx, e->srclen.as<int>(), s); }
        st.desc_self = (int4*)(base + o_ds); st.desc_cross = (int4*)(base + o_dc); st.src_len = e->srclen.as<int>();
    }
    ActT* kc_cur = cached ? e->kcache.as<ActT>() : nullptr;
    ActT* vc_cur = cached ? e->vcache.as<ActT>() : nullptr;
    ActT* kc_next = cached ? e->kcache2.as<ActT>() : nullptr;
    ActT* vc_next = cached ? e->vcache2.as<ActT>() : nullptr;
    const int* n_live_cands = st.ctrl + BC_NLIVE_CANDS;
    float* xg = (float*)(base + o_xg);
    ActT* xgh = Prec<ActT>::lowp ? (ActT*)(base + o_xgh) : nullptr;
    float* logits = (float*)(base + o_lg);
    { Scope sc(e, KC_MISC, s); launch_beam_init(st, s); }
    if (smart) { Scope sc(e, KC_MISC, s); launch_beam_build_lib(st, s); }

    float* x = e->x.as<float>();
    ActT* xh = Prec<ActT>::lowp ? e->xh.as<ActT>() : nullptr;
    const int* n_live = st.ctrl + BC_NLIVE_ROWS;
    int W = 1, empty_cols = 0, filled = 1, budget = max_len - filled - 1, dl = smart ? D_lib - 1 : D0, C = B, beam = 1, iters = 0;
    // CUDA graphs of the steady-state iteration (one per parity), kept across calls while nothing they bake in changes
    const int dl_steady = dl;
    // Opt-in (TTB_BEAM_GRAPH=1, read per call): measured neutral on one B200 -- the search is bound by the GPU work of its
    // 5-20 k token rows, not by the ~33 launches per iteration (63.8 ms per batch launched directly, 64.5 ms replayed, the
    // graphs being re-captured whenever the padded source length changes), see DESIGN.md §2b.
    const char* ng = getenv("TTB_BEAM_GRAPH");
    const bool use_graph = cached && st.host_ctrl && !trace_nacc && !trace_pick && e->prof.mask == 0 && ng && ng[0] == '1';
    if (use_graph) {
        const long long key[20] = {B, K, N, D0, Ls, max_len, V, smart ? 1 : 0, pad, bos, eos, c_token, tie_break, e->alloc_signature(),
                                   (long long)sizeof(ActT), (long long)reinterpret_cast<uintptr_t>(st.host_ctrl), n_dec, 0, 0, 0};
        if (memcmp(key, e->beam_graph_key, sizeof(key)) != 0) {
            for (auto& gp : e->beam_graph) for (auto& g : gp) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
            memcpy(e->beam_graph_key, key, sizeof(key));
        }
    }
    int* hc = e->h_ctrl;
    hc[BC_ERROR] = 0;
    // the two hypothesis / score / cache buffer sets: iteration i (0-based) reads set i & 1 and writes the other one
    int* const cand_set[2] = {st.cand_cur, st.cand_next};
    float* const logp_set[2] = {st.logp_cur, st.logp_next};
    ActT* const kc_set[2] = {kc_cur, kc_next};
    ActT* const vc_set[2] = {vc_cur, vc_next};

    // One iteration = a fixed kernel sequence whose arguments depend on (C, beam, dl) and on the ping-pong parity of the
    // hypothesis / cache buffers only: the width of the token matrix and the iteration number live in device memory
    // (beam.cu: BCX_W, BCX_ITER), so the steady state (C = B K, full draft length) can be replayed as a CUDA graph and
    // -- because the device also applies the stop rule itself (BCX_DONE) -- be enqueued AHEAD of the host reading the
    // outcome of the iteration before it.
    // `Cb`: upper bound of the candidates that are still unfinished in this iteration (exact when the host has read the word
    // of the iteration before, C otherwise): the decoder-side grids are sized for Cb * N rows instead of C * N, so that the
    // no-op CTAs of finished candidates do not queue for SMs that the other batches in flight are using (§4c of DESIGN.md)
    auto enqueue_iteration = [&](int it, int C, int beam, int dl, int Cb) -> int {
        const int parity = it & 1;
        st.cand_cur = cand_set[parity]; st.cand_next = cand_set[parity ^ 1];
        st.logp_cur = logp_set[parity]; st.logp_next = logp_set[parity ^ 1];
        kc_cur = kc_set[parity]; kc_next = kc_set[parity ^ 1];
        vc_cur = vc_set[parity]; vc_next = vc_set[parity ^ 1];
        // replayed iterations bake their grids in: the live-candidate bound is rounded up to one of eight buckets of C
        const bool graph_it = use_graph && C == B * K && beam == K && dl == dl_steady;
        int gb = 0;
        if (graph_it && !smart) {
            const int want = std::min(C, std::max(1, Cb));
            while (gb + 1 < ttb_engine::kBeamBuckets && (C * (ttb_engine::kBeamBuckets - gb - 1) + ttb_engine::kBeamBuckets - 1) / ttb_engine::kBeamBuckets >= want) ++gb;
            Cb = (C * (ttb_engine::kBeamBuckets - gb) + ttb_engine::kBeamBuckets - 1) / ttb_engine::kBeamBuckets;
        }
        auto body = [&]() -> int {
        const int R = (smart ? C : std::min(C, std::max(1, Cb))) * N;
        bool fused_stats = false;
        { Scope sc(e, KC_MISC, s); launch_beam_prepare(st, C, beam, dl, s); }
        if (cached) {
            RowCount rows(R * (dl + 1), n_live, dl + 1);
            { Scope sc(e, KC_EMBED, s); launch_beam_embed_cached<ActT>(st, beam, R, dl, e->tgt_emb, e->pe, E, x, xh, s); }
            // attention groups: the N rows of a live candidate share its cache prefix; in smart mode candidates own a
            // different number of rows, so every row is a group of its own
            const int G = smart ? R : R / N, n_per_group = smart ? 1 : N;
            auto self_attn = [&](int l, ActT* qkv, ActT* att) {
                spec_attn(qkv, 3 * E, kc_cur + l * cache_l_stride, vc_cur + l * cache_l_stride, cache_c_stride, E, att, E, G, n_live_cands,
                          st.live_cand, st.c_front, st.cand_cur, ldw, e->d.tgt_pad_token_idx, n_per_group, dl, H, HD, ldw, s, st.desc_self);
            };
            auto cross_attn = [&](int l, ActT* q2, ActT* att) {
                const ActT* kv = crosskv + (long long)l * ckv_l_stride;
                attn(q2, E, kv, kv + E, 2 * E, att, E, G, n_live_cands, n_per_group * (dl + 1), Ls, Ls, st.live_query,
                     src32, Ls, e->d.src_pad_token_idx, false, H, HD, s, nullptr, nullptr, st.desc_cross);
            };
            if (decoder_stack<ActT>(e, rows, n_dec, Tc * 3 * E, self_attn, cross_attn, s)) return 1;
            // every decoder row is a scored position.  bf16 path: the projection is fused with the statistics the search reads
            // (logits stay in tensor memory); otherwise logits straight from the residual stream + beam_stats below
            if constexpr (Prec<ActT>::lowp) {
                const char* nf = getenv("TTB_NO_FUSED_STATS");
                if (!(nf && nf[0] == '1')) {
                    Scope sc(e, KC_GEMM_CLASSIFIER, s);
                    const int rc = launch_classifier_stats(xh, E, e->classifier.wh, e->classifier.b, rows, V, E, K, st.row_tok, st.tokv, st.lmax, st.lsum,
                                                           st.nkeep, st.topv, st.topi, s);
                    if (rc > 0) return 1;
                    fused_stats = rc == 0;
                    if (!fused_stats) e->launches--;
                }
            }
            if (!fused_stats && linear<float>(e, KC_GEMM_CLASSIFIER, a_view<ActT>(x, xh), E, e->classifier, logits, V, rows, false, s)) return 1;
        } else {
            const int Wit = W;   // full-prefix recomputation (A/B path): never enqueued ahead, the host's width is exact
            { Scope sc(e, KC_MISC, s); launch_beam_fill_rows(st, C, beam, Wit, dl, s); }
            RowCount rows(R * Wit, n_live, Wit);
            { Scope sc(e, KC_EMBED, s); launch_embed_seq_rows<ActT>(st.rows_tok, rows, Wit, e->tgt_emb, e->pe, E, x, xh, s); }
            auto self_attn = [&](int, ActT* qkv, ActT* att) {
                attn(qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, R, n_live, Wit, Wit, Wit, nullptr,
                     st.rows_tok, Wit, e->d.tgt_pad_token_idx, true, H, HD, s);
            };
            auto cross_attn = [&](int l, ActT* q2, ActT* att) {
                const ActT* kv = crosskv + (long long)l * ckv_l_stride;
                attn(q2, E, kv, kv + E, 2 * E, att, E, R, n_live, Wit, Ls, Ls, st.row_query,
                     src32, Ls, e->d.src_pad_token_idx, false, H, HD, s);
            };
            if (decoder_stack<ActT>(e, rows, 1, 0, self_attn, cross_attn, s)) return 1;
            { Scope sc(e, KC_MISC, s); launch_beam_gather<ActT>(st, x, xh, R, Wit, dl, E, Prec<ActT>::lowp ? nullptr : xg, xgh, s); }
            RowCount rp_rows(R * (dl + 1), n_live, dl + 1);
            if (linear<float>(e, KC_GEMM_CLASSIFIER, a_view<ActT>(xg, xgh), E, e->classifier, logits, V, rp_rows, false, s)) return 1;
        }
        if (!fused_stats) { Scope sc(e, KC_ARGMAX, s); launch_beam_stats(st, logits, R, dl, s); }
        { Scope sc(e, KC_ACCEPT, s); launch_beam_choose(st, C, beam, dl, s); }
        // the expand kernel also closes the iteration: its last CTA writes the control words and, last, the iteration's
        // sequence number into pinned host memory; the host reads them while the caches are still being re-parented
        { Scope sc(e, KC_ACCEPT, s); launch_beam_expand(st, beam, dl, logits, s); }
        if (cached) {
            Scope sc(e, KC_CACHE_APPEND, s);
            launch_beam_cache_update<ActT>(st, dl, e->qkv.as<ActT>(), Tc * 3 * E, n_dec, 3 * E, E, kc_cur, vc_cur, kc_next, vc_next,
                                           cache_l_stride, cache_c_stride, s);
        }
        return 0;
        };   // body
        if (graph_it) {
            if (!e->beam_graph[parity][gb]) {
                const long long l0 = e->launches;
                cudaGraph_t graph = nullptr;
                TTB_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
                const int rc = body();
                const cudaError_t ce = cudaStreamEndCapture(s, &graph);
                if (rc || ce != cudaSuccess) {
                    if (graph) cudaGraphDestroy(graph);
                    if (!rc) set_last_error(std::string("CUDA graph capture of the beam-search iteration failed: ") + cudaGetErrorString(ce));
                    return 1;
                }
                TTB_CUDA_OK(cudaGraphInstantiate(&e->beam_graph[parity][gb], graph, 0));
                cudaGraphDestroy(graph);
                e->beam_graph_launches = e->launches - l0;
                e->launches = l0;   // capturing did not launch anything
            }
            TTB_CUDA_OK(cudaGraphLaunch(e->beam_graph[parity][gb], s));
            e->launches += e->beam_graph_launches;
            return 0;
        }
        return body();
    };
    // wait for the packed word the expand kernel posts for iteration `seq` (sequence | error | all finished | empty columns)
    static const bool pure_spin = [] { const char* v = getenv("TTB_BEAM_SPIN"); return v && v[0] == '1'; }();
    int live_next = B;      // unfinished candidates entering the next iteration, from the newest word read
    auto wait_word = [&](int seq) -> int {
        volatile unsigned long long* word = reinterpret_cast<volatile unsigned long long*>(hc + 16) + (seq & 3);
        unsigned long long w = 0;
        bool seen = false;
        for (long spin = 0; !seen; ++spin) {
            w = *word;
            seen = (unsigned)((w >> 40) & 0xffffff) == (unsigned)(seq & 0xffffff);
            // the word normally arrives within a few microseconds of the last launch call; past that (long iterations,
            // several engines per GPU and ranks per box sharing the cores) the thread gives its core away between polls
            if (!seen) { if (spin < 256 || pure_spin) cpu_relax(); else sched_yield(); }
            if (!seen && (spin & 0x3FFF) == 0x3FFF) {
                const cudaError_t q = cudaStreamQuery(s);
                if (q == cudaSuccess) {   // stream drained: the word is there, or the mapping is unavailable: take a copy
                    w = *word;
                    if ((unsigned)((w >> 40) & 0xffffff) != (unsigned)(seq & 0xffffff)) {
                        TTB_CUDA_OK(cudaMemcpy(hc, st.ctrl, BC_COUNT * sizeof(int), cudaMemcpyDeviceToHost));
                        w = ((unsigned long long)(seq & 0xffffff) << 40) | ((unsigned long long)(hc[BC_ERROR] & 0xf) << 36) |
                            ((unsigned long long)(hc[BC_ALL_FINISHED] & 1) << 35) | ((unsigned long long)(hc[BC_EMPTY_COLS] & 0xfff) << 23) |
                            ((unsigned long long)0xffff << 7);   // live candidates unknown: full grids
                    }
                    seen = true;
                } else if (q != cudaErrorNotReady) {
                    TTB_CUDA_OK(q);
                }
            }
        }
        hc[BC_ERROR] = (int)((w >> 36) & 0xf);
        hc[BC_ALL_FINISHED] = (int)((w >> 35) & 1);
        hc[BC_EMPTY_COLS] = (int)((w >> 23) & 0xfff);
        live_next = (int)((w >> 7) & 0xffff);
        return 0;
    };
    // Host loop.  `iters` iterations have been read back (their outcome is known exactly: W, filled, budget, dl below are
    // the reference's values before iteration `iters`); `enq` iterations have been enqueued.  In the steady state the next
    // iteration is enqueued BEFORE the word of the current one is awaited whenever its arguments cannot depend on that
    // word: C = B K and beam = K from the second iteration on, and the draft length stays dl as long as the length budget
    // cannot drop below it (one iteration fills at most dl + 1 more columns).  If the current iteration ends the loop the
    // device has raised BCX_DONE and the iteration enqueued ahead does nothing.
    static const bool no_ahead = [] { const char* v = getenv("TTB_BEAM_NO_AHEAD"); return v && v[0] == '1'; }();
    const bool ahead_ok = cached && st.host_ctrl && !no_ahead && !trace_nacc && !trace_pick;
    int enq = 0, dl_enq = dl;
    bool ended = false;
    while (!ended) {
        if (enq == iters) {                              // nothing in flight: the coming iteration from exact knowledge
            if (!(budget >= 1 && filled <= max_len)) break;
            dl = std::min(budget, dl);
            const int grow = dl + 1 - empty_cols;
            if (grow > 0) W += grow;
            TTB_CHECK(W <= ldw, "beam search token matrix outgrew its buffer");
            if (enqueue_iteration(enq, C, beam, dl, std::min(live_next, C))) return 1;
            dl_enq = dl;
            ++enq;
        }
        // ahead of the word only while every candidate is alive (iteration 0 aside): once hypotheses finish, the exact live
        // count of the word sizes the next iteration's grids, which is worth more than the host's head start (both measured)
        if (ahead_ok && enq == iters + 1 && dl_enq == dl_steady && max_len - (filled + 2 * (dl_enq + 1)) - 1 >= dl_enq &&
            W + 2 * (dl_enq + 1) <= ldw && (iters == 0 || live_next >= B * K)) {
            // `filled` is the exact count before the iteration in flight: after it and one more at most 2 (dl + 1) columns are added
            if (enqueue_iteration(enq, B * K, K, dl_enq, B * K)) return 1;
            ++enq;
        }
        if (wait_word(iters + 1)) return 1;
        ++iters;
        if (hc[BC_ERROR]) break;
        C = B * K;
        beam = K;
        if (hc[BC_ALL_FINISHED]) break;
        empty_cols = hc[BC_EMPTY_COLS];
        filled = W - empty_cols;
        budget = max_len - filled - 1;
        if (enq > iters) {
            // the iteration enqueued ahead runs with the values the reference would use: same dl (checked above), and
            // its token matrix is this one's plus the columns the device adds with the same rule
            if (!(budget >= 1 && filled <= max_len)) { ended = true; break; }   // the device stopped as well (BCX_DONE)
            const int grow = dl + 1 - empty_cols;
            if (grow > 0) W += grow;
        }
    }
    // the hypotheses of the last iteration that really ran are in the set it wrote
    st.cand_cur = cand_set[iters & 1];
    TTB_CUDA_OK(cudaGetLastError());
    const int err = hc[BC_ERROR];
    if (!err) {
        TTB_CHECK(iters > 0, "max_len too small for the speculative beam search (the reference fails as well)");
        launch_beam_export(st.cand_cur, ldw, B * K, W, reinterpret_cast<long long*>(out_dev), s);
        e->launches++;
    }
    TTB_CUDA_OK(cudaEventRecord(e->t1, s));
    TTB_CUDA_OK(cudaStreamSynchronize(s));
    TTB_CUDA_OK(cudaGetLastError());
    prof_collect(e);
    if (out_width) *out_width = W;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->t0, e->t1);
    if (stats) {
        int fin_ctrl[BC_COUNT] = {};
        if (iters > 0) TTB_CUDA_OK(cudaMemcpy(fin_ctrl, st.ctrl, sizeof(fin_ctrl), cudaMemcpyDeviceToHost));
        stats->model_calls = iters;
        stats->accepted_tokens = fin_ctrl[BC_ACCEPTED];
        stats->produced_tokens = fin_ctrl[BC_PRODUCED];
        stats->unfinished = 0;
        stats->error = err;
        stats->gpu_launches = (int32_t)(e->launches - launches0);
        stats->gpu_ms = ms;
    }
    if (err == 3) {
        set_last_error("PAD predicted inside a hypothesis: draft slots are not contiguous (the reference fails on the reshape at speculative_decoding.py:524-526)");
        return TTB_ERR_REF_SHAPE;
    }
    if (err == 4) {
        set_last_error("fewer candidate continuations than n_best (reference assert in topk_in_each_group, speculative_decoding.py:195)");
        return TTB_ERR_REF_ASSERT;
    }
    return 0;
}

// ---- standard beam search (standard_decoding.py:90-174), host loop + device kernels -----------------------------
// One decoder call per generated column on the hypotheses that have not produced EOS (full prefix, no KV cache yet);
// tokens, scores and logits stay on the device, the host reads two control words per step for the stop test.
template <typename ActT>
static int std_beam_api(ttb_engine* e, const int64_t* src_dev, int B, int Ls, int max_len, int K, int pad, int bos, int eos,
                        int64_t* out_dev, int32_t* out_width, ttb_generate_stats* stats, cudaStream_t user_stream) {
    const int E = e->E(), H = e->d.num_heads, HD = e->HD(), V = e->d.tgt_vocab_size;
    const int n_dec = (int)e->dec.size();
    cudaStream_t s = e->stream;
    TTB_CUDA_OK(cudaEventRecord(e->join_ev, user_stream));
    TTB_CUDA_OK(cudaStreamWaitEvent(s, e->join_ev, 0));
    const long long launches0 = e->launches;
    const long long TS = (long long)B * Ls;
    const long long TS_cap = (long long)B * std::max(256, (Ls + 63) / 64 * 64);   // no reallocation from one batch length to the next
    const int Cmax = B * K, ldw = max_len;
    const long long Tmax = (long long)Cmax * max_len;

    if (e->src32.ensure(TS_cap * sizeof(int)) || e->memory.ensure(TS_cap * E * sizeof(float))) return 1;
    if (Prec<ActT>::lowp && e->memh.ensure(TS_cap * E * sizeof(ActT))) return 1;
    if (e->crosskv.ensure(TS_cap * 2 * E * sizeof(ActT) * n_dec)) return 1;
    {
        const char* v = getenv("TTB_SBEAM_NO_CACHE");
        if (ensure_work<ActT>(e, (v && v[0] == '1') ? std::max(Tmax, TS_cap) : std::max<long long>(Cmax, TS_cap), 1)) return 1;
    }
    DevBuf& bb = e->beam;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_y0 = take((size_t)Cmax * ldw * 4), o_y1 = take((size_t)Cmax * ldw * 4);
    const size_t o_s0 = take(Cmax * 4), o_s1 = take(Cmax * 4), o_fin = take(Cmax * 4), o_cr = take(Cmax * 4), o_rc = take(Cmax * 4),
                 o_rq = take(Cmax * 4), o_rows = take((size_t)Cmax * ldw * 4), o_tot = take((size_t)Cmax * V * 4), o_ctrl = take(64),
                 o_xg = take((size_t)Cmax * E * 4), o_xgh = take((size_t)Cmax * E * 2), o_lg = take((size_t)Cmax * V * 4),
                 o_cf = take(Cmax * 4), o_par = take(Cmax * 4), o_ds = take(Cmax * 16), o_dc = take(Cmax * 16);
    if (bb.ensure(off)) return 1;
    // KV-cached pass (default): one new token per live hypothesis and step, the prefix of a hypothesis is served from its
    // self-attention cache, which follows the hypotheses through the re-parenting of every step (two ping-pong copies).
    // TTB_SBEAM_NO_CACHE=1 keeps the reference-like full-prefix recomputation (A/B runs).
    static const bool sb_no_cache = [] { const char* v = getenv("TTB_SBEAM_NO_CACHE"); return v && v[0] == '1'; }();
    const bool cached = !sb_no_cache;
    const long long cache_c_stride = (long long)max_len * E, cache_l_stride = (long long)Cmax * max_len * E;
    if (cached) {
        const size_t cbytes = (size_t)n_dec * cache_l_stride * sizeof(ActT);
        if (e->kcache.ensure(cbytes) || e->vcache.ensure(cbytes) || e->kcache2.ensure(cbytes) || e->vcache2.ensure(cbytes)) return 1;
        if (e->srclen.ensure((size_t)B * sizeof(int))) return 1;
        if (ensure_work<ActT>(e, std::max<long long>(Cmax, TS_cap), n_dec)) return 1;
    }
    char* base = bb.as<char>();

    TTB_CUDA_OK(cudaEventRecord(e->t0, s));
    int* src32 = e->src32.as<int>();
    { Scope sc(e, KC_MISC, s); launch_i64_to_i32(reinterpret_cast<const long long*>(src_dev), src32, TS, s); }
    float* mem = e->memory.as<float>();
    ActT* memh = Prec<ActT>::lowp ? e->memh.as<ActT>() : nullptr;
    if (encode_impl<ActT>(e, src32, src32, B, Ls, mem, memh, s)) return 1;
    ActT* crosskv = e->crosskv.as<ActT>();
    if (cross_kv_impl<ActT>(e, mem, memh, (int)TS, crosskv, s)) return 1;

    StdBeamState st{};
    st.B = B; st.K = K; st.V = V; st.pad = pad; st.bos = bos; st.eos = eos; st.ldw = ldw;
    st.y_cur = (int*)(base + o_y0); st.y_next = (int*)(base + o_y1);
    st.score_cur = (float*)(base + o_s0); st.score_next = (float*)(base + o_s1);
    st.fin = (int*)(base + o_fin); st.cand_row = (int*)(base + o_cr); st.row_cand = (int*)(base + o_rc); st.row_query = (int*)(base + o_rq);
    st.rows_tok = (int*)(base + o_rows); st.total = (float*)(base + o_tot); st.ctrl = (int*)(base + o_ctrl);
    float* xg = (float*)(base + o_xg);
    ActT* xgh = Prec<ActT>::lowp ? (ActT*)(base + o_xgh) : nullptr;
    float* logits = (float*)(base + o_lg);
    ActT *kc_cur = nullptr, *vc_cur = nullptr, *kc_next = nullptr, *vc_next = nullptr;
    int* const y_set[2] = {st.y_cur, st.y_next};
    if (cached) {
        st.c_front = (int*)(base + o_cf); st.parent = (int*)(base + o_par);
        st.desc_self = (int4*)(base + o_ds); st.desc_cross = (int4*)(base + o_dc);
        { Scope sc(e, KC_MISC, s); launch_row_lengths(src32, B, Ls, e->d.src_pad_token_idx, e->srclen.as<int>(), s); }
        st.src_len = e->srclen.as<int>();
        kc_cur = e->kcache.as<ActT>(); vc_cur = e->vcache.as<ActT>(); kc_next = e->kcache2.as<ActT>(); vc_next = e->vcache2.as<ActT>();
    }
    { Scope sc(e, KC_MISC, s); launch_sbeam_init(st, s); }

    float* x = e->x.as<float>();
    ActT* xh = Prec<ActT>::lowp ? e->xh.as<ActT>() : nullptr;
    const int* n_live = st.ctrl;
    int* hc = e->h_ctrl;
    int W = 1, beam = 1, calls = 0;
    constexpr int sb_lag = 1;   // steps the host stays ahead of the control words it reads (blocking-sync events: the wait sleeps)
    // step 0 on the BOS column (:106), then at most max_len - 2 further columns (:127)
    for (int step = 0; step < max_len - 1; ++step) {
        const int C = B * beam;
        { Scope sc(e, KC_MISC, s); launch_sbeam_prepare(st, C, beam, W, s); }
        if (cached) {
            RowCount rows1(C, n_live, 1);
            { Scope sc(e, KC_EMBED, s); launch_sbeam_embed_last<ActT>(st, C, W, e->tgt_emb, e->pe, E, x, xh, s); }
            auto self_attn = [&](int l, ActT* qkv, ActT* att) {
                spec_attn(qkv, 3 * E, kc_cur + l * cache_l_stride, vc_cur + l * cache_l_stride, cache_c_stride, E, att, E, C, n_live, st.row_cand,
                          st.c_front, st.y_cur, ldw, e->d.tgt_pad_token_idx, 1, 0, H, HD, max_len, s, st.desc_self);
            };
            auto cross_attn = [&](int l, ActT* q2, ActT* att) {
                const ActT* kv = crosskv + (long long)l * TS * 2 * E;
                attn(q2, E, kv, kv + E, 2 * E, att, E, C, n_live, 1, Ls, Ls, st.row_query, src32, Ls, e->d.src_pad_token_idx, false, H, HD, s,
                     nullptr, nullptr, st.desc_cross);
            };
            if (decoder_stack<ActT>(e, rows1, n_dec, (long long)Cmax * 3 * E, self_attn, cross_attn, s)) return 1;
            // every live row is the last position of its hypothesis: logits straight from the residual stream
            if (linear<float>(e, KC_GEMM_CLASSIFIER, a_view<ActT>(x, xh), E, e->classifier, logits, V, rows1, false, s)) return 1;
            { Scope sc(e, KC_ARGMAX, s); launch_sbeam_scores(st, C, logits, s); }
            {
                Scope sc(e, KC_ACCEPT, s);
                TTB_CHECK(launch_sbeam_select(st, beam, W, s) == 0, "beam_size * vocabulary too large for the selection kernel");
            }
            {
                Scope sc(e, KC_CACHE_APPEND, s);
                launch_sbeam_cache_update<ActT>(st, W, e->qkv.as<ActT>(), (long long)Cmax * 3 * E, n_dec, 3 * E, E, kc_cur, vc_cur, kc_next, vc_next,
                                                cache_l_stride, cache_c_stride, s);
            }
            std::swap(kc_cur, kc_next);
            std::swap(vc_cur, vc_next);
            // the stop test runs on the device (std_beam.cu: ctrl[2]); the host reads the control words of step - 1 while
            // this step is already enqueued (a step behind the end is a no-op), so the GPU never waits for the host
            TTB_CUDA_OK(cudaMemcpyAsync(hc + 32 + (step & 3) * 8, st.ctrl, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
            TTB_CUDA_OK(cudaEventRecord(e->poll_ev[step & 3], s));
            std::swap(st.y_cur, st.y_next);
            std::swap(st.score_cur, st.score_next);
            beam = K;
            W += 1;
            if (step >= sb_lag) {
                const int j = (step - sb_lag) & 3;
                TTB_CUDA_OK(cudaEventSynchronize(e->poll_ev[j]));
                if (hc[32 + j * 8 + 2]) break;       // every hypothesis contains EOS (:169)
            }
            continue;
        }
        RowCount rows(C * W, n_live, W);
        { Scope sc(e, KC_EMBED, s); launch_embed_seq_rows<ActT>(st.rows_tok, rows, W, e->tgt_emb, e->pe, E, x, xh, s); }
        auto self_attn = [&](int, ActT* qkv, ActT* att) {
            attn(qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, C, n_live, W, W, W, nullptr,
                 st.rows_tok, W, e->d.tgt_pad_token_idx, true, H, HD, s);
        };
        auto cross_attn = [&](int l, ActT* q2, ActT* att) {
            const ActT* kv = crosskv + (long long)l * TS * 2 * E;
            attn(q2, E, kv, kv + E, 2 * E, att, E, C, n_live, W, Ls, Ls, st.row_query,
                 src32, Ls, e->d.src_pad_token_idx, false, H, HD, s);
        };
        if (decoder_stack<ActT>(e, rows, 1, 0, self_attn, cross_attn, s)) return 1;
        { Scope sc(e, KC_MISC, s); launch_sbeam_gather_last<ActT>(st, x, xh, C, W, E, Prec<ActT>::lowp ? nullptr : xg, xgh, s); }
        RowCount last_rows(C, n_live, 1);
        if (linear<float>(e, KC_GEMM_CLASSIFIER, a_view<ActT>(xg, xgh), E, e->classifier, logits, V, last_rows, false, s)) return 1;
        { Scope sc(e, KC_ARGMAX, s); launch_sbeam_scores(st, C, logits, s); }
        {
            Scope sc(e, KC_ACCEPT, s);
            TTB_CHECK(launch_sbeam_select(st, beam, W, s) == 0, "beam_size * vocabulary too large for the selection kernel");
        }
        TTB_CUDA_OK(cudaMemcpyAsync(hc, st.ctrl, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        TTB_CUDA_OK(cudaStreamSynchronize(s));
        ++calls;
        std::swap(st.y_cur, st.y_next);
        std::swap(st.score_cur, st.score_next);
        beam = K;
        W += 1;
        if (step > 0 && hc[1] == B * K) break;   // every hypothesis contains EOS (:169); the first step never breaks (:106-125)
    }
    if (cached) {
        // steps that really ran (the host may have enqueued one more, a no-op): width, call count, and the hypothesis set
        // the last real step wrote (step i reads set i & 1 and writes the other one)
        int fin[4] = {};
        TTB_CUDA_OK(cudaMemcpyAsync(fin, st.ctrl, sizeof(fin), cudaMemcpyDeviceToHost, s));
        TTB_CUDA_OK(cudaStreamSynchronize(s));
        calls = fin[3];
        W = 1 + calls;
        st.y_cur = (calls & 1) ? y_set[1] : y_set[0];
    }
    launch_beam_export(st.y_cur, ldw, B * K, W, reinterpret_cast<long long*>(out_dev), s);
    e->launches++;
    TTB_CUDA_OK(cudaEventRecord(e->t1, s));
    TTB_CUDA_OK(cudaStreamSynchronize(s));
    TTB_CUDA_OK(cudaGetLastError());
    prof_collect(e);
    if (out_width) *out_width = W;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->t0, e->t1);
    if (stats) {
        stats->model_calls = calls;
        stats->accepted_tokens = 0;
        stats->produced_tokens = 0;
        stats->unfinished = 0;
        stats->error = 0;
        stats->gpu_launches = (int32_t)(e->launches - launches0);
        stats->gpu_ms = ms;
    }
    return 0;
}

}  // namespace ttb

// =================================================================================================
extern "C" {

int ttb_abi_version(void) { return TTB_ABI_VERSION; }
const char* ttb_last_error(void) { return g_last_error.c_str(); }

int ttb_device_check(int device) {
    int n = 0;
    TTB_CUDA_OK(cudaGetDeviceCount(&n));
    TTB_CHECK(device >= 0 && device < n, "no such CUDA device");
    cudaDeviceProp p{};
    TTB_CUDA_OK(cudaGetDeviceProperties(&p, device));
    TTB_CHECK(p.major == 10, "libttb200 is built for sm_100a (B200) only");
    return 0;
}

int ttb_engine_create(const ttb_model_desc* desc, int device, ttb_engine** out) {
    TTB_CHECK(desc && out, "null argument");
    TTB_CHECK(desc->embedding_dim % desc->num_heads == 0, "embedding_dim must be divisible by num_heads");
    const int hd = desc->embedding_dim / desc->num_heads;
    TTB_CHECK(hd == 16 || hd == 32 || hd == 64, "head_dim must be 16, 32 or 64");
    TTB_CHECK(desc->embedding_dim % 16 == 0 && desc->feedforward_dim % 16 == 0, "embedding_dim and feedforward_dim must be multiples of 16");
    TTB_CHECK(desc->embedding_dim <= 1024, "embedding_dim above 1024 is not supported");
    if (desc->precision == TTB_PRECISION_BF16)
        TTB_CHECK(desc->embedding_dim % 64 == 0 && desc->feedforward_dim % 64 == 0, "bf16 path needs embedding_dim and feedforward_dim multiples of 64");
    if (int rc = ttb_device_check(device)) return rc;
    TTB_CUDA_OK(cudaSetDevice(device));
    ttb_engine* e = new ttb_engine();
    e->d = *desc;
    e->device = device;
    if (register_params(e)) { ttb_engine_destroy(e); return 1; }
    TTB_CUDA_OK(cudaMallocHost(&e->h_ctrl, 4 * CTRL_COUNT * sizeof(int)));
    // blocking sync: the decoding loop polls one graph (four iterations, ~1.2 ms) behind the GPU, so a sleeping host thread
    // costs nothing, and several engines per GPU x several ranks per box do not each burn a core spinning
    for (auto& ev : e->poll_ev) TTB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
    TTB_CUDA_OK(cudaEventCreate(&e->t0));
    TTB_CUDA_OK(cudaEventCreate(&e->t1));
    TTB_CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    TTB_CUDA_OK(cudaEventCreateWithFlags(&e->join_ev, cudaEventDisableTiming));
    *out = e;
    return 0;
}

void ttb_engine_destroy(ttb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (auto& kv : e->params) {
        if (kv.second.dev) cudaFree(kv.second.dev);
        if (kv.second.devh) cudaFree(kv.second.devh);
    }
    DevBuf* bufs[] = {&e->x, &e->xh, &e->y, &e->qkv, &e->att, &e->q2, &e->hid, &e->logits, &e->tok32, &e->keytok32, &e->pred,
                      &e->src32, &e->memory, &e->memh, &e->crosskv, &e->kcache, &e->vcache, &e->drafts, &e->gen, &e->front,
                      &e->active, &e->ctrl, &e->sel, &e->out64, &e->hist, &e->beam, &e->srclen, &e->desc, &e->kcache2, &e->vcache2, &e->crossvt};
    for (DevBuf* b : bufs) b->release();
    if (e->h_ctrl) cudaFreeHost(e->h_ctrl);
    for (auto& ev : e->poll_ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->prof.pool) cudaEventDestroy(ev);
    for (auto& gb : e->graph_exec) for (auto& g : gb) if (g) cudaGraphExecDestroy(g);
    for (auto& gp : e->beam_graph) for (auto& g : gp) if (g) cudaGraphExecDestroy(g);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->join_ev) cudaEventDestroy(e->join_ev);
    if (e->t0) cudaEventDestroy(e->t0);
    if (e->t1) cudaEventDestroy(e->t1);
    delete e;
}

int ttb_engine_set_param(ttb_engine* e, const char* name, const float* data, int64_t numel) {
    TTB_CHECK(e && name && data, "null argument");
    std::string n(name);
    if (n.rfind("model.", 0) == 0) n = n.substr(6);
    auto it = e->params.find(n);
    TTB_CHECK(it != e->params.end(), std::string("unexpected parameter name: ") + n);
    TTB_CHECK(it->second.numel == numel, std::string("size mismatch for ") + n + ": expected " +
              std::to_string(it->second.numel) + " elements, got " + std::to_string(numel));
    TTB_CUDA_OK(cudaSetDevice(e->device));
    TTB_CUDA_OK(cudaMemcpy(it->second.dev, data, (size_t)numel * sizeof(float), cudaMemcpyDefault));
    it->second.set = true;
    e->finalized = false;
    return 0;
}

int ttb_engine_finalize(ttb_engine* e) {
    TTB_CHECK(e, "null engine");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    for (auto& kv : e->params) TTB_CHECK(kv.second.set, std::string("missing parameter: ") + kv.first);
    if (e->d.precision == TTB_PRECISION_BF16) {
        for (auto& kv : e->params) {
            Param& p = kv.second;
            if (!p.is_gemm_weight) continue;
            if (!p.devh) TTB_CUDA_OK(cudaMalloc(&p.devh, (size_t)p.numel * sizeof(__nv_bfloat16)));
            launch_f32_to_bf16(p.dev, p.devh, p.numel, 0);
        }
        TTB_CUDA_OK(cudaDeviceSynchronize());
    }
    const int E = e->d.embedding_dim, F = e->d.feedforward_dim;
    e->enc.clear();
    e->dec.clear();
    for (int i = 0; i < e->d.num_encoder_layers; ++i) {
        std::string p = "transformer.encoder.layers." + std::to_string(i);
        EncLayer L;
        L.in_proj = make_lin(e, p + ".self_attn.in_proj_weight", p + ".self_attn.in_proj_bias", 3 * E, E);
        L.out_proj = make_lin(e, p + ".self_attn.out_proj.weight", p + ".self_attn.out_proj.bias", E, E);
        L.ff1 = make_lin(e, p + ".linear1.weight", p + ".linear1.bias", F, E);
        L.ff2 = make_lin(e, p + ".linear2.weight", p + ".linear2.bias", E, F);
        L.n1 = make_norm(e, p + ".norm1");
        L.n2 = make_norm(e, p + ".norm2");
        e->enc.push_back(L);
    }
    for (int i = 0; i < e->d.num_decoder_layers; ++i) {
        std::string p = "transformer.decoder.layers." + std::to_string(i);
        DecLayer L;
        L.self_in = make_lin(e, p + ".self_attn.in_proj_weight", p + ".self_attn.in_proj_bias", 3 * E, E);
        L.self_out = make_lin(e, p + ".self_attn.out_proj.weight", p + ".self_attn.out_proj.bias", E, E);
        L.cross_in = make_lin(e, p + ".multihead_attn.in_proj_weight", p + ".multihead_attn.in_proj_bias", 3 * E, E);
        L.cross_out = make_lin(e, p + ".multihead_attn.out_proj.weight", p + ".multihead_attn.out_proj.bias", E, E);
        L.ff1 = make_lin(e, p + ".linear1.weight", p + ".linear1.bias", F, E);
        L.ff2 = make_lin(e, p + ".linear2.weight", p + ".linear2.bias", E, F);
        L.n1 = make_norm(e, p + ".norm1");
        L.n2 = make_norm(e, p + ".norm2");
        L.n3 = make_norm(e, p + ".norm3");
        e->dec.push_back(L);
    }
    e->enc_norm = make_norm(e, "transformer.encoder.norm");
    e->dec_norm = make_norm(e, "transformer.decoder.norm");
    e->classifier = make_lin(e, "next_token_classifier.weight", "next_token_classifier.bias", e->d.tgt_vocab_size, E);
    e->src_emb = e->params["src_token_featurizer.embedding.weight"].dev;
    e->tgt_emb = e->params["tgt_token_featurizer.embedding.weight"].dev;
    e->pe = e->params["positional_encoding.pe"].dev;
    e->finalized = true;
    return 0;
}

int ttb_make_drafts(const int64_t* src_dev, int64_t src_ld, int32_t B, int32_t L, int32_t draft_len,
                    int32_t n_drafts, int32_t min_draft_len, int32_t max_draft_len, int32_t eos,
                    int32_t pad, int32_t replace, int64_t* out_dev, int32_t* d_out, void* stream) {
    // same argument checks, same order, as drafting.py:38-42 (the wrappers raise AssertionError)
    TTB_CHECK(n_drafts > 0, "The number of drafts must be greater than 0");
    TTB_CHECK(min_draft_len <= max_draft_len, "The minimum draft length must not be greater than the maximum draft length");
    TTB_CHECK(pad != replace, "The pad token and the replace token must be different");
    TTB_CHECK(eos != replace, "The eos token and the replace token must be different");
    TTB_CHECK(eos != pad, "The eos token and the pad token must be different");
    TTB_CHECK(src_dev && out_dev && B > 0 && L > 0, "bad source tensor");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int D = std::min(std::max(min_draft_len, draft_len), max_draft_len);
    if (d_out) *d_out = D;
    int *src32 = nullptr, *out32 = nullptr;
    TTB_CUDA_OK(cudaMallocAsync(&src32, (size_t)B * L * sizeof(int), s));
    TTB_CUDA_OK(cudaMallocAsync(&out32, (size_t)B * n_drafts * D * sizeof(int), s));
    if (src_ld == L) {
        launch_i64_to_i32(reinterpret_cast<const long long*>(src_dev), src32, (long long)B * L, s);
    } else {
        for (int b = 0; b < B; ++b)
            launch_i64_to_i32(reinterpret_cast<const long long*>(src_dev) + (long long)b * src_ld, src32 + (long long)b * L, L, s);
    }
    launch_make_drafts(src32, L, B, L, D, n_drafts, eos, pad, replace, out32, s);
    launch_i32_to_i64(out32, reinterpret_cast<long long*>(out_dev), (long long)B * n_drafts * D, s);
    TTB_CUDA_OK(cudaFreeAsync(src32, s));
    TTB_CUDA_OK(cudaFreeAsync(out32, s));
    TTB_CUDA_OK(cudaGetLastError());
    return 0;
}

#define TTB_DISPATCH(e, call_f32, call_bf16) ((e)->d.precision == TTB_PRECISION_FP32 ? (call_f32) : (call_bf16))

// Every exit of a decoding loop leaves the engine's stream drained: kernels still queued there write to the caller's
// `out` / `trace` buffers, which the caller is free to release as soon as the call returns (error exits included).
static int drained(ttb_engine* e, int rc) {
    if (rc != 0 && e->stream) {
        const std::string keep = g_last_error;
        cudaStreamSynchronize(e->stream);
        (void)cudaGetLastError();
        g_last_error = keep;
    }
    return rc;
}

int ttb_encode_src(ttb_engine* e, const int64_t* src_dev, const uint8_t* src_pad_mask_dev, int32_t B,
                   int32_t Ls, float* memory_out_dev, void* stream) {
    TTB_CHECK(e && e->finalized, "engine not finalized");
    TTB_CHECK(src_dev && memory_out_dev && B > 0 && Ls > 0, "bad arguments");
    TTB_CHECK(Ls <= e->d.max_positions, "source longer than the positional table");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return TTB_DISPATCH(e, encode_api<float>(e, src_dev, src_pad_mask_dev, B, Ls, memory_out_dev, s),
                        encode_api<__nv_bfloat16>(e, src_dev, src_pad_mask_dev, B, Ls, memory_out_dev, s));
}

int ttb_decode_tgt(ttb_engine* e, const int64_t* tgt_dev, int32_t B, int32_t Lt, const float* memory_dev,
                   const uint8_t* memory_pad_mask_dev, int32_t Ls, float* logits_out_dev, void* stream) {
    TTB_CHECK(e && e->finalized, "engine not finalized");
    TTB_CHECK(tgt_dev && memory_dev && memory_pad_mask_dev && logits_out_dev && B > 0 && Lt > 0 && Ls > 0, "bad arguments");
    TTB_CHECK(Lt <= e->d.max_positions, "target longer than the positional table");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return TTB_DISPATCH(e, decode_api<float>(e, tgt_dev, B, Lt, memory_dev, memory_pad_mask_dev, Ls, logits_out_dev, s),
                        decode_api<__nv_bfloat16>(e, tgt_dev, B, Lt, memory_dev, memory_pad_mask_dev, Ls, logits_out_dev, s));
}

int ttb_greedy_speculative_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls,
                                    int32_t max_len, int32_t draft_len, int32_t n_drafts, int32_t pad_token,
                                    int32_t bos_token, int32_t eos_token, int32_t replace_token,
                                    int32_t tie_break, int64_t* out_dev, int32_t* trace_dev,
                                    ttb_generate_stats* stats, void* stream) {
    TTB_CHECK(e && e->finalized, "engine not finalized");
    TTB_CHECK(src_dev && out_dev && B > 0 && Ls > 1 && max_len > 1, "bad arguments");
    TTB_CHECK(n_drafts > 0, "The number of drafts must be greater than 0");
    TTB_CHECK(pad_token != replace_token, "The pad token and the replace token must be different");
    TTB_CHECK(eos_token != replace_token, "The eos token and the replace token must be different");
    TTB_CHECK(eos_token != pad_token, "The eos token and the pad token must be different");
    TTB_CHECK(Ls <= e->d.max_positions && max_len + draft_len + 2 <= e->d.max_positions, "sequence longer than the positional table");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return drained(e, TTB_DISPATCH(e, greedy_api<float>(e, src_dev, B, Ls, max_len, draft_len, n_drafts, pad_token, bos_token, eos_token,
                                             replace_token, tie_break, out_dev, trace_dev, stats, s),
                        greedy_api<__nv_bfloat16>(e, src_dev, B, Ls, max_len, draft_len, n_drafts, pad_token, bos_token, eos_token,
                                                  replace_token, tie_break, out_dev, trace_dev, stats, s)));
}

int ttb_greedy_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls, int32_t max_len, int32_t pad_token,
                        int32_t bos_token, int32_t eos_token, int64_t* out_dev, ttb_generate_stats* stats, void* stream) {
    TTB_CHECK(e && e->finalized, "engine not finalized");
    TTB_CHECK(src_dev && out_dev && B > 0 && Ls > 0 && max_len > 1, "bad arguments");
    TTB_CHECK(Ls <= e->d.max_positions && max_len + 2 <= e->d.max_positions, "sequence longer than the positional table");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return drained(e, TTB_DISPATCH(e, greedy_api<float>(e, src_dev, B, Ls, max_len, 0, 1, pad_token, bos_token, eos_token, -1, 1, out_dev, nullptr, stats, s, true),
                        greedy_api<__nv_bfloat16>(e, src_dev, B, Ls, max_len, 0, 1, pad_token, bos_token, eos_token, -1, 1, out_dev, nullptr, stats, s, true)));
}

int ttb_beam_search_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls, int32_t max_len, int32_t beam_size,
                             int32_t pad_token, int32_t bos_token, int32_t eos_token, int64_t* out_dev, int32_t* out_width,
                             ttb_generate_stats* stats, void* stream) {
    TTB_CHECK(e && e->finalized, "engine not finalized");
    TTB_CHECK(src_dev && out_dev && B > 0 && Ls > 0, "bad arguments");
    TTB_CHECK(max_len > 1 && beam_size > 0, "max_len must be greater than 1 and beam_size greater than 0");
    TTB_CHECK(beam_size <= e->d.tgt_vocab_size, "beam_size larger than the vocabulary (the reference's first topk fails as well)");
    TTB_CHECK(Ls <= e->d.max_positions && max_len + 2 <= e->d.max_positions, "sequence longer than the positional table");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return drained(e, TTB_DISPATCH(e, std_beam_api<float>(e, src_dev, B, Ls, max_len, beam_size, pad_token, bos_token, eos_token, out_dev, out_width, stats, s),
                        std_beam_api<__nv_bfloat16>(e, src_dev, B, Ls, max_len, beam_size, pad_token, bos_token, eos_token, out_dev, out_width, stats, s)));
}

int ttb_beam_speculative_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls, int32_t max_len,
                                  int32_t n_best, int32_t draft_len, int32_t n_drafts, int32_t smart_drafts_mode, int32_t pad_token,
                                  int32_t bos_token, int32_t eos_token, int32_t c_token, int32_t tie_break, int64_t* out_dev,
                                  int32_t* out_width, int32_t* trace_nacc_dev, int32_t* trace_pick_dev, ttb_generate_stats* stats,
                                  void* stream) {
    TTB_CHECK(e && e->finalized, "engine not finalized");
    TTB_CHECK(src_dev && out_dev && B > 0 && Ls > 1 && max_len > 2, "bad arguments");
    TTB_CHECK(n_best >= 1 && n_best <= 32, "n_best must be in [1, 32]");
    TTB_CHECK(n_drafts > 0, "The number of drafts must be greater than 0");
    TTB_CHECK(pad_token != c_token, "The pad token and the replace token must be different");
    TTB_CHECK(eos_token != c_token, "The eos token and the replace token must be different");
    TTB_CHECK(eos_token != pad_token, "The eos token and the pad token must be different");
    TTB_CHECK(Ls <= e->d.max_positions && max_len + 210 <= e->d.max_positions, "sequence longer than the positional table");
    TTB_CHECK(e->d.tgt_vocab_size <= 1024, "vocabularies above 1024 tokens are not supported by the beam statistics kernel");
    TTB_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return drained(e, TTB_DISPATCH(e, beam_api<float>(e, src_dev, B, Ls, max_len, n_best, draft_len, n_drafts, pad_token, bos_token, eos_token,
                                           c_token, tie_break, out_dev, out_width, trace_nacc_dev, trace_pick_dev, stats, s,
                                           smart_drafts_mode != 0),
                        beam_api<__nv_bfloat16>(e, src_dev, B, Ls, max_len, n_best, draft_len, n_drafts, pad_token, bos_token,
                                                eos_token, c_token, tie_break, out_dev, out_width, trace_nacc_dev, trace_pick_dev, stats, s,
                                                smart_drafts_mode != 0)));
}

int ttb_kernel_class_count(void) { return KC_COUNT; }
const char* ttb_kernel_class_name(int32_t id) { return (id >= 0 && id < KC_COUNT) ? kKcNames[id] : ""; }

int ttb_engine_set_profiling(ttb_engine* e, uint32_t class_mask) {
    TTB_CHECK(e, "null engine");
    e->prof.mask = class_mask;
    for (int i = 0; i < KC_COUNT; ++i) { e->prof.ms[i] = 0.0; e->prof.n[i] = 0; }
    return 0;
}

int ttb_engine_get_profile(ttb_engine* e, int32_t n_classes, double* ms_out, int64_t* launches_out) {
    TTB_CHECK(e && ms_out && launches_out, "null argument");
    for (int i = 0; i < n_classes && i < KC_COUNT; ++i) { ms_out[i] = e->prof.ms[i]; launches_out[i] = e->prof.n[i]; }
    return 0;
}

int ttb_engine_get_history(ttb_engine* e, int32_t* live_queries_out, int32_t capacity) {
    TTB_CHECK(e && live_queries_out, "null argument");
    const int n = (int)e->h_hist.size();
    for (int i = 0; i < n && i < capacity; ++i) live_queries_out[i] = e->h_hist[i];
    return n;
}

int ttb_gemm(int32_t precision, const void* A_dev, const void* W_dev, const float* bias_dev, float* C_dev,
             int32_t M, int32_t N, int32_t K, int32_t relu, void* stream) {
    TTB_CHECK(A_dev && W_dev && C_dev && M > 0 && N > 0 && K > 0, "bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (precision == TTB_PRECISION_FP32) {
        TTB_CHECK(K % 16 == 0, "K must be a multiple of 16");
        launch_gemm_f32<float>(static_cast<const float*>(A_dev), K, static_cast<const float*>(W_dev), bias_dev, C_dev, N,
                               RowCount(M), N, K, relu != 0, s);
    } else {
        TTB_CHECK(K % 64 == 0, "K must be a multiple of 64");
        if (int rc = launch_gemm_bf16_tc<float>(static_cast<const __nv_bfloat16*>(A_dev), K, static_cast<const __nv_bfloat16*>(W_dev),
                                                bias_dev, C_dev, N, RowCount(M), N, K, relu != 0, s))
            return rc;
    }
    TTB_CUDA_OK(cudaGetLastError());
    return 0;
}

int ttb_gemm_bf16_out(const void* A_dev, const void* W_dev, const float* bias_dev, void* C_dev, int32_t M, int32_t N, int32_t K,
                      int32_t relu, void* stream) {
    TTB_CHECK(A_dev && W_dev && C_dev && M > 0 && N > 0 && K > 0, "bad arguments");
    TTB_CHECK(K % 64 == 0, "K must be a multiple of 64");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (int rc = launch_gemm_bf16_tc<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(A_dev), K, static_cast<const __nv_bfloat16*>(W_dev), bias_dev,
                                                    static_cast<__nv_bfloat16*>(C_dev), N, RowCount(M), N, K, relu != 0, s))
        return rc;
    TTB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
