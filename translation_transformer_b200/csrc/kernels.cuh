// Kernel launchers of libttb200.  Every launcher enqueues on `stream` and never synchronises.
//
// Row counts that are only known on the device (the number of live queries shrinks while
// decoding) are passed as `RowCount{max_rows, units_dev, rows_per_unit}`: the grid is sized for
// `max_rows`, each block reads `*units_dev * rows_per_unit` and exits early beyond it.
#pragma once
#include "common.cuh"

namespace ttb {

struct RowCount {
    int max_rows;            // host-side upper bound (grid sizing)
    const int* units_dev;    // nullptr -> all max_rows rows are live
    int rows_per_unit;
    __host__ __device__ RowCount(int m = 0, const int* u = nullptr, int r = 1) : max_rows(m), units_dev(u), rows_per_unit(r) {}
    __device__ __forceinline__ int live() const { return units_dev ? (*units_dev) * rows_per_unit : max_rows; }
};

// ---- elementwise.cu ----------------------------------------------------------------------
void launch_i64_to_i32(const long long* in, int* out, long long n, cudaStream_t s);
void launch_i32_to_i64(const int* in, long long* out, long long n, cudaStream_t s);
void launch_mask_to_tokens(const unsigned char* mask, int* out, long long n, int pad_id, cudaStream_t s);
// out[b] = 1 + index of the last token of row b that differs from pad_id (0 for an all-pad row)
void launch_row_lengths(const int* tok, int B, int L, int pad_id, int* out, cudaStream_t s);
// x[t] = table[tok[t]] + pe[(t % L) + 1]; also writes the low-precision copy when xh != nullptr
template <typename ActT>
void launch_embed_seq(const int* tok, int T, int L, const float* table, const float* pe, int E,
                      float* x, ActT* xh, cudaStream_t s);
// out = LN(resid + y) * g + b ; optional second LN (final norm of the stack) applied on top
template <typename ActT>
void launch_add_layernorm(const float* resid, const float* y, const float* g, const float* b,
                          const float* g2, const float* b2, float* out, ActT* outh,
                          RowCount rows, int E, cudaStream_t s);
void launch_argmax_rows(const float* logits, int ld, int V, int* out, RowCount rows, cudaStream_t s);

// ---- gemm_simt.cu : C[M,N] = A[M,K] * W[N,K]^T + bias (fp32 FMA, exact-precision path) -------
template <typename OutT>
void launch_gemm_f32(const float* A, int lda, const float* W, const float* bias, OutT* C, int ldc,
                     RowCount rows, int N, int K, bool relu, cudaStream_t s);

// ---- gemm_tcgen05.cu : same contract, bf16 operands, fp32 accumulation in TMEM (tcgen05.mma) ----
// A is [M,K] bf16 row-major (lda elements), W is [N,K] bf16 row-major; both are TMA-loaded.
template <typename OutT>
int launch_gemm_bf16_tc(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias,
                        OutT* C, int ldc, RowCount rows, int N, int K, bool relu, cudaStream_t s);

// Fused sub-layer tail for embedding_dim 256 (bf16 path):  x <- LN2?(LN1(x + A W^T + bias)), x fp32 updated in
// place, xh = bf16 copy.  W is [256, K] bf16; g2/b2 = nullptr without the second LayerNorm.  Optional chained
// projection of the result (K == 256 only): q2 = xh W2^T + bias2 with W2 [256, 256] bf16 (the cross-attention query).
int launch_gemm_resid_ln(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, float* x, __nv_bfloat16* xh,
                         const float* g1, const float* b1, const float* g2, const float* b2, RowCount rows, int K, cudaStream_t s,
                         const __nv_bfloat16* W2 = nullptr, const float* bias2 = nullptr, __nv_bfloat16* q2 = nullptr);

// Vocabulary projection fused with arg-max (greedy loop, bf16 path): pred[row] = argmax_v(A[row] . W[v] + bias[v]).
// Returns -1 when the shape does not fit the kernel (K % 64, V <= 512, shared memory) so the caller can fall back.
int launch_classifier_argmax(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, int* pred, RowCount rows,
                             int V, int K, cudaStream_t s);

// Vocabulary projection fused with the statistics of the speculative beam search (bf16 path, n_best <= 16): per row the soft-max
// maximum / sum, the n_best largest logits (descending, ties: lower id) and their ids, the nucleus-truncated support size
// and the logit of token row_tok[row] -- everything beam_choose / beam_expand read; the logits never leave tensor memory.
// Returns -1 when the shape does not fit (caller: logits GEMM + launch_beam_stats).
int launch_classifier_stats(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, RowCount rows, int V, int Kdim, int n_best,
                            const int* row_tok, float* tokv, float* lmax, float* lsum, int* nkeep, float* topv, int* topi, cudaStream_t s);

// Fused feed-forward sub-layer for embedding_dim 256 (bf16 path):
//   x <- LN2?(LN1(x + relu(xh W1^T + b1) W2^T + b2)), x fp32 and its bf16 copy xh both updated in place.
// Chained form (att, Wo given; only with the CTA-pair kernel, see ffn_pair_available): the same launch first computes the
// preceding sub-layer tail x <- LN0(x + att Wo^T + bias_o) (the cross-attention out-projection), i.e. xh is not read.
int launch_ffn_fused(__nv_bfloat16* xh, const __nv_bfloat16* W1, const float* bias1, const __nv_bfloat16* W2, const float* bias2,
                     float* x, const float* g1, const float* b1, const float* g2, const float* b2, RowCount rows, int F, cudaStream_t s,
                     const __nv_bfloat16* att = nullptr, const __nv_bfloat16* Wo = nullptr, const float* bias_o = nullptr,
                     const float* g0 = nullptr, const float* b0 = nullptr);
bool ffn_pair_available(int max_rows);   // the cta_group::2 feed-forward kernel can run all its clusters co-resident

// ---- attention.cu ---------------------------------------------------------------------------
// Generic masked attention over "groups".  Queries of group g are tokens g*Lq .. g*Lq+Lq-1 of
// `q`; its keys/values are rows kv_row0 .. kv_row0+Lk-1 with kv_row0 = (kvmap ? kvmap[g] : g) *
// kv_group_stride.  Key j is masked when key_tok[(kvmap?kvmap[g]:g)*key_tok_stride + j] == pad_id,
// and, when `causal`, when j > i.
template <typename ActT>
void launch_attention(const ActT* q, int q_ld, const ActT* k, const ActT* v, int kv_ld,
                      ActT* out, int out_ld, int n_groups_max, const int* n_groups_dev,
                      int Lq, int Lk, long long kv_group_stride, const int* kvmap,
                      const int* key_tok, int key_tok_stride, int pad_id, bool causal,
                      int heads, int head_dim, cudaStream_t s, const int* lk_dev = nullptr);

// Speculative self-attention: group g = live query slot; b = active[g]; f = front[b].
// Queries: rows (g*N + n)*(D+1) + i of `qkv` (q | k | v packed, row stride qkv_ld).
// Keys: cache positions 0..f-1 of query b (masked where gen[b][j] == pad) followed by the new
// keys i' <= i of the same draft row (i' = 0 is masked when gen[b][f] == pad).
template <typename ActT>
void launch_spec_self_attention(const ActT* qkv, int qkv_ld, const ActT* kcache, const ActT* vcache,
                                long long cache_query_stride, int cache_ld, ActT* out, int out_ld,
                                int B_max, const int* n_active_dev, const int* active, const int* front,
                                const int* gen, int gen_ld, int pad_id, int N, int D,
                                int heads, int head_dim, int max_cache_len, cudaStream_t s);

// ---- attention_mma.cu : tensor-core (mma.sync bf16) versions of the two entry points above ----------
void launch_attention_mma(const __nv_bfloat16* q, int q_ld, const __nv_bfloat16* k, const __nv_bfloat16* v, int kv_ld,
                          __nv_bfloat16* out, int out_ld, int n_groups_max, const int* n_groups_dev,
                          int Lq, int Lk, long long kv_group_stride, const int* kvmap,
                          const int* key_tok, int key_tok_stride, int pad_id, bool causal,
                          int heads, int head_dim, cudaStream_t s, const int* lk_dev = nullptr, const int* lk_group = nullptr,
                          const int4* desc = nullptr);
void launch_spec_self_attention_mma(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* kcache, const __nv_bfloat16* vcache,
                                    long long cache_query_stride, int cache_ld, __nv_bfloat16* out, int out_ld,
                                    int B_max, const int* n_active_dev, const int* active, const int* front,
                                    const int* gen, int gen_ld, int pad_id, int N, int D,
                                    int heads, int head_dim, cudaStream_t s, const int4* desc = nullptr);

// ---- attention_tc.cu : tcgen05 attention of the decoding loops (head_dim 32, S in tensor memory, V caches kept transposed) ----
bool attention_tc_supported(int head_dim, int Lq, int max_shared_keys, int row_len);
// out[(g * E + c) * pitch + j] = in[(g * L + j) * ld + c]
void launch_transpose_v(const __nv_bfloat16* in, int ld, int groups, int L, int E, __nv_bfloat16* out, int pitch, cudaStream_t s);
int launch_cross_attention_tc(const __nv_bfloat16* q, int q_ld, const __nv_bfloat16* k, int k_ld, long long k_group_stride,
                              const __nv_bfloat16* vt, long long vt_group_stride, int vt_pitch, __nv_bfloat16* out, int out_ld,
                              int n_groups_max, const int* n_groups_dev, int Lq, const int* key_tok, int key_tok_stride, int pad_id,
                              const int* lk_dev, int heads, const int4* desc, cudaStream_t s);
int launch_spec_self_attention_tc(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* kcache, long long k_query_stride, int k_ld,
                                  const __nv_bfloat16* vtcache, long long vt_query_stride, int vt_pitch, __nv_bfloat16* out, int out_ld,
                                  int B_max, const int* n_active_dev, const int* gen, int gen_ld, int pad_id, int N, int D, int heads,
                                  const int4* desc, cudaStream_t s);

// ---- drafting.cu ------------------------------------------------------------------------------
// Mirrors utils/drafting.py::make_drafts on device; src is (B, L) int32 with row stride src_ld (the
// caller skips the BOS column by passing src + 1, L - 1).  out is (B, N, Deff) int32.
void launch_make_drafts(const int* src, int src_ld, int B, int L, int Deff, int N, int eos, int pad, int replace,
                        int* out, cudaStream_t s);

// ---- greedy.cu -------------------------------------------------------------------------------
struct GreedyState {
    int B, N, D, max_len, gen_ld, pad, bos, eos, Ls;
    int* gen;          // [B][gen_ld] generated tokens (row b of the reference's token matrix)
    int* front;        // [B] index of the last generated token
    int* active;       // [B] compact list of live query ids (order preserved, like boolean masking)
    int* ctrl;         // [CTRL_COUNT]
    const int* drafts; // [B][N][D]
    int* pred;         // [B*N*(D+1)] argmax predictions of the current iteration
    long long* out;    // [B][max_len] finished predictions (int64 like the reference)
    int* sel;          // [B][4] per pre-retirement slot: {query id, old front, draft index, n_accepted}
    int* trace;        // optional [max_len][B][4] = {query id, n_accepted, draft index, width}
    int tie_break;     // 0 = torch-CPU topk(1) emulation, 1 = lowest index
    int* hist;         // [max_len + 2] live queries at the start of every iteration
    int4* desc;        // [B] per live slot of the COMING iteration: {query id, front, token at front, source length};
                       // written by the init / accept kernels so that the kernels of an iteration need one load
                       // instead of the chain active[g] -> front[b] -> gen[b][front]
    const int* src_len;  // [B] source length per query (cross-attention key bound), may be nullptr
};
void launch_greedy_init(const GreedyState& st, cudaStream_t s);
// First kernel of an iteration: appends the K/V rows of the tokens accepted in the previous iteration (chosen
// draft recorded in st.sel) to the self-attention cache of every layer, and embeds the (D+1) step tokens of
// every live draft row.
template <typename ActT>
void launch_greedy_advance(const GreedyState& st, const float* table, const float* pe, int E, float* x, ActT* xh,
                           const ActT* qkv_all, long long qkv_layer_stride, int n_layers, int qkv_ld, ActT* kcache, ActT* vcache,
                           long long cache_layer_stride, long long cache_query_stride, int cache_ld, cudaStream_t s,
                           long long vt_layer_stride = 0, long long vt_query_stride = 0, int vt_pitch = 0);
// vt_pitch > 0: `vcache` is the TRANSPOSED value cache [layer][query][E][vt_pitch] of the tcgen05 attention kernel
// picks the best draft per live query, appends tokens, retires finished queries, plans next width
void launch_greedy_accept(const GreedyState& st, cudaStream_t s);
// standard greedy decoding (no drafts: N = 1, D = 0): appends the predicted token of every row, stop test
void launch_greedy_std_step(const GreedyState& st, cudaStream_t s);

// ---- beam.cu --------------------------------------------------------------------------------------
struct BeamState {
    int B, K, N, dl0, V, pad, bos, eos, ldw, tie_break;
    int max_len, w0;                      // length budget of the loop; width of the token matrix in the first iteration
    int* cand_cur; int* cand_next;        // [B*K][ldw] hypotheses (query-major), ping-pong
    float* logp_cur; float* logp_next;    // [B*K]
    const int* drafts;                    // [B][N][dl0]
    int* c_slot0; int* c_fin; int* c_rowbase; int* c_nacc; int* c_pick; int* acc_stat; int* ctrl;
    // per live attention group (written by beam_prepare): {candidate, front, token at front, -} for the self-attention
    // over the candidate caches and {query, -, -, source length} for the cross-attention: one 16-byte load per CTA instead
    // of the chain live list -> front -> token, and everything the attention kernels may read ahead of their dependency wait
    int4* desc_self; int4* desc_cross; const int* src_len;
    int* host_ctrl;   // pinned host mirror of ctrl (device-accessible): BC_COUNT words + a sequence word written last
    int* rows_tok; int* row_cand; int* row_query; int* row_slot0;      // live decoder rows
    float* topv; int* topi; int* nkeep; float* lmax; float* lsum;      // per (row, position) statistics
    // KV-cached pass: token whose logit the accepted-path sums need at (row, position) = the row's next draft token (-1
    // behind the last one), written by the embedding kernel; its logit, written by the statistics kernel
    int* row_tok; float* tokv;
    int* trace_nacc; int* trace_pick;     // optional [iter][B*K][N] / [iter][B*K]
    // KV-cached decoder pass: live candidates as attention groups, cache bookkeeping of the new candidates
    int* live_cand; int* live_query;      // [live candidates] candidate index / query index
    int* c_front;                         // [B*K] cached positions of a candidate = first free slot - 1
    int* n_parent; int* n_keep; int* n_row;   // [B*K] new candidate: parent (-1: finished parent), accepted draft tokens, decoder row
    // smart_drafts_mode (speculative_decoding.py:600-845): `drafts` is the window library [B][n_lib][dl0] whose first
    // token is the key; a candidate tries the windows that start with its last token
    int smart, n_lib;
    int* tok_cnt; int* tok_list;          // [B][V] number of windows per first token (1..N), [B][V][N] their library indices
    int* c_cnt; int* c_last;              // [B*K] rows (drafts) of a candidate, its last token
    int* row_draft;                       // [rows] library index of the row's draft (row_cand / row_query: its candidate / query)
};
void launch_beam_build_lib(const BeamState& st, cudaStream_t s);
void launch_beam_init(const BeamState& st, cudaStream_t s);
void launch_beam_prepare(const BeamState& st, int C, int beam, int dl, cudaStream_t s);
void launch_beam_fill_rows(const BeamState& st, int C, int beam, int W, int dl, cudaStream_t s);
template <typename ActT>
void launch_beam_gather(const BeamState& st, const float* x, const ActT* xh, int max_rows, int W, int dl, int E,
                        float* xg, ActT* xgh, cudaStream_t s);
// KV-cached pass: embeds (last token, draft tokens) of every live (candidate, draft) row at positions front .. front+dl
template <typename ActT>
void launch_beam_embed_cached(const BeamState& st, int beam, int max_rows, int dl, const float* table, const float* pe, int E,
                              float* x, ActT* xh, cudaStream_t s);
// new candidate c' <- cache of its parent [0, front) + K/V of positions front .. front + n_keep of the chosen draft row
template <typename ActT>
void launch_beam_cache_update(const BeamState& st, int dl, const ActT* qkv_all, long long qkv_layer_stride, int n_layers, int qkv_ld,
                              int E, const ActT* kc_cur, const ActT* vc_cur, ActT* kc_next, ActT* vc_next, long long cache_layer_stride,
                              long long cache_cand_stride, cudaStream_t s);
void launch_beam_stats(const BeamState& st, const float* logits, int max_rows, int dl, cudaStream_t s);
void launch_beam_choose(const BeamState& st, int C, int beam, int dl, cudaStream_t s);
void launch_beam_expand(const BeamState& st, int beam, int dl, const float* logits, cudaStream_t s);
void launch_beam_export(const int* cand, int ldw, int R, int W, long long* out, cudaStream_t s);
// ---- std_beam.cu : standard beam search (standard_decoding.py:90-174) -------------------------------------------
struct StdBeamState {
    int B, K, V, pad, bos, eos, ldw;
    int* y_cur; int* y_next;              // [B*K][ldw] hypotheses (query-major), ping-pong
    float* score_cur; float* score_next;  // [B*K]
    int* fin; int* cand_row;              // [B*K] contains EOS / live-row index or -1
    int* row_cand; int* row_query;        // [live rows]
    int* rows_tok;                        // [live rows][W] decoder input
    float* total;                         // [B*K][V] score + log-softmax
    int* ctrl;                            // [0] live rows, [1] hypotheses with EOS after the selection
    // KV-cached pass (one new token per live hypothesis and step; nullptr: full-prefix recomputation)
    int* c_front;                         // [B*K] cached positions of a hypothesis (= W - 1 for all of them)
    int* parent;                          // [B*K] hypothesis of the previous step a new hypothesis continues
    int4* desc_self; int4* desc_cross;    // per live row: {hypothesis, front, token at front, -} / {query, -, -, source length}
    const int* src_len;
};
void launch_sbeam_init(const StdBeamState& st, cudaStream_t s);
void launch_sbeam_prepare(const StdBeamState& st, int C, int beam, int W, cudaStream_t s);
template <typename ActT>
void launch_sbeam_gather_last(const StdBeamState& st, const float* x, const ActT* xh, int max_rows, int W, int E, float* xg, ActT* xgh,
                              cudaStream_t s);
void launch_sbeam_scores(const StdBeamState& st, int C, const float* logits, cudaStream_t s);
int launch_sbeam_select(const StdBeamState& st, int beam, int W, cudaStream_t s);
// KV-cached pass: embedding of the last token (position W - 1) of every live hypothesis
template <typename ActT>
void launch_sbeam_embed_last(const StdBeamState& st, int max_rows, int W, const float* table, const float* pe, int E, float* x, ActT* xh,
                             cudaStream_t s);
// cache of a new hypothesis = cache of its parent [0, W - 1) + K/V of the parent's row of this step (position W - 1)
template <typename ActT>
void launch_sbeam_cache_update(const StdBeamState& st, int W, const ActT* qkv_all, long long qkv_layer_stride, int n_layers, int qkv_ld, int E,
                               const ActT* kc_cur, const ActT* vc_cur, ActT* kc_next, ActT* vc_next, long long cache_layer_stride,
                               long long cache_cand_stride, cudaStream_t s);

// embedding of (rows, L) token matrices whose live row count is on the device
template <typename ActT>
void launch_embed_seq_rows(const int* tok, RowCount rows, int L, const float* table, const float* pe, int E,
                           float* x, ActT* xh, cudaStream_t s);

}  // namespace ttb
