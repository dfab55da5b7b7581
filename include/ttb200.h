/* libttb200 — C ABI of the B200-native translation-transformer inference hot path.
 *
 * Every entry point replaces one Python-level interface of Academich/translation-transformer
 * (the reference is pure Python/PyTorch; the file:line given with each function is the call a
 * reference maintainer would re-point, see INTEGRATION.md for the ctypes binding).
 *
 * Conventions
 *   - plain pointers and sizes only; no framework types.  `*_dev` pointers are device pointers of
 *     the engine's CUDA device, `stream` is a `cudaStream_t` passed as `void*` (NULL = default).
 *   - token ids are int64 at the boundary (torch.LongTensor in the reference).
 *   - every function returns 0 on success; otherwise `ttb_last_error()` describes the failure.
 *   - return code 10..19 = the reference itself would have raised at this point
 *     (TTB_ERR_REF_*), the Python wrappers re-raise the same exception type.
 *   - there is no CPU implementation behind this ABI: without a CUDA device every call fails.
 */
#ifndef TTB200_H
#define TTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTB_ABI_VERSION 2

#define TTB_PRECISION_FP32 0 /* fp32 FMA GEMMs, exact-parity path (1e-5 logits)            */
#define TTB_PRECISION_BF16 1 /* tcgen05 bf16 GEMMs, fp32 accumulate (1e-2 logits)          */

#define TTB_ERR_REF_INDEX 10 /* reference: scatter index out of bounds (speculative_decoding.py:111) */
#define TTB_ERR_REF_SHAPE 11 /* reference: finished row wider than max_len (speculative_decoding.py:158) */
#define TTB_ERR_REF_ASSERT 12 /* reference: assert in topk_in_each_group (speculative_decoding.py:195) */

typedef struct ttb_engine ttb_engine;

/* Architecture of VanillaTransformer (src/model/modules.py:12-37). */
typedef struct {
    int32_t src_vocab_size, tgt_vocab_size;
    int32_t embedding_dim, feedforward_dim;
    int32_t num_encoder_layers, num_decoder_layers, num_heads;
    int32_t src_pad_token_idx, tgt_pad_token_idx;
    int32_t precision;      /* TTB_PRECISION_* */
    int32_t max_positions;  /* rows of the positional table minus one (the reference uses 5000) */
} ttb_model_desc;

/* Statistics of one generate() call (model_calls_num etc. of the reference generators). */
typedef struct {
    int32_t model_calls;      /* decoder invocations (speculative_decoding.py:121, :496, :742) */
    int32_t accepted_tokens;  /* draft tokens accepted                                         */
    int32_t produced_tokens;  /* accepted + bonus tokens                                       */
    int32_t unfinished;       /* queries still alive when the loop stopped                     */
    int32_t error;            /* 0, TTB_ERR_REF_INDEX or TTB_ERR_REF_SHAPE                     */
    int32_t gpu_launches;     /* kernels launched by this call                                 */
    float   gpu_ms;           /* device time of the call measured with CUDA events             */
    float   reserved;
} ttb_generate_stats;

int ttb_abi_version(void);
const char* ttb_last_error(void);
/* 0 if a usable sm_100 device is present */
int ttb_device_check(int device);

/* ---- engine lifecycle: replaces VanillaTransformer.__init__ + load_state_dict (modules.py:10-83) */
int ttb_engine_create(const ttb_model_desc* desc, int device, ttb_engine** out);
void ttb_engine_destroy(ttb_engine* e);
/* `name` is a key of VanillaTransformer.state_dict() (optionally prefixed "model.") or
 * "positional_encoding.pe"; `data` may be a host or a device pointer to `numel` floats. */
int ttb_engine_set_param(ttb_engine* e, const char* name, const float* data, int64_t numel);
int ttb_engine_finalize(ttb_engine* e);

/* ---- utils/drafting.py:5 make_drafts(src, draft_len, n_drafts, min_draft_len, max_draft_len,
 *      eos_token_idx, pad_token_idx, replace_token_idx) -> (B, N, D) int64
 * src_dev is (B, L) int64 with row stride src_ld.  Returns D (clamped draft length) in *d_out. */
int ttb_make_drafts(const int64_t* src_dev, int64_t src_ld, int32_t B, int32_t L, int32_t draft_len,
                    int32_t n_drafts, int32_t min_draft_len, int32_t max_draft_len, int32_t eos,
                    int32_t pad, int32_t replace, int64_t* out_dev, int32_t* d_out, void* stream);

/* ---- modules.py:108 encode_src(src, src_pad_mask) -> memory (B, Ls, E) fp32
 * src_pad_mask_dev may be NULL (then src == src_pad_token_idx, as every caller does). */
int ttb_encode_src(ttb_engine* e, const int64_t* src_dev, const uint8_t* src_pad_mask_dev, int32_t B,
                   int32_t Ls, float* memory_out_dev, void* stream);

/* ---- modules.py:117 decode_tgt(tgt, memory, memory_pad_mask) -> logits (B, Lt, V) fp32 */
int ttb_decode_tgt(ttb_engine* e, const int64_t* tgt_dev, int32_t B, int32_t Lt, const float* memory_dev,
                   const uint8_t* memory_pad_mask_dev, int32_t Ls, float* logits_out_dev, void* stream);

/* ---- speculative_decoding.py:39 TranslationInferenceGreedySpeculative.generate(src)
 * Whole decoding loop on the device (encoder, drafts, KV-cached decoder steps, verification,
 * retirement).  out_dev is (B, max_len) int64 (the reference returns it with a singleton middle
 * dimension).  trace_dev is optional: (max_len, B, 4) int32 = {query, n_accepted, draft, width}
 * per iteration and live slot.  tie_break: 0 = torch-CPU topk(1) order, 1 = lowest index. */
int ttb_greedy_speculative_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls,
                                    int32_t max_len, int32_t draft_len, int32_t n_drafts, int32_t pad_token,
                                    int32_t bos_token, int32_t eos_token, int32_t replace_token,
                                    int32_t tie_break, int64_t* out_dev, int32_t* trace_dev,
                                    ttb_generate_stats* stats, void* stream);

/* ---- standard_decoding.py:29 TranslationInferenceGreedy.generate(src)
 * Plain greedy decoding, one token per row and step, on the same KV-cached device loop (no drafts).  Like the
 * reference, rows keep decoding after their EOS until every row predicts EOS or PAD in the same step.
 * out_dev is (B, max_len) int64, PAD behind the last generated column. */
int ttb_greedy_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls, int32_t max_len, int32_t pad_token,
                        int32_t bos_token, int32_t eos_token, int64_t* out_dev, ttb_generate_stats* stats, void* stream);

/* ---- standard_decoding.py:90 TranslationInferenceBeamSearch.generate(src)
 * Standard beam search: one decoder call per generated column on the hypotheses without EOS; scores are
 * log(softmax(logits)) sums, finished hypotheses continue with PAD (artificial logits), the beam_size best of the
 * beam x vocabulary continuations survive (sorted).  out_dev must hold B * beam_size * max_len int64; the
 * hypotheses are written densely as (B, beam_size, *out_width). */
int ttb_beam_search_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls, int32_t max_len, int32_t beam_size,
                             int32_t pad_token, int32_t bos_token, int32_t eos_token, int64_t* out_dev, int32_t* out_width,
                             ttb_generate_stats* stats, void* stream);

/* ---- speculative_decoding.py:422 TranslationInferenceBeamSearchSpeculative.generate(src)
 * smart_drafts_mode = 0: "try all the drafts" (:428-598), every candidate tries the n_drafts source windows;
 * smart_drafts_mode = 1: :600-845, a candidate tries the windows of a (Ls - 5)-window library whose first token equals
 * its last token (at most n_drafts of them).  out_dev must hold B * n_best * (max_len + clamp(draft_len,5,200) + 4)
 * int64; the hypotheses are written densely as (B, n_best, *out_width).  Optional traces: trace_nacc_dev
 * (max_len, B*n_best, n_drafts) accepted length of every draft (-1 behind a candidate's last draft), trace_pick_dev
 * (max_len, B*n_best) chosen draft, per iteration. */
int ttb_beam_speculative_generate(ttb_engine* e, const int64_t* src_dev, int32_t B, int32_t Ls, int32_t max_len,
                                  int32_t n_best, int32_t draft_len, int32_t n_drafts, int32_t smart_drafts_mode, int32_t pad_token,
                                  int32_t bos_token, int32_t eos_token, int32_t c_token, int32_t tie_break, int64_t* out_dev,
                                  int32_t* out_width, int32_t* trace_nacc_dev, int32_t* trace_pick_dev, ttb_generate_stats* stats,
                                  void* stream);

/* ---- instrumentation (bench.py): per-kernel-class CUDA-event timing on the launching stream.
 * class_mask bit i enables class i (names via ttb_kernel_class_name); totals accumulate over
 * generate() calls until the next ttb_engine_set_profiling. */
int ttb_kernel_class_count(void);
const char* ttb_kernel_class_name(int32_t id);
int ttb_engine_set_profiling(ttb_engine* e, uint32_t class_mask);
int ttb_engine_get_profile(ttb_engine* e, int32_t n_classes, double* ms_out, int64_t* launches_out);
/* live queries at the start of every decoder iteration of the last generate(); returns their number */
int ttb_engine_get_history(ttb_engine* e, int32_t* live_queries_out, int32_t capacity);

/* Standalone GEMM entry used by the kernel unit tests and the roofline microbenchmark:
 * C[M,N] = A[M,K] * W[N,K]^T + bias (+ReLU).  precision selects the kernel; A/W are fp32 for
 * TTB_PRECISION_FP32 and bf16 (uint16 storage) for TTB_PRECISION_BF16; C is fp32. */
int ttb_gemm(int32_t precision, const void* A_dev, const void* W_dev, const float* bias_dev, float* C_dev,
             int32_t M, int32_t N, int32_t K, int32_t relu, void* stream);

/* Same with a bf16 result (the form the decoder uses for the QKV / cross K-V projections: wide K = 256 projections run
 * on the CTA-pair kernel). */
int ttb_gemm_bf16_out(const void* A_dev, const void* W_dev, const float* bias_dev, void* C_dev, int32_t M, int32_t N, int32_t K,
                      int32_t relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTB200_H */
