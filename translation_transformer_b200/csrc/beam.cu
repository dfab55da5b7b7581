// Device side of the speculative beam search ("try all the drafts" mode).
// Reference: /root/reference/src/decoding/speculative_decoding.py:428-598 (loop), :294-400 (sample),
// :847-869 (accepted lengths), :871-904 (nucleus truncation), :177-238 (top-k per query).
//
// Per iteration the engine runs
//   beam_prepare  -> per-candidate bookkeeping (first free slot, finished flag, live-row prefix)
//   beam_fill_rows-> (candidate, draft) decoder input rows, live candidates only
//   decoder stack on the live rows (full-prefix, causal), gather of the dl+1 scored positions,
//   classifier GEMM
//   beam_stats    -> per (row, position): softmax max/sum, the n_best largest logits (sorted), size of
//                    the nucleus-truncated support
//   beam_choose   -> accepted length of every draft, best draft per candidate (torch-CPU topk(1) order)
//   beam_expand   -> per query: scores of all leaves of the continuation trees of its candidates, the
//                    n_best best become the next candidates
//   (loop control: all-finished flag, trailing empty columns, acceptance statistics -> the last CTA of beam_expand)
#include "kernels.cuh"
#include "topk_emul.cuh"

namespace ttb {
// Accumulators of the loop control behind the BC_COUNT control words (reset by the CTA that finalises an iteration), and the
// loop state the device carries itself so that an iteration's kernel arguments do not change from one iteration to the next
// (the iteration is replayed as a CUDA graph): width of the token matrix and number of the CURRENT iteration, advanced by
// the CTA that closes the iteration with the same rule the host applies (engine.cu:beam_api, speculative_decoding.py:452-470)
constexpr int BCX_ALL_FIN = BC_COUNT, BCX_MIN_PAD = BC_COUNT + 1, BCX_ACC = BC_COUNT + 2, BCX_CNT = BC_COUNT + 3, BCX_TICKET = BC_COUNT + 4,
              BCX_W = BC_COUNT + 5, BCX_ITER = BC_COUNT + 6,
              BCX_NLIVE_NEXT = BC_COUNT + 8,   // accumulator: new hypotheses without EOS = live candidates of the next iteration
              BCX_DONE = BC_COUNT + 7;   // the loop has ended (every hypothesis finished, length budget used up, or an error): set by the CTA that
                                         // closes an iteration with the reference's own stop rule, so that an iteration the host has
                                         // enqueued ahead of reading this one's outcome is a no-op on the device


// drafts of candidate c (query q): all N source drafts, or in smart mode the library windows keyed by its last token
__device__ __forceinline__ int cand_n_drafts(const BeamState& st, int c) { return st.smart ? st.c_cnt[c] : st.N; }
__device__ __forceinline__ const int* cand_draft(const BeamState& st, int c, int q, int j) {
    if (!st.smart) return st.drafts + ((long long)q * st.N + j) * st.dl0;
    const int n = st.tok_list[((long long)q * st.V + st.c_last[c]) * st.N + j];
    return st.drafts + ((long long)q * st.n_lib + n) * st.dl0 + 1;      // skip the key token
}

// ---- smart drafts: windows of the library grouped by their first token (get_vocab_tokens_bool_lib, :402-420) ------
__global__ void beam_build_lib_kernel(BeamState st) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= st.B * st.V) return;
    const int b = w / st.V, v = w % st.V;
    int* list = st.tok_list + (long long)w * st.N;
    int cnt = 0;
    for (int n0 = 0; n0 < st.n_lib && cnt < st.N; n0 += 32) {
        const int n = n0 + lane;
        const bool hit = n < st.n_lib && st.drafts[((long long)b * st.n_lib + n) * st.dl0] == v;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < st.N) list[pos] = n;                 // the first N windows in library order (:419)
        cnt += __popc(m);
    }
    if (lane == 0) {
        if (cnt == 0) { list[0] = 0; cnt = 1; }              // "each line needs at least one draft" (:418)
        st.tok_cnt[w] = cnt < st.N ? cnt : st.N;
    }
}
void launch_beam_build_lib(const BeamState& st, cudaStream_t s) {
    const int warps = st.B * st.V;
    beam_build_lib_kernel<<<(warps + 7) / 8, 256, 0, s>>>(st);
}

// ---- prepare ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) beam_prepare_kernel(BeamState st, int C, int beam, int dl) {
    if (st.ctrl[BCX_DONE]) {   // enqueued ahead of the end of the loop: no live rows, every kernel behind this one exits at once
        if (threadIdx.x == 0) { st.ctrl[BC_NLIVE_ROWS] = 0; st.ctrl[BC_NLIVE_CANDS] = 0; }
        return;
    }
    const int W = st.ctrl[BCX_W];
    // one warp per candidate row, coalesced scan: first PAD column, EOS anywhere, a real token behind the first PAD
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int c = warp; c < C; c += n_warps) {
        const int* row = st.cand_cur + (long long)c * st.ldw;
        int first_pad = 0x7fffffff, last_real = -1, fin = 0;
        for (int j = lane; j < W; j += 32) {
            const int t = row[j];
            if (t == st.eos) fin = 1;
            if (t == st.pad) first_pad = min(first_pad, j); else last_real = max(last_real, j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            first_pad = min(first_pad, __shfl_xor_sync(0xffffffffu, first_pad, o));
            last_real = max(last_real, __shfl_xor_sync(0xffffffffu, last_real, o));
            fin |= __shfl_xor_sync(0xffffffffu, fin, o);
        }
        if (lane == 0) {
            const int slot0 = first_pad == 0x7fffffff ? -1 : first_pad;
            const int hole = slot0 >= 0 && last_real > slot0;   // a real token after a PAD: draft slots not contiguous
            st.c_slot0[c] = slot0;
            st.c_fin[c] = fin;
            if (hole || slot0 < 1 || slot0 + dl + 1 > W) atomicExch(&st.ctrl[BC_ERROR], 3);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, live = 0, lmax = 0;
        for (int c = 0; c < C; ++c) {
            st.c_rowbase[c] = run;
            const int slot0 = st.c_slot0[c];
            st.c_front[c] = max(0, slot0 - 1);   // clamped: an inconsistent hypothesis raises BC_ERROR, its rows are discarded
            int cnt = st.N;
            if (st.smart) {
                // last meaningful token = the one before the first PAD (hypotheses without interior PADs), or the last
                // column when the row has no PAD at all (:693-697)
                const int last = st.cand_cur[(long long)c * st.ldw + (slot0 >= 1 ? slot0 - 1 : (slot0 < 0 ? W - 1 : 0))];
                st.c_last[c] = last;
                cnt = st.tok_cnt[(c / beam) * st.V + last];
                st.c_cnt[c] = cnt;
            }
            lmax = max(lmax, cnt);
            if (!st.c_fin[c]) {
                run += cnt;
                if (!st.smart) {
                    st.live_cand[live] = c;
                    st.live_query[live] = c / beam;
                    if (st.desc_self) {
                        const int f = st.c_front[c];
                        st.desc_self[live] = make_int4(c, f, st.cand_cur[(long long)c * st.ldw + f], 0x7fffffff);
                        st.desc_cross[live] = make_int4(c / beam, 0, 0, st.src_len ? st.src_len[c / beam] : 0x7fffffff);
                    }
                }
                ++live;
            }
        }
        st.ctrl[BC_NLIVE_ROWS] = run;
        st.ctrl[BC_NLIVE_CANDS] = st.smart ? run : live;     // attention groups: candidates, or single rows in smart mode
        st.ctrl[BC_LMAX] = lmax;
    }
    __syncthreads();
    if (st.smart) {   // every live row is its own attention group: row -> (candidate, query, library window)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (st.c_fin[c]) continue;
            const int q = c / beam, base = st.c_rowbase[c], cnt = st.c_cnt[c];
            const int* list = st.tok_list + ((long long)q * st.V + st.c_last[c]) * st.N;
            const int f = st.c_front[c];
            const int4 ds = make_int4(c, f, st.cand_cur[(long long)c * st.ldw + f], 0x7fffffff);
            const int4 dc = make_int4(q, 0, 0, st.src_len ? st.src_len[q] : 0x7fffffff);
            for (int j = 0; j < cnt; ++j) {
                st.live_cand[base + j] = c;
                st.live_query[base + j] = q;
                st.row_draft[base + j] = list[j];
                if (st.desc_self) { st.desc_self[base + j] = ds; st.desc_cross[base + j] = dc; }
            }
        }
    }
}
void launch_beam_prepare(const BeamState& st, int C, int beam, int dl, cudaStream_t s) {
    beam_prepare_kernel<<<1, 1024, 0, s>>>(st, C, beam, dl);
}

// ---- KV-cached decoder pass ---------------------------------------------------------------------------------
// Row t = (live candidate g, draft n, position i): token = last token of the candidate (i = 0) or draft token i-1,
// sequence position front + i.  Row order matches c_rowbase (rows of live candidate g start at g * N).
template <typename ActT>
__global__ void beam_embed_cached_kernel(BeamState st, int beam, int dl, const float* __restrict__ table, const float* __restrict__ pe,
                                         int E, float* __restrict__ x, ActT* __restrict__ xh) {
    pdl_wait();   // launched with the programmatic attribute: scheduled while the predecessor drains, starts when it is complete
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= st.ctrl[BC_NLIVE_ROWS] * (dl + 1)) return;
    const int r = t / (dl + 1), i = t % (dl + 1);
    int c, f, tok;
    if (st.smart) {
        c = st.live_cand[r];
        f = st.c_front[c];
        tok = (i == 0) ? st.cand_cur[(long long)c * st.ldw + f]
                       : st.drafts[((long long)(c / beam) * st.n_lib + st.row_draft[r]) * st.dl0 + i];   // window token i (0 = key)
    } else {
        const int g = r / st.N, n = r % st.N;
        c = st.live_cand[g];
        f = st.c_front[c];
        tok = (i == 0) ? st.cand_cur[(long long)c * st.ldw + f] : st.drafts[((long long)(c / beam) * st.N + n) * st.dl0 + i - 1];
    }
    if (lane == 0 && st.row_tok) {   // the draft token at input position i is the "next draft token" of scored position i - 1
        if (i >= 1) st.row_tok[t - 1] = tok;
        if (i == dl) st.row_tok[t] = -1;
    }
    const float* e = table + (long long)tok * E;
    const float* p = pe + (long long)(f + i + 1) * E;
    for (int col = lane; col < E; col += 32) {
        const float v = e[col] + p[col];
        x[(long long)t * E + col] = v;
        if (xh) xh[(long long)t * E + col] = from_f32<ActT>(v);
    }
}
template <typename ActT>
void launch_beam_embed_cached(const BeamState& st, int beam, int max_rows, int dl, const float* table, const float* pe, int E,
                              float* x, ActT* xh, cudaStream_t s) {
    const int T = max_rows * (dl + 1);
    if (T <= 0) return;
    launch_pdl(beam_embed_cached_kernel<ActT>, dim3((T + 7) / 8), dim3(256), 0, s, st, beam, dl, table, pe, E, x, xh);
}
template void launch_beam_embed_cached<float>(const BeamState&, int, int, int, const float*, const float*, int, float*, float*, cudaStream_t);
template void launch_beam_embed_cached<__nv_bfloat16>(const BeamState&, int, int, int, const float*, const float*, int, float*, __nv_bfloat16*,
                                                      cudaStream_t);

template <typename ActT>
__global__ void beam_cache_update_kernel(BeamState st, int dl, const ActT* __restrict__ qkv_all, long long qkv_layer_stride, int qkv_ld,
                                         int E, const ActT* __restrict__ kc_cur, const ActT* __restrict__ vc_cur, ActT* __restrict__ kc_next,
                                         ActT* __restrict__ vc_next, long long cache_layer_stride, long long cache_cand_stride) {
    pdl_wait();   // launched with the programmatic attribute: scheduled while the predecessor drains, starts when it is complete
    if (st.ctrl[BCX_DONE]) return;   // the loop has ended (this iteration closed it, or it was enqueued behind the end): the caches are not read again
    const int cn = blockIdx.x, l = blockIdx.y;
    const int parent = st.n_parent[cn];
    if (parent < 0) return;                         // child of a finished hypothesis: never decoded again
    const int f = st.c_front[parent], keep = st.n_keep[cn], r = st.n_row[cn];
    const long long lo = (long long)l * cache_layer_stride;
    const ActT* ks = kc_cur + lo + (long long)parent * cache_cand_stride;
    const ActT* vs = vc_cur + lo + (long long)parent * cache_cand_stride;
    ActT* kd = kc_next + lo + (long long)cn * cache_cand_stride;
    ActT* vd = vc_next + lo + (long long)cn * cache_cand_stride;
    // parent prefix: f rows of E elements, 16-byte vectors (E * sizeof(ActT) is a multiple of 16)
    constexpr int VEC = 16 / (int)sizeof(ActT);
    const long long nvec = (long long)f * E / VEC;
    const uint4* ks4 = reinterpret_cast<const uint4*>(ks);
    const uint4* vs4 = reinterpret_cast<const uint4*>(vs);
    uint4* kd4 = reinterpret_cast<uint4*>(kd);
    uint4* vd4 = reinterpret_cast<uint4*>(vd);
    for (long long i = (long long)blockIdx.z * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.z * blockDim.x) { kd4[i] = ks4[i]; vd4[i] = vs4[i]; }
    if (blockIdx.z != 0) return;
    // positions f .. f + keep from the chosen draft row of this iteration
    const ActT* src = qkv_all + (long long)l * qkv_layer_stride + (long long)r * (dl + 1) * qkv_ld;
    for (int idx = threadIdx.x; idx < (keep + 1) * E; idx += blockDim.x) {
        const int i = idx / E, col = idx % E;
        kd[(long long)(f + i) * E + col] = src[(long long)i * qkv_ld + E + col];
        vd[(long long)(f + i) * E + col] = src[(long long)i * qkv_ld + 2 * E + col];
    }
}
template <typename ActT>
void launch_beam_cache_update(const BeamState& st, int dl, const ActT* qkv_all, long long qkv_layer_stride, int n_layers, int qkv_ld,
                              int E, const ActT* kc_cur, const ActT* vc_cur, ActT* kc_next, ActT* vc_next, long long cache_layer_stride,
                              long long cache_cand_stride, cudaStream_t s) {
    dim3 grid(st.B * st.K, n_layers, 4);   // z: four slices of the parent's prefix
    launch_pdl(beam_cache_update_kernel<ActT>, dim3(grid), dim3(256), 0, s, st, dl, qkv_all, qkv_layer_stride, qkv_ld, E, kc_cur, vc_cur, kc_next, vc_next,
                                                        cache_layer_stride, cache_cand_stride);
}
template void launch_beam_cache_update<float>(const BeamState&, int, const float*, long long, int, int, int, const float*, const float*, float*,
                                              float*, long long, long long, cudaStream_t);
template void launch_beam_cache_update<__nv_bfloat16>(const BeamState&, int, const __nv_bfloat16*, long long, int, int, int,
                                                      const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, long long,
                                                      long long, cudaStream_t);

__global__ void beam_fill_rows_kernel(BeamState st, int C, int beam, int W, int dl) {
    const int c = blockIdx.x / st.N, n = blockIdx.x % st.N;
    if (c >= C || st.c_fin[c]) return;
    const int q = c / beam;
    const int r = st.c_rowbase[c] + n;
    const int slot0 = st.c_slot0[c];
    const int* src = st.cand_cur + (long long)c * st.ldw;
    const int* dr = st.drafts + ((long long)q * st.N + n) * st.dl0;
    int* dst = st.rows_tok + (long long)r * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        int t = src[j];
        if (j >= slot0 && j < slot0 + dl) t = dr[j - slot0];
        dst[j] = t;
    }
    if (threadIdx.x == 0) { st.row_cand[r] = c; st.row_query[r] = q; st.row_slot0[r] = slot0; }
}
void launch_beam_fill_rows(const BeamState& st, int C, int beam, int W, int dl, cudaStream_t s) {
    beam_fill_rows_kernel<<<C * st.N, 128, 0, s>>>(st, C, beam, W, dl);
}

// rows of the residual stream that are scored: positions slot0-1 .. slot0+dl-1 of every live row
template <typename ActT>
__global__ void beam_gather_kernel(BeamState st, const float* __restrict__ x, const ActT* __restrict__ xh, int W, int dl, int E,
                                   float* __restrict__ xg, ActT* __restrict__ xgh) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= st.ctrl[BC_NLIVE_ROWS] * (dl + 1)) return;
    const int r = t / (dl + 1), i = t % (dl + 1);
    const long long src = ((long long)r * W + st.row_slot0[r] - 1 + i) * E, dst = (long long)t * E;
    for (int c = lane; c < E; c += 32) {
        if (xg) xg[dst + c] = x[src + c];
        if (xgh) xgh[dst + c] = xh[src + c];
    }
}
template <typename ActT>
void launch_beam_gather(const BeamState& st, const float* x, const ActT* xh, int max_rows, int W, int dl, int E,
                        float* xg, ActT* xgh, cudaStream_t s) {
    const int T = max_rows * (dl + 1);
    if (T <= 0) return;
    beam_gather_kernel<ActT><<<(T + 7) / 8, 256, 0, s>>>(st, x, xh, W, dl, E, xg, xgh);
}
template void launch_beam_gather<float>(const BeamState&, const float*, const float*, int, int, int, int, float*, float*, cudaStream_t);
template void launch_beam_gather<__nv_bfloat16>(const BeamState&, const float*, const __nv_bfloat16*, int, int, int, int, float*, __nv_bfloat16*, cudaStream_t);

// ---- per (row, position) statistics -------------------------------------------------------------------
// One warp per distribution: softmax max / sum, the K largest logits in descending order (ties: lower
// token id first) and the size of the truncated support (exclusive cumulative probability < 0.9975,
// the best token always kept; speculative_decoding.py:886-899).
template <int VPL>
__global__ void __launch_bounds__(256) beam_stats_kernel(BeamState st, const float* __restrict__ logits, int dl) {
    pdl_wait();   // launched with the programmatic attribute: scheduled while the predecessor drains, starts when it is complete
    const int rp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (rp >= st.ctrl[BC_NLIVE_ROWS] * (dl + 1)) return;
    const int V = st.V, K = st.K;
    const float* p = logits + (long long)rp * V;
    float v[VPL];
    float mx = -INFINITY;
    const int want = st.row_tok ? st.row_tok[rp] : -1;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int c = lane + 32 * k;
        v[k] = c < V ? p[c] : -INFINITY;
        mx = fmaxf(mx, v[k]);
        if (c == want) st.tokv[rp] = v[k];
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) sum += (lane + 32 * k < V) ? expf(v[k] - mx) : 0.f;
    sum = warp_sum(sum);
    float* topv = st.topv + (long long)rp * K;
    int* topi = st.topi + (long long)rp * K;
    float my_v = -INFINITY;
    int my_i = -1;
    for (int j = 0; j < K; ++j) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int c = lane + 32 * k;
            if (c < V && (v[k] > bv || (v[k] == bv && c < bi))) { bv = v[k]; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k)
            if (lane + 32 * k == bi) v[k] = -INFINITY;                          // taken
        // every lane holds (bv, bi) after the butterfly: lane j keeps entry j for the support scan below
        if (lane == (j & 31)) { my_v = bv; my_i = bi == 0x7fffffff ? -1 : bi; }
        if (lane == 0) { topv[j] = bv; topi[j] = bi == 0x7fffffff ? -1 : bi; }
    }
    {
        // truncated support: entry j is kept while the exclusive cumulative probability of entries 0..j-1 is < 0.9975
        // (sequential sum in entry order, as the reference's cumsum; K <= 32 entries live one per lane)
        int keep = 1;
        float cum = 0.f;
        for (int j = 1; j < K; ++j) {
            const float pv = __shfl_sync(0xffffffffu, my_v, j - 1);
            const int ij = __shfl_sync(0xffffffffu, my_i, j);
            if (ij < 0) break;
            cum += expf(pv - mx) / sum;
            if (cum < 0.9975f) keep = j + 1; else break;
        }
        if (lane == 0) {
            st.lmax[rp] = mx;
            st.lsum[rp] = sum;
            st.nkeep[rp] = keep;
        }
    }
}
void launch_beam_stats(const BeamState& st, const float* logits, int max_rows, int dl, cudaStream_t s) {
    const int T = max_rows * (dl + 1);
    if (T <= 0) return;
    if (st.V <= 320) launch_pdl(beam_stats_kernel<10>, dim3((T + 7) / 8), dim3(256), 0, s, st, logits, dl);
    else if (st.V <= 512) launch_pdl(beam_stats_kernel<16>, dim3((T + 7) / 8), dim3(256), 0, s, st, logits, dl);
    else launch_pdl(beam_stats_kernel<32>, dim3((T + 7) / 8), dim3(256), 0, s, st, logits, dl);
}

// ---- accepted lengths + best draft per candidate ------------------------------------------------------------
// One CTA per candidate.  "Is draft token a of draft n inside the truncated support of its position?" is evaluated for all
// (draft, position) pairs in parallel (two dependent global loads each instead of a chain of up to 2 dl per draft); the
// accepted length of a draft is the length of its leading run of hits.
__global__ void __launch_bounds__(256) beam_choose_kernel(BeamState st, int C, int beam, int dl, int par) {
    pdl_wait();   // launched with the programmatic attribute: scheduled while the predecessor drains, starts when it is complete
    if (st.ctrl[BCX_DONE]) return;   // no-op iteration behind the end of the loop
    const int iter = st.ctrl[BCX_ITER];
    const int c = blockIdx.x;
    __shared__ int s_nacc[64];
    extern __shared__ unsigned char s_hit[];                 // [m][dl] when par
    if (c >= C) return;
    const int q = c / beam, N = st.N, K = st.K;
    const bool fin = st.c_fin[c] != 0;
    const int m = cand_n_drafts(st, c);                      // drafts of this candidate
    const int lmax = st.smart ? st.ctrl[BC_LMAX] : N;        // ragged groups are padded with -1 up to the longest (:206-223)
    const int rb = st.c_rowbase[c];
    auto hit = [&](int n, int a) -> bool {
        const long long rp = (long long)(rb + n) * (dl + 1) + a;
        const int keep = st.nkeep[rp];
        const int* ti = st.topi + rp * K;
        const int tok = cand_draft(st, c, q, n)[a];
        bool in = false;
        for (int j = 0; j < keep; ++j) in |= (ti[j] == tok);
        return in;
    };
    if (par && !fin) {
        for (int idx = threadIdx.x; idx < m * dl; idx += blockDim.x) s_hit[idx] = hit(idx / dl, idx % dl) ? 1 : 0;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        int a = -1;
        if (n < m) {
            a = 0;
            if (!fin) {
                if (par) { while (a < dl && s_hit[n * dl + a]) ++a; }
                else { while (a < dl && hit(n, a)) ++a; }
            }
        }
        st.c_nacc[(long long)c * N + n] = a;
        if (n < 64) s_nacc[n] = a;
        if (st.trace_nacc) st.trace_nacc[((long long)iter * st.B * K + c) * N + n] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int pick = 0;
        if (st.tie_break == 0 && lmax < 64) {
            pick = topk1_torch_cpu(s_nacc, lmax);
        } else {
            for (int n = 1; n < m; ++n) if (st.c_nacc[(long long)c * N + n] > st.c_nacc[(long long)c * N + pick]) pick = n;
        }
        st.c_pick[c] = pick;
        if (st.trace_pick) st.trace_pick[(long long)iter * st.B * K + c] = pick;
    }
}
void launch_beam_choose(const BeamState& st, int C, int beam, int dl, cudaStream_t s) {
    const size_t smem = (size_t)st.N * (dl > 0 ? dl : 1);
    const int par = smem <= 40 * 1024 ? 1 : 0;
    launch_pdl(beam_choose_kernel, dim3(C), dim3(256), par ? smem : 0, s, st, C, beam, dl, par);
}

// ---- leaves of the continuation trees, n_best best per query, loop control ---------------------------------------
// Accumulators of the loop control behind the BC_COUNT control words (reset by the CTA that finalises an iteration)
// log softmax exactly as the reference evaluates it: log(exp(x - max) / sum)   (:378)
__device__ __forceinline__ float ref_logprob(float logit, float mx, float sum) { return logf(expf(logit - mx) / sum); }

__global__ void __launch_bounds__(256) beam_expand_kernel(BeamState st, int beam, int dl, const float* __restrict__ logits) {
    pdl_wait();   // launched with the programmatic attribute: scheduled while the predecessor drains, starts when it is complete
    if (st.ctrl[BCX_DONE]) return;   // no-op iteration behind the end of the loop (the word of the last real iteration stays posted)
    extern __shared__ float s_score[];                 // [beam][(dl+1)][K] leaf scores, -inf when absent
    __shared__ float s_red_v[256];
    __shared__ int s_red_i[256];
    __shared__ int s_sel[64];
    __shared__ int s_valid;
    const int q = blockIdx.x, K = st.K, V = st.V, N = st.N;
    const int W = st.ctrl[BCX_W];                      // rewritten only by the CTA that closes the iteration, after every CTA has arrived
    const int per_c = (dl + 1) * K, total = beam * per_c;
    float* s_pre = s_score + total;                    // [beam][dl+2] running log-prob of the accepted path
    if (threadIdx.x == 0) s_valid = 0;
    // accepted-path prefix sums (sequential fp32 adds, same association as the reference's cumsum): one warp per
    // candidate, the terms are fetched by the lanes in parallel and added in order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int cb = warp; cb < beam; cb += n_warps) {
        const int c = q * beam + cb;
        float run = 0.f;
        if (lane == 0) s_pre[cb * (dl + 2)] = 0.f;
        if (!st.c_fin[c]) {
            const int n = st.c_pick[c], a = st.c_nacc[(long long)c * N + n], r = st.c_rowbase[c] + n;
            const int* dr = cand_draft(st, c, q, n);
            for (int i0 = 0; i0 < a; i0 += 32) {
                float term = 0.f;
                if (i0 + lane < a) {
                    const long long rp = (long long)r * (dl + 1) + i0 + lane;
                    term = ref_logprob(st.tokv ? st.tokv[rp] : logits[rp * V + dr[i0 + lane]], st.lmax[rp], st.lsum[rp]);
                }
                const int cnt = min(32, a - i0);
                for (int i = 0; i < cnt; ++i) {
                    run += __shfl_sync(0xffffffffu, term, i);
                    if (lane == 0) s_pre[cb * (dl + 2) + i0 + i + 1] = run;
                }
            }
        }
    }
    __syncthreads();
    int my_valid = 0;
    for (int L = threadIdx.x; L < total; L += blockDim.x) {
        const int cb = L / per_c, p = (L % per_c) / K, j = L % K;
        const int c = q * beam + cb;
        float score = -INFINITY;
        if (st.c_fin[c]) {
            // finished candidate: the artificial distribution leaves exactly one leaf, PAD at position 0,
            // with log-probability log(1) = 0
            if (p == 0 && j == 0) score = st.logp_cur[c] + 0.f;
        } else {
            const int n = st.c_pick[c], a = st.c_nacc[(long long)c * N + n], r = st.c_rowbase[c] + n;
            if (p <= a) {
                const long long rp = (long long)r * (dl + 1) + p;
                const int tok = st.topi[rp * K + j];
                const float lg = st.topv[rp * K + j];
                const int* dr = cand_draft(st, c, q, n);
                // excluded: the draft token that continues the accepted path (p < a), BOS in the slot of the first
                // rejected draft token (p == a < dl); entries whose logit is exactly 0.0 vanish in the reference
                const int excl = (p < dl) ? (p < a ? dr[p] : st.bos) : -1;
                if (tok >= 0 && tok != excl && lg != 0.0f)
                    score = st.logp_cur[c] + (s_pre[cb * (dl + 2) + p] + ref_logprob(lg, st.lmax[rp], st.lsum[rp]));
            }
        }
        s_score[L] = score;
        my_valid += score != -INFINITY ? 1 : 0;
    }
    if (my_valid) atomicAdd(&s_valid, my_valid);
    __syncthreads();
    const bool ok = s_valid >= K;
    if (!ok && threadIdx.x == 0) atomicExch(&st.ctrl[BC_ERROR], 4);   // reference: assert min(group length) >= k (:195)
    for (int k = 0; ok && k < K; ++k) {                // K rounds of block-wide argmax (ties: lower leaf index)
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int L = threadIdx.x; L < total; L += blockDim.x) {
            const float sc = s_score[L];
            if (sc > bv || (sc == bv && sc != -INFINITY && L < bi)) { bv = sc; bi = L; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_red_v[warp] = bv; s_red_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = lane < n_warps ? s_red_v[lane] : -INFINITY;
            bi = lane < n_warps ? s_red_i[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                s_sel[k] = bi;
                st.logp_next[q * K + k] = bv;
                s_score[bi] = -INFINITY;
            }
        }
        __syncthreads();
    }
    // materialise the K new candidates: one warp per candidate; the loop control's view of the new rows (PAD columns,
    // EOS anywhere, accepted draft tokens) is accumulated on the way
    for (int k = warp; ok && k < K; k += n_warps) {
        const int L = s_sel[k];
        const int cb = L / per_c, p = (L % per_c) / K, j = L % K;
        const int c = q * beam + cb;
        const int* src = st.cand_cur + (long long)c * st.ldw;
        int* dst = st.cand_next + (long long)(q * K + k) * st.ldw;
        const int slot0 = st.c_slot0[c];
        const bool fin = st.c_fin[c] != 0;
        int n = 0, r = 0;
        if (!fin) { n = st.c_pick[c]; r = st.c_rowbase[c] + n; }
        const int* dr = fin ? st.drafts : cand_draft(st, c, q, n);
        const int tok = fin ? st.pad : st.topi[((long long)r * (dl + 1) + p) * K + j];
        int pads = 0, has_eos = 0;
        for (int col = lane; col < W; col += 32) {
            int t = src[col];
            if (col >= slot0 && col <= slot0 + dl) {
                const int o = col - slot0;
                t = o < p ? dr[o] : (o == p ? tok : st.pad);
            }
            dst[col] = t;
            pads += t == st.pad;
            has_eos |= t == st.eos;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pads += __shfl_xor_sync(0xffffffffu, pads, o);
            has_eos |= __shfl_xor_sync(0xffffffffu, has_eos, o);
        }
        if (lane == 0) {
            st.acc_stat[q * K + k] = fin ? -1 : p;
            st.n_parent[q * K + k] = fin ? -1 : c;
            st.n_keep[q * K + k] = p;
            st.n_row[q * K + k] = r;
            if (!has_eos) { atomicAnd(&st.ctrl[BCX_ALL_FIN], 0); atomicAdd(&st.ctrl[BCX_NLIVE_NEXT], 1); }
            atomicMin(&st.ctrl[BCX_MIN_PAD], pads);
            if (!fin) { atomicAdd(&st.ctrl[BCX_ACC], p); atomicAdd(&st.ctrl[BCX_CNT], 1); }
        }
    }
    // ---- loop control: the CTA that arrives last publishes the control words of the iteration (also into the pinned
    // host mirror, sequence number last) and resets the accumulators
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&st.ctrl[BCX_TICKET], 1) == (int)gridDim.x - 1) {
            __threadfence();
            const int acc = atomicAdd(&st.ctrl[BCX_ACC], 0), cnt = atomicAdd(&st.ctrl[BCX_CNT], 0);
            st.ctrl[BC_ALL_FINISHED] = atomicAdd(&st.ctrl[BCX_ALL_FIN], 0);
            st.ctrl[BC_EMPTY_COLS] = atomicAdd(&st.ctrl[BCX_MIN_PAD], 0);
            st.ctrl[BC_ACCEPTED] += acc;
            st.ctrl[BC_PRODUCED] += acc + cnt;
            const int nlive_next = atomicAdd(&st.ctrl[BCX_NLIVE_NEXT], 0);
            st.ctrl[BCX_ALL_FIN] = 1; st.ctrl[BCX_MIN_PAD] = 0x7fffffff; st.ctrl[BCX_ACC] = 0; st.ctrl[BCX_CNT] = 0; st.ctrl[BCX_TICKET] = 0;
            st.ctrl[BCX_NLIVE_NEXT] = 0;
            // width of the next iteration's token matrix (speculative_decoding.py:452-470, mirrored by the host loop)
            const int seq = st.ctrl[BCX_ITER] + 1;
            {
                const int empty = st.ctrl[BC_EMPTY_COLS], filled = W - empty, budget = st.max_len - filled - 1;
                const int grow = min(budget, dl) + 1 - empty;
                st.ctrl[BCX_W] = W + max(grow, 0);
                st.ctrl[BCX_ITER] = seq;
                // the reference's stop rule (:570-598 break on all finished; :452 while budget >= 1 and filled <= max_len)
                if (atomicAdd(&st.ctrl[BC_ERROR], 0) != 0 || st.ctrl[BC_ALL_FINISHED] != 0 || budget < 1 || filled > st.max_len) st.ctrl[BCX_DONE] = 1;
            }
            if (st.host_ctrl) {
                // what the host needs per iteration, packed into ONE 64-bit word of its own (pinned, mapped) memory:
                // sequence number | error | all finished | empty columns.  One posted store: no copy engine, no stream
                // synchronisation, no system-scope fence between this kernel and the cache update that follows it.
                // sequence (24 bits) | error (4) | all finished (1) | empty columns (12) | live candidates of the next iteration (16)
                const unsigned long long word = ((unsigned long long)(seq & 0xffffff) << 40) |
                                                ((unsigned long long)(atomicAdd(&st.ctrl[BC_ERROR], 0) & 0xf) << 36) |
                                                ((unsigned long long)(st.ctrl[BC_ALL_FINISHED] & 1) << 35) |
                                                ((unsigned long long)(st.ctrl[BC_EMPTY_COLS] & 0xfff) << 23) |
                                                ((unsigned long long)(nlive_next & 0xffff) << 7);
                // ring of four words: the host may read iteration i's word after iteration i + 1 (enqueued ahead) has posted its own
                reinterpret_cast<volatile unsigned long long*>(st.host_ctrl)[seq & 3] = word;
            }
        }
    }
}
void launch_beam_expand(const BeamState& st, int beam, int dl, const float* logits, cudaStream_t s) {
    const size_t smem = ((size_t)beam * (dl + 1) * st.K + (size_t)beam * (dl + 2)) * sizeof(float);
    if (ensure_dyn_smem(beam_expand_kernel, (int)smem)) return;   // error recorded; the launch below fails and cudaGetLastError reports it
    launch_pdl(beam_expand_kernel, dim3(st.B), dim3(256), smem, s, st, beam, dl, logits);
}

__global__ void beam_init_kernel(BeamState st) {
    const long long n = (long long)st.B * st.K * st.ldw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        st.cand_cur[i] = (i % st.ldw == 0 && i / st.ldw < st.B) ? st.bos : st.pad;
        st.cand_next[i] = st.pad;
    }
    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b < st.B * st.K; b += blockDim.x) { st.logp_cur[b] = 0.f; st.logp_next[b] = 0.f; }
        if (threadIdx.x < BC_COUNT) st.ctrl[threadIdx.x] = 0;
        if (threadIdx.x == 0) {
            st.ctrl[BCX_ALL_FIN] = 1; st.ctrl[BCX_MIN_PAD] = 0x7fffffff; st.ctrl[BCX_ACC] = 0; st.ctrl[BCX_CNT] = 0; st.ctrl[BCX_TICKET] = 0;
            st.ctrl[BCX_W] = st.w0; st.ctrl[BCX_ITER] = 0; st.ctrl[BCX_DONE] = 0; st.ctrl[BCX_NLIVE_NEXT] = 0;
        }
    }
}
void launch_beam_init(const BeamState& st, cudaStream_t s) { beam_init_kernel<<<64, 256, 0, s>>>(st); }

// (rows, W) int32 with row stride ldw -> dense int64
__global__ void beam_export_kernel(const int* __restrict__ cand, int ldw, int R, int W, long long* __restrict__ out) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx < (long long)R * W) out[idx] = cand[(idx / W) * ldw + idx % W];
}
void launch_beam_export(const int* cand, int ldw, int R, int W, long long* out, cudaStream_t s) {
    const long long n = (long long)R * W;
    if (n > 0) beam_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cand, ldw, R, W, out);
}

}  // namespace ttb
