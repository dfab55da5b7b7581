// Device-resident state machine of speculative greedy decoding
// (reference: /root/reference/src/decoding/speculative_decoding.py:39-174).
//
// The host never looks at tokens while decoding.  Per iteration the engine launches
//   greedy_embed -> decoder layers -> classifier + argmax -> greedy_accept -> cache_append
// and the accept kernel (a single CTA: the bookkeeping of a whole batch is a few KB) does what the
// reference does with ~40 tensor ops and two host syncs per iteration:
//   * accepted length of every draft  (verification / cumsum / sum, :129-131)
//   * best draft per query            (topk(1), :133; tie-break = torch CPU topk emulation)
//   * token append + front index      (:136-146)
//   * retirement of finished queries into the output buffer and order-preserving compaction of
//     the live list (:149-168)
//   * the width bookkeeping of the reference's shared token matrix (:93-102), which decides when
//     the loop stops and when the reference itself would fail (see oracle/greedy_speculative.py).
#include "kernels.cuh"
#include "topk_emul.cuh"

namespace ttb {

// ---- width plan of the coming iteration (speculative_decoding.py:93-102) -----------------------
// Executed by one CTA after the live list is final.  `W` is the current width of the reference's
// token matrix.  `active` / `front` / `gen` may point to shared-memory copies of the state (accept
// kernel) or to the global arrays (init kernel); `col_live` is [W] ints of shared scratch.
__device__ void plan_next_iteration(const GreedyState& st, const int* active, const int* front, const int* gen,
                                    int n_active, int W, int* /*s_tmp*/, int* /*col_live*/) {
    __syncthreads();
    if (n_active == 0 || W >= st.max_len) {
        if (threadIdx.x == 0) { st.ctrl[CTRL_DONE] = 1; st.ctrl[CTRL_N_LEFT] = n_active; st.ctrl[CTRL_N_ACTIVE] = 0; }
        return;
    }
    // a column is dead when every live row holds PAD there: one column per thread (and pass), rows in the inner loop, the
    // dead columns are counted by the barrier itself
    int dead_total = 0;
    for (int c0 = 0; c0 < W; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        int live = 0;
        if (c < W)
            for (int g = 0; g < n_active; ++g) {
                const int b = active[g];
                live |= (c <= front[b] && gen[(long long)b * st.gen_ld + c] != st.pad) ? 1 : 0;
            }
        dead_total += __syncthreads_count(c < W && !live);
    }
    const int Wn = W + st.D + 1 - dead_total;   // the barrier counts are the same in every thread
    int oob = 0;
    for (int g = threadIdx.x; g < n_active; g += blockDim.x)
        if (front[active[g]] + 1 + st.D > Wn - 1) oob = 1;
    const int any_oob = __syncthreads_or(oob);
    if (threadIdx.x == 0) {
        st.ctrl[CTRL_PREV_WIDTH] = W;
        st.ctrl[CTRL_WIDTH] = Wn;
        if (any_oob) { st.ctrl[CTRL_ERROR] = 1; st.ctrl[CTRL_DONE] = 1; st.ctrl[CTRL_N_LEFT] = n_active; st.ctrl[CTRL_N_ACTIVE] = 0; }
    }
}

__global__ void __launch_bounds__(256) greedy_init_kernel(GreedyState st) {
    __shared__ int s_tmp[2];
    extern __shared__ int s_init_dyn[];   // [gen_ld] column flags
    for (long long idx = threadIdx.x; idx < (long long)st.B * st.gen_ld; idx += blockDim.x)
        st.gen[idx] = (idx % st.gen_ld == 0) ? st.bos : st.pad;
    for (long long idx = threadIdx.x; idx < (long long)st.B * st.max_len; idx += blockDim.x) st.out[idx] = st.pad;
    for (int b = threadIdx.x; b < st.B; b += blockDim.x) {
        st.front[b] = 0;
        st.active[b] = b;
        st.desc[b] = make_int4(b, 0, st.bos, st.src_len ? st.src_len[b] : 0x7fffffff);
    }
    if (threadIdx.x < CTRL_COUNT) st.ctrl[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        st.ctrl[CTRL_N_ACTIVE] = st.B;
        st.ctrl[CTRL_LS] = st.Ls;
        // the most frequent tie (no draft accepted at all: every length is 0) has a fixed answer per n_drafts
        int pick = 0;
        if (st.tie_break == 0 && st.N > 1 && st.N < 64) {
            for (int n = 0; n < st.N; ++n) s_init_dyn[n] = n;   // value 0, index n
            pick = topk1_torch_cpu_packed(s_init_dyn, st.N);
        }
        st.ctrl[CTRL_ALLEQ_PICK] = pick;
    }
    plan_next_iteration(st, st.active, st.front, st.gen, st.B, 1, s_tmp, s_init_dyn);
}
void launch_greedy_init(const GreedyState& st, cudaStream_t s) {
    greedy_init_kernel<<<1, 256, (size_t)(st.gen_ld > 64 ? st.gen_ld : 64) * sizeof(int), s>>>(st);
}

// ---- first kernel of an iteration: KV-cache append of the previous iteration + step-token embedding ----------
template <typename ActT>
__global__ void greedy_advance_kernel(GreedyState st, const float* __restrict__ table, const float* __restrict__ pe,
                                      int E, float* __restrict__ x, ActT* __restrict__ xh, int n_embed_blocks,
                                      const ActT* __restrict__ qkv_all, long long qkv_layer_stride, int n_layers, int qkv_ld,
                                      ActT* __restrict__ kcache, ActT* __restrict__ vcache, long long cache_layer_stride,
                                      long long cache_query_stride, int cache_ld, long long vt_layer_stride, long long vt_query_stride,
                                      int vt_pitch) {
    // Launched with the programmatic attribute behind the accept kernel (scheduled while it runs, starts when it is
    // complete).  The kernels of the decoding iteration read the control words and descriptor table of the accept kernel
    // ahead of their dependency waits, so none of them may be scheduled before this kernel has SEEN the accept kernel
    // complete: its dependents are released right behind its own wait, not before (their prologues -- barrier set-up,
    // tensor-memory allocation, weight requests -- then overlap this kernel instead of following it).  Nothing this kernel
    // writes (embeddings, the appended K/V rows) is read by a later kernel ahead of that kernel's wait.
    pdl_wait();
#ifndef TTB_ADVANCE_NO_TRIGGER
    pdl_launch_dependents();
#endif
    // the control words and the per-slot records are fetched together, ahead of the tests that use them (as dependent
    // loads they would be three to five global round trips in a row); the slot indices are in range for every CTA
    const int done = st.ctrl[CTRL_DONE], n_sel = st.ctrl[CTRL_N_SEL], n_active = st.ctrl[CTRL_N_ACTIVE];
    if ((int)blockIdx.x >= n_embed_blocks) {
        // K/V of the accepted positions of the chosen draft (recorded in st.sel by the accept kernel) -> cache
        const int idx = blockIdx.x - n_embed_blocks;
        const int g = idx / n_layers, l = idx % n_layers;
        const int4 sel = *reinterpret_cast<const int4*>(st.sel + g * 4);
        if (done || g >= n_sel) return;   // after DONE neither the cache nor the embeddings are read again
        const int b = sel.x, f = sel.y, pick = sel.z, a = sel.w;
        const ActT* src = qkv_all + (long long)l * qkv_layer_stride + ((long long)g * st.N + pick) * (st.D + 1) * qkv_ld;
        ActT* kd = kcache + (long long)l * cache_layer_stride + (long long)b * cache_query_stride + (long long)f * cache_ld;
        if (vt_pitch > 0) {   // value cache kept transposed ([dim][position]) for the tcgen05 attention kernel
            ActT* vt = vcache + (long long)l * vt_layer_stride + (long long)b * vt_query_stride + f;
            for (int i2 = threadIdx.x; i2 < (a + 1) * E; i2 += blockDim.x) {
                const int i = i2 / E, c = i2 % E;
                kd[(long long)i * cache_ld + c] = src[(long long)i * qkv_ld + E + c];
                vt[(long long)c * vt_pitch + i] = src[(long long)i * qkv_ld + 2 * E + c];
            }
            return;
        }
        ActT* vd = vcache + (long long)l * cache_layer_stride + (long long)b * cache_query_stride + (long long)f * cache_ld;
        for (int i2 = threadIdx.x; i2 < (a + 1) * E; i2 += blockDim.x) {
            const int i = i2 / E, c = i2 % E;
            kd[(long long)i * cache_ld + c] = src[(long long)i * qkv_ld + E + c];
            vd[(long long)i * cache_ld + c] = src[(long long)i * qkv_ld + 2 * E + c];
        }
        return;
    }
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int per_q = st.N * (st.D + 1);
    const int g = min(t / per_q, st.B - 1), r = t % per_q, n = r / (st.D + 1), i = r % (st.D + 1);
    const int4 d = st.desc[g];
    if (done || t >= n_active * per_q) return;
    const int b = d.x, f = d.y;
    const int tok = (i == 0) ? d.z : st.drafts[((long long)b * st.N + n) * st.D + i - 1];
    const float* e = table + (long long)tok * E;
    const float* p = pe + (long long)(f + i + 1) * E;
    for (int c = lane; c < E; c += 32) {
        float v = e[c] + p[c];
        x[(long long)t * E + c] = v;
        if (xh) xh[(long long)t * E + c] = from_f32<ActT>(v);
    }
}
template <typename ActT>
void launch_greedy_advance(const GreedyState& st, const float* table, const float* pe, int E, float* x, ActT* xh,
                           const ActT* qkv_all, long long qkv_layer_stride, int n_layers, int qkv_ld, ActT* kcache, ActT* vcache,
                           long long cache_layer_stride, long long cache_query_stride, int cache_ld, cudaStream_t s,
                           long long vt_layer_stride, long long vt_query_stride, int vt_pitch) {
    const int T = st.B * st.N * (st.D + 1);
    const int n_embed_blocks = (T + 7) / 8;
    // first kernel of the captured iteration: plain launch (its predecessor is the previous graph launch)
    launch_pdl(greedy_advance_kernel<ActT>, dim3(n_embed_blocks + st.B * n_layers), dim3(256), 0, s, st, table, pe, E, x, xh, n_embed_blocks, qkv_all,
                                                                              qkv_layer_stride, n_layers, qkv_ld, kcache, vcache,
                                                                              cache_layer_stride, cache_query_stride, cache_ld, vt_layer_stride,
                                                                              vt_query_stride, vt_pitch);
}
template void launch_greedy_advance<float>(const GreedyState&, const float*, const float*, int, float*, float*, const float*, long long, int, int,
                                           float*, float*, long long, long long, int, cudaStream_t, long long, long long, int);
template void launch_greedy_advance<__nv_bfloat16>(const GreedyState&, const float*, const float*, int, float*, __nv_bfloat16*,
                                                   const __nv_bfloat16*, long long, int, int, __nv_bfloat16*, __nv_bfloat16*, long long,
                                                   long long, int, cudaStream_t, long long, long long, int);

// ---- accept / retire / plan ---------------------------------------------------------------------
// One CTA, one warp per live query: lanes score the drafts in parallel (accepted length = leading
// matches between draft tokens and the predictions one position earlier), lane 0 applies the
// tie-break rule, the warp appends the tokens and retires the query if it produced EOS.
// The whole bookkeeping state (live list, fronts, token matrix) is copied to shared memory with one
// round of independent loads at kernel entry, so the dependent chains (live list -> front -> tokens,
// compaction, dead-column scan) never wait on global memory; updates are written through.
constexpr int ACCEPT_MAX_SMEM_INTS = 50 * 1024;   // token matrices larger than 200 KB stay in global memory

#ifdef TTB_ACC_TIMELINE
__device__ long long g_acc_ts[16];
#define ACC_TS(i) do { if (threadIdx.x == 0) g_acc_ts[i] = clock64(); } while (0)
#else
#define ACC_TS(i) do { } while (0)
#endif
__global__ void __launch_bounds__(1024) greedy_accept_kernel(GreedyState st, int stage_gen) {
    __shared__ int s_tmp[2];
    __shared__ int s_acc, s_tok, s_err;
    extern __shared__ int s_dyn[];
    const int B = st.B, D = st.D, N = st.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    int* s_active = s_dyn;                       // [B]
    int* s_front = s_active + B;                 // [B]
    int* s_fin = s_front + B;                    // [B]
    int* s_newact = s_fin + B;                   // [B]
    int* s_srclen = s_newact + B;                // [B]
    int* s_nacc_all = s_srclen + B;              // [warps][64]
    int* s_tok_all = s_nacc_all + n_warps * 64;  // [warps][16] predictions of the chosen draft row
    int* col_live = s_tok_all + n_warps * 16;    // [gen_ld]
    int* s_gen = col_live + st.gen_ld;           // [B][gen_ld] when stage_gen
    int* s_nacc = s_nacc_all + warp * 64;
    int* s_tokw = s_tok_all + warp * 16;
    // every lane scores one draft and D + 1 <= 16: the predictions of the chosen row are still in its lane's registers,
    // the token append takes them from there (through shared memory) instead of a second global round trip
    const bool keep_pred = N <= 32 && D + 1 <= 16;
    ACC_TS(0);
    pdl_launch_dependents();
#ifdef TTB_ACC_LATE_LOADS
    pdl_wait();   // A/B build: everything behind the dependency wait
#endif
    // ---- round 1: independent loads, AHEAD of the dependency wait.  Everything read here (control words, live list, fronts,
    // token matrix) was written by the accept kernel of the previous iteration (or by the init kernel), and the first kernel
    // of every iteration (greedy_advance_kernel) releases its dependents only behind its own wait: no kernel of this iteration
    // is scheduled before it has seen that accept kernel complete.  Only the predictions of the classifier (st.pred) need the wait.
    const int done = st.ctrl[CTRL_DONE];
    const int n_active = st.ctrl[CTRL_N_ACTIVE];
    const int Wn = st.ctrl[CTRL_WIDTH];
    const int iter = st.ctrl[CTRL_ITERS];
    const int alleq_pick = st.ctrl[CTRL_ALLEQ_PICK];
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        s_active[b] = st.active[b];
        s_front[b] = st.front[b];
        s_srclen[b] = st.src_len ? st.src_len[b] : 0x7fffffff;
    }
    if (stage_gen) {
        // all loads of a thread in flight before the first store (a plain copy loop waits for every load: seven global
        // round trips for the 27 KB of B = 32, max_len = 200)
        const int total = B * st.gen_ld;
        for (int i0 = 0; i0 < total; i0 += 8 * blockDim.x) {
            int tmp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int idx = i0 + u * blockDim.x + threadIdx.x; tmp[u] = idx < total ? st.gen[idx] : 0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int idx = i0 + u * blockDim.x + threadIdx.x; if (idx < total) s_gen[idx] = tmp[u]; }
        }
    }
    ACC_TS(1);
#ifndef TTB_ACC_LATE_LOADS
    pdl_wait();
#endif
    if (done) return;
    int* G = stage_gen ? s_gen : st.gen;
    if (threadIdx.x == 0) { s_acc = 0; s_tok = 0; s_err = 0; if (st.hist) st.hist[iter] = n_active; }
    __syncthreads();
    ACC_TS(2);
    for (int g = warp; g < n_active; g += n_warps) {
        const int b = s_active[g];
        const int f = s_front[b];
        int best_val = -1, best_first = 0x7fffffff;
        int kept[16];
        for (int n = lane; n < N; n += 32) {
            const int* pr = st.pred + ((long long)g * N + n) * (D + 1);
            const int* dr = st.drafts + ((long long)b * N + n) * D;
            int a = 0;
            bool open = true;
            for (int a0 = 0; a0 < D && open; a0 += 16) {         // 32 independent loads per round
                int dv[16], pv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    dv[u] = (a0 + u < D) ? dr[a0 + u] : -1;
                    pv[u] = (a0 + u <= D) ? pr[a0 + u] : -2;   // pr[D]: the token behind a fully accepted draft (dv = -1 there)
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    if (open && dv[u] == pv[u]) ++a; else open = false;
                    kept[u] = pv[u];
                }
            }
            if (n < 64) s_nacc[n] = (a << 8) | n;   // packed for the tie-break emulation
            if (a > best_val) { best_val = a; best_first = n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // max accepted length, lowest draft index among equals
            const int ov = __shfl_xor_sync(0xffffffffu, best_val, o), oi = __shfl_xor_sync(0xffffffffu, best_first, o);
            if (ov > best_val || (ov == best_val && oi < best_first)) { best_val = ov; best_first = oi; }
        }
        __syncwarp();
        int pick = best_first;
        if (st.tie_break == 0 && N < 64 && N > 1) {
            // a unique maximum is the answer of any selection algorithm; only ties need the emulation
            int ties = 0;
            for (int n = lane; n < N; n += 32) ties += ((s_nacc[n] >> 8) == best_val) ? 1 : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ties += __shfl_xor_sync(0xffffffffu, ties, o);
            if (ties == N) {
                pick = alleq_pick;            // all lengths equal: answer precomputed by the init kernel
            } else if (ties > 1) {
                if (lane == 0) pick = topk1_torch_cpu_packed(s_nacc, N);
                pick = __shfl_sync(0xffffffffu, pick, 0);
            }
        }
        const int a = best_val;
        const int* pr = st.pred + ((long long)g * N + pick) * (D + 1);
        int* row = G + (long long)b * st.gen_ld;
        int* grow = st.gen + (long long)b * st.gen_ld;
        bool fin_l = false;
        if (keep_pred) {
            if (lane == pick) {
#pragma unroll
                for (int u = 0; u < 16; ++u) s_tokw[u] = kept[u];
            }
            __syncwarp();
        }
        for (int j = lane; j <= D; j += 32) {
            const int t = (j <= a) ? (keep_pred ? s_tokw[j] : pr[j]) : st.pad;
            row[f + 1 + j] = t;
            if (stage_gen) grow[f + 1 + j] = t;
            fin_l |= (j <= a) && (t == st.eos);
        }
        const bool fin = __any_sync(0xffffffffu, fin_l);
        __syncwarp();
        if (lane == 0) {
            s_front[b] = f + a + 1;
            st.front[b] = f + a + 1;
            st.sel[g * 4 + 0] = b; st.sel[g * 4 + 1] = f; st.sel[g * 4 + 2] = pick; st.sel[g * 4 + 3] = a;
            if (st.trace) {
                int* tr = st.trace + ((long long)iter * st.B + g) * 4;
                tr[0] = b; tr[1] = a; tr[2] = pick; tr[3] = Wn;
            }
            atomicAdd(&s_acc, a);
            atomicAdd(&s_tok, a + 1);
            s_fin[g] = fin ? 1 : 0;
            if (fin && Wn > st.max_len) s_err = 2;  // the reference cannot store a finished row wider than max_len (:158)
        }
        if (fin && Wn <= st.max_len) {
            long long* o = st.out + (long long)b * st.max_len;
            for (int c = lane; c < Wn; c += 32) o[c] = (c <= f + a + 1) ? row[c] : st.pad;
        }
    }
    ACC_TS(3);
    __syncthreads();
    ACC_TS(4);
    // order-preserving compaction of the live list (boolean masking in the reference): warp 0, ballot prefix sums
    if (warp == 0) {
        int w = 0;
        for (int g0 = 0; g0 < n_active; g0 += 32) {
            const int g = g0 + lane;
            const bool alive = g < n_active && !s_fin[g];
            const unsigned m = __ballot_sync(0xffffffffu, alive);
            if (alive) {
                const int pos = w + __popc(m & ((1u << lane) - 1u));
                const int b = s_active[g];
                s_newact[pos] = b;
                st.active[pos] = b;
                const int fb = s_front[b];
                st.desc[pos] = make_int4(b, fb, G[(long long)b * st.gen_ld + fb], s_srclen[b]);
            }
            w += __popc(m);
        }
        if (lane == 0) {
            st.ctrl[CTRL_N_SEL] = n_active;
            st.ctrl[CTRL_N_ACTIVE] = w;
            st.ctrl[CTRL_ITERS] = iter + 1;
            st.ctrl[CTRL_ACCEPTED] += s_acc;
            st.ctrl[CTRL_TOKENS] += s_tok;
            if (s_err) { st.ctrl[CTRL_ERROR] = s_err; st.ctrl[CTRL_DONE] = 1; st.ctrl[CTRL_N_LEFT] = w; st.ctrl[CTRL_N_ACTIVE] = 0; }
            s_tmp[0] = w;
        }
    }
    __syncthreads();
    ACC_TS(5);
    if (s_err) return;
    const int n_left = s_tmp[0];
    plan_next_iteration(st, s_newact, s_front, G, n_left, Wn, s_tmp, col_live);
    ACC_TS(6);
}
void launch_greedy_accept(const GreedyState& st, cudaStream_t s) {
    const int warps = st.B < 32 ? (st.B < 4 ? 4 : st.B) : 32;
    const size_t base_ints = (size_t)5 * st.B + (size_t)warps * 80 + st.gen_ld;
    const size_t gen_ints = (size_t)st.B * st.gen_ld;
    const int stage_gen = base_ints + gen_ints <= (size_t)ACCEPT_MAX_SMEM_INTS ? 1 : 0;
    const size_t smem = (base_ints + (stage_gen ? gen_ints : 0)) * sizeof(int);
    if (smem > 48 * 1024 && ensure_dyn_smem(greedy_accept_kernel, (int)(ACCEPT_MAX_SMEM_INTS * sizeof(int)))) return;   // the launch below then fails loudly
#ifdef TTB_ACC_TIMELINE
    {
        static int n_launch = 0;
        if (++n_launch == 100) {
            cudaStreamSynchronize(s);
            long long h[16];
            cudaMemcpyFromSymbol(h, g_acc_ts, sizeof(h));
            FILE* f = fopen("gpurun_out/acc_timeline.txt", "w");
            if (f) { for (int i = 0; i < 7; ++i) fprintf(f, "%d %lld\n", i, h[i] - h[1]); fclose(f); }
        }
    }
#endif
    launch_pdl(greedy_accept_kernel, dim3(1), dim3(warps * 32), smem, s, st, stage_gen);
}

// ---- standard greedy decoding (standard_decoding.py:29-56): one token per row and iteration ------------------
// Runs in place of the accept kernel when the loop is used without drafts (N = 1, D = 0).  Every row stays live
// until the first step in which ALL rows predict EOS or PAD (rows that already emitted EOS keep decoding, like in
// the reference), or until the token matrix is full.
__global__ void __launch_bounds__(1024) greedy_std_step_kernel(GreedyState st) {
    pdl_launch_dependents();
    pdl_wait();
    if (st.ctrl[CTRL_DONE]) return;
    const int iter = st.ctrl[CTRL_ITERS];
    const int f = iter;                     // every row holds iter + 1 tokens
    int stop = 1;
    for (int b = threadIdx.x; b < st.B; b += blockDim.x) {
        const int tok = st.pred[b];
        st.gen[(long long)b * st.gen_ld + f + 1] = tok;
        st.front[b] = f + 1;
        st.desc[b] = make_int4(b, f + 1, tok, st.src_len ? st.src_len[b] : 0x7fffffff);
        st.sel[b * 4 + 0] = b; st.sel[b * 4 + 1] = f; st.sel[b * 4 + 2] = 0; st.sel[b * 4 + 3] = 0;
        if (tok != st.eos && tok != st.pad) stop = 0;
    }
    const int all_stop = __syncthreads_and(stop);
    const bool done = all_stop || f + 2 >= st.max_len;
    if (done) {   // export: tokens 0 .. f+1, PAD behind them (the reference's pre-filled token matrix)
        for (long long idx = threadIdx.x; idx < (long long)st.B * st.max_len; idx += blockDim.x) {
            const int b = (int)(idx / st.max_len), c = (int)(idx % st.max_len);
            st.out[idx] = c <= f + 1 ? st.gen[(long long)b * st.gen_ld + c] : st.pad;
        }
    }
    if (threadIdx.x == 0) {
        if (st.hist) st.hist[iter] = st.B;
        st.ctrl[CTRL_N_SEL] = st.B;
        st.ctrl[CTRL_ITERS] = iter + 1;
        st.ctrl[CTRL_TOKENS] += st.B;
        if (done) { st.ctrl[CTRL_DONE] = 1; st.ctrl[CTRL_N_LEFT] = 0; st.ctrl[CTRL_N_ACTIVE] = 0; }
    }
}
void launch_greedy_std_step(const GreedyState& st, cudaStream_t s) {
    launch_pdl(greedy_std_step_kernel, dim3(1), dim3(st.B < 1024 ? ((st.B + 31) / 32 * 32) : 1024), 0, s, st);
}

}  // namespace ttb
