// Short-sequence attention kernels (head_dim 16/32/64, sequences of a few hundred tokens).
//
// One thread owns one query: q and the output accumulator live in registers, keys/values of the
// (group, head) are staged tile by tile in shared memory and read as warp-wide broadcasts, the
// softmax is computed online in fp32.  The two entry points differ only in where keys come from:
//   * attention_kernel       : keys/values of a whole group (encoder self-attention, decoder
//                              cross-attention over the cached source memory, full-prefix decoder
//                              self-attention of decode_tgt);
//   * spec_self_attn_kernel  : keys = per-query KV cache of the accepted prefix (shared by all
//                              drafts of the query) + the causal part of the draft row itself.
// Masking mirrors torch: masked keys get probability 0; a query whose keys are all masked
// yields NaN (0/0), exactly like softmax over a row of -inf.
#include "kernels.cuh"

namespace ttb {

template <int HD>
struct QueryState {
    float q[HD];
    float acc[HD];
    float m, l;
    __device__ __forceinline__ void init() {
        m = -INFINITY;
        l = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    }
    // one key/value pair given as fp32 pointers (shared memory broadcast or registers)
    __device__ __forceinline__ void push(const float* __restrict__ kf, const float* __restrict__ vf, float scale) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) s = fmaf(q[d], kf[d], s);
        s *= scale;
        if (s > m) {
            float c = expf(m - s);
            l *= c;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] *= c;
            m = s;
        }
        float p = expf(s - m);
        l += p;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, vf[d], acc[d]);
    }
};

constexpr int ATT_THREADS = 128;
constexpr int ATT_KT = 64;

template <typename ActT, int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const ActT* __restrict__ q, int q_ld, const ActT* __restrict__ k, const ActT* __restrict__ v, int kv_ld,
                 ActT* __restrict__ out, int out_ld, const int* __restrict__ n_groups_dev,
                 int Lq, int Lk, long long kv_group_stride, const int* __restrict__ kvmap,
                 const int* __restrict__ key_tok, int key_tok_stride, int pad_id, int causal, float scale,
                 const int* __restrict__ lk_dev) {
    const int g = blockIdx.z, h = blockIdx.y;
    if (n_groups_dev && g >= *n_groups_dev) return;
    if (lk_dev) { Lk = *lk_dev; kv_group_stride = Lk; key_tok_stride = Lk; }  // graph replay: source length read on device
    const int kvg = kvmap ? kvmap[g] : g;
    const int i = blockIdx.x * ATT_THREADS + threadIdx.x;
    const bool live = i < Lq;
    __shared__ float Ks[ATT_KT][HD];
    __shared__ float Vs[ATT_KT][HD];
    __shared__ int Msk[ATT_KT];

    QueryState<HD> st;
    st.init();
    if (live) {
        const ActT* qp = q + ((long long)g * Lq + i) * q_ld + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) st.q[d] = to_f32(qp[d]);
    }
    const ActT* kbase = k + (long long)kvg * kv_group_stride * kv_ld + h * HD;
    const ActT* vbase = v + (long long)kvg * kv_group_stride * kv_ld + h * HD;
    // keys beyond the last query of this block are never needed under the causal mask
    const int k_end = causal ? min(Lk, blockIdx.x * ATT_THREADS + ATT_THREADS) : Lk;
    for (int j0 = 0; j0 < k_end; j0 += ATT_KT) {
        const int nk = min(ATT_KT, k_end - j0);
        for (int idx = threadIdx.x; idx < nk * HD; idx += ATT_THREADS) {
            int j = idx / HD, d = idx % HD;
            Ks[j][d] = to_f32(kbase[(long long)(j0 + j) * kv_ld + d]);
            Vs[j][d] = to_f32(vbase[(long long)(j0 + j) * kv_ld + d]);
        }
        for (int j = threadIdx.x; j < nk; j += ATT_THREADS)
            Msk[j] = key_tok ? (key_tok[(long long)kvg * key_tok_stride + j0 + j] == pad_id) : 0;
        __syncthreads();
        if (live) {
            for (int j = 0; j < nk; ++j) {
                if (Msk[j]) continue;
                if (causal && j0 + j > i) break;
                st.push(Ks[j], Vs[j], scale);
            }
        }
        __syncthreads();
    }
    if (live) {
        ActT* op = out + ((long long)g * Lq + i) * out_ld + h * HD;
        const float inv = 1.0f / st.l;  // l == 0 -> inf -> 0 * inf = NaN, like torch on a fully masked row
#pragma unroll
        for (int d = 0; d < HD; ++d) op[d] = from_f32<ActT>(st.l == 0.f ? __int_as_float(0x7fc00000) : st.acc[d] * inv);
    }
}

template <typename ActT>
void launch_attention(const ActT* q, int q_ld, const ActT* k, const ActT* v, int kv_ld,
                      ActT* out, int out_ld, int n_groups_max, const int* n_groups_dev,
                      int Lq, int Lk, long long kv_group_stride, const int* kvmap,
                      const int* key_tok, int key_tok_stride, int pad_id, bool causal,
                      int heads, int head_dim, cudaStream_t s, const int* lk_dev) {
    if (n_groups_max <= 0 || Lq <= 0) return;
    dim3 grid((Lq + ATT_THREADS - 1) / ATT_THREADS, heads, n_groups_max);
    const float scale = 1.0f / sqrtf((float)head_dim);
#define TTB_ATT(HDV)                                                                                         \
    attention_kernel<ActT, HDV><<<grid, ATT_THREADS, 0, s>>>(q, q_ld, k, v, kv_ld, out, out_ld, n_groups_dev, \
                                                             Lq, Lk, kv_group_stride, kvmap, key_tok,        \
                                                             key_tok_stride, pad_id, causal ? 1 : 0, scale, lk_dev)
    if (head_dim == 16) TTB_ATT(16);
    else if (head_dim == 32) TTB_ATT(32);
    else if (head_dim == 64) TTB_ATT(64);
#undef TTB_ATT
}
template void launch_attention<float>(const float*, int, const float*, const float*, int, float*, int, int, const int*, int, int, long long, const int*, const int*, int, int, bool, int, int, cudaStream_t, const int*);
template void launch_attention<__nv_bfloat16>(const __nv_bfloat16*, int, const __nv_bfloat16*, const __nv_bfloat16*, int, __nv_bfloat16*, int, int, const int*, int, int, long long, const int*, const int*, int, int, bool, int, int, cudaStream_t, const int*);

// ------------------------------------------------------------------------------------------------
constexpr int SPEC_THREADS = 256;

template <typename ActT, int HD>
__global__ void __launch_bounds__(SPEC_THREADS)
spec_self_attn_kernel(const ActT* __restrict__ qkv, int qkv_ld, const ActT* __restrict__ kcache,
                      const ActT* __restrict__ vcache, long long cache_query_stride, int cache_ld,
                      ActT* __restrict__ out, int out_ld, const int* __restrict__ n_active_dev,
                      const int* __restrict__ active, const int* __restrict__ front,
                      const int* __restrict__ gen, int gen_ld, int pad_id, int N, int D, int E, float scale) {
    const int g = blockIdx.y, h = blockIdx.x;
    if (g >= *n_active_dev) return;
    const int b = active[g];
    const int f = front[b];
    const int rows = N * (D + 1);
    __shared__ float Ks[ATT_KT][HD];
    __shared__ float Vs[ATT_KT][HD];
    __shared__ int Msk[ATT_KT];
    const ActT* kb = kcache + (long long)b * cache_query_stride + h * HD;
    const ActT* vb = vcache + (long long)b * cache_query_stride + h * HD;
    const bool first_new_masked = gen[(long long)b * gen_ld + f] == pad_id;

    for (int r0 = 0; r0 < rows; r0 += SPEC_THREADS) {
        const int r = r0 + threadIdx.x;
        const bool live = r < rows;
        const long long tok = (long long)g * rows + r;
        QueryState<HD> st;
        st.init();
        if (live) {
            const ActT* qp = qkv + tok * qkv_ld + h * HD;
#pragma unroll
            for (int d = 0; d < HD; ++d) st.q[d] = to_f32(qp[d]);
        }
        // (1) accepted prefix, shared by all drafts of the query: staged through shared memory
        for (int j0 = 0; j0 < f; j0 += ATT_KT) {
            const int nk = min(ATT_KT, f - j0);
            for (int idx = threadIdx.x; idx < nk * HD; idx += SPEC_THREADS) {
                int j = idx / HD, d = idx % HD;
                Ks[j][d] = to_f32(kb[(long long)(j0 + j) * cache_ld + d]);
                Vs[j][d] = to_f32(vb[(long long)(j0 + j) * cache_ld + d]);
            }
            for (int j = threadIdx.x; j < nk; j += SPEC_THREADS)
                Msk[j] = gen[(long long)b * gen_ld + j0 + j] == pad_id;
            __syncthreads();
            if (live) {
                for (int j = 0; j < nk; ++j) {
                    if (Msk[j]) continue;
                    st.push(Ks[j], Vs[j], scale);
                }
            }
            __syncthreads();
        }
        // (2) the draft row itself (causal), straight from the freshly projected K/V (L2 resident)
        if (live) {
            const int i = r % (D + 1);
            const long long row0 = tok - i;
            for (int ii = 0; ii <= i; ++ii) {
                if (ii == 0 && first_new_masked) continue;
                const ActT* kp = qkv + (row0 + ii) * qkv_ld + E + h * HD;
                const ActT* vp = kp + E;
                float kf[HD], vf[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) { kf[d] = to_f32(kp[d]); vf[d] = to_f32(vp[d]); }
                st.push(kf, vf, scale);
            }
            ActT* op = out + tok * out_ld + h * HD;
            const float inv = 1.0f / st.l;
#pragma unroll
            for (int d = 0; d < HD; ++d) op[d] = from_f32<ActT>(st.l == 0.f ? __int_as_float(0x7fc00000) : st.acc[d] * inv);
        }
    }
}

template <typename ActT>
void launch_spec_self_attention(const ActT* qkv, int qkv_ld, const ActT* kcache, const ActT* vcache,
                                long long cache_query_stride, int cache_ld, ActT* out, int out_ld,
                                int B_max, const int* n_active_dev, const int* active, const int* front,
                                const int* gen, int gen_ld, int pad_id, int N, int D,
                                int heads, int head_dim, int max_cache_len, cudaStream_t s) {
    (void)max_cache_len;
    if (B_max <= 0) return;
    dim3 grid(heads, B_max);
    const float scale = 1.0f / sqrtf((float)head_dim);
    const int E = heads * head_dim;
#define TTB_SPEC(HDV)                                                                                        \
    spec_self_attn_kernel<ActT, HDV><<<grid, SPEC_THREADS, 0, s>>>(qkv, qkv_ld, kcache, vcache,             \
        cache_query_stride, cache_ld, out, out_ld, n_active_dev, active, front, gen, gen_ld, pad_id, N, D, E, scale)
    if (head_dim == 16) TTB_SPEC(16);
    else if (head_dim == 32) TTB_SPEC(32);
    else if (head_dim == 64) TTB_SPEC(64);
#undef TTB_SPEC
}
template void launch_spec_self_attention<float>(const float*, int, const float*, const float*, long long, int, float*, int, int, const int*, const int*, const int*, const int*, int, int, int, int, int, int, int, cudaStream_t);
template void launch_spec_self_attention<__nv_bfloat16>(const __nv_bfloat16*, int, const __nv_bfloat16*, const __nv_bfloat16*, long long, int, __nv_bfloat16*, int, int, const int*, const int*, const int*, const int*, int, int, int, int, int, int, int, cudaStream_t);

}  // namespace ttb
