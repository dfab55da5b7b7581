// Blackwell-native attention for the decoding loops (bf16 path, head_dim 32): S = Q K^T and O = P V on tcgen05.mma with
// the score tile in tensor memory.
//
// One CTA of 128 threads per (group, head, block of 128 query rows); thread t owns query row t = TMEM lane t.
//   * shared keys (accepted-prefix KV cache of a query / cached source memory): K rows are staged K-major in shared memory
//     (rows of 128 bytes, 128-byte swizzle, only the first 64 bytes = 32 dims of a row are used), V comes from a cache that
//     is kept TRANSPOSED ([dim][position]), so the second product's B operand (N = 32 dims, K = keys) is K-major as well
//     and needs no transposing instruction: four 64-key blocks of 32 x 128 bytes;
//   * S[128 x Nk] = Q K^T: two tcgen05.mma (M = 128, N = Nk rounded to 16, K = 16) into Nk fp32 TMEM columns;
//   * softmax by the row's own thread straight from tensor memory (tcgen05.ld, 32 columns at a time): pass 1 row maximum,
//     pass 2 exp2 / row sum / bf16 rounding of P, which is written back IN PLACE over the consumed score columns with
//     tcgen05.st (two keys per 32-bit column) and becomes the A operand of the second product (TS form);
//   * O[128 x 32] += P V: Nk / 16 tcgen05.mma (A from tensor memory, B = V^T blocks), accumulator in 32 more TMEM columns;
//   * "private" keys of speculative self-attention (the freshly projected K/V of the query's own draft row, at most 16,
//     causal inside the row) never form a dense tile: the row's thread scores them on the SIMT pipes while the tensor core
//     works on the prefix, and both parts share one softmax normalisation.
// TMEM columns are allocated per CTA for what its key count needs (64 / 128 / 256), so several CTAs share an SM.
// Masking mirrors torch: masked keys get probability 0, a fully masked query row yields NaN.
// Replaces amma::attn_mma_kernel (mma.sync) inside the decoding loops; the encoder / full-sequence decoder paths and every
// shape outside the limits below stay on that kernel.
#include "kernels.cuh"

namespace ttb {
namespace atc {

constexpr int HD = 32, ROWS = 128, THREADS = 128;
constexpr int KMAX = 224;            // shared keys per group (TMEM: KMAX + 32 accumulator columns = 256)
constexpr int RLMAX = 16;            // tokens per draft row (private keys per query)
constexpr int PRIV = ROWS + RLMAX;   // private key rows a 128-row query block can see (its draft rows, the first one from its start)
constexpr int PPITCH = 80;           // bytes per staged private K / V row (64 + 16: rows 20 words apart -> conflict-free 16-byte reads)
constexpr int OFF_Q = 0, OFF_K = ROWS * 128, OFF_VT = OFF_K + KMAX * 128, OFF_PK = OFF_VT + 4 * 4096, OFF_PV = OFF_PK + PRIV * PPITCH,
              OFF_BIAS = OFF_PV + PRIV * PPITCH;   // [8] key-mask words (bit j of word c: key 32 c + j visible), barriers, TMEM slot
constexpr int SMEM = OFF_BIAS + 128 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "ATC_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra ATC_WAIT_DONE;\n\t"
        "bra ATC_WAIT_LOOP;\n\t"
        "ATC_WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major tile, 128-byte swizzle: rows of 128 bytes, 8-row atoms 1024 bytes apart (same descriptor as the GEMM kernels)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {   // D = f32, A = B = bf16, K-major, N >> 3 at bit 17, M >> 4 at bit 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {   // ex2.approx: exact 0 for -inf, 1 for 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct Params {
    const __nv_bfloat16* q; int q_ld;                       // query rows: token (g * Lq + r), head offset h * 32
    const __nv_bfloat16* k; int k_ld; long long k_group_stride;      // shared keys, row-major: k + kvg * stride + j * k_ld + h * 32
    const __nv_bfloat16* vt; long long vt_group_stride; int vt_pitch; // shared values, transposed: vt + kvg * stride + (h * 32 + n) * pitch + j
    __nv_bfloat16* out; int out_ld;
    const int* n_groups_dev;
    int Lq;
    const int4* desc;                                       // per live slot {kv group, front, token at front, key bound}
    const int* key_tok; int key_tok_stride; int pad_id;     // key j of group kvg is masked when key_tok[kvg * stride + j] == pad_id
    const int* lk_dev;                                      // cross-attention: key count of the batch (device word, graph replay)
    float scale_log2e;
    int spec;                                               // 1: keys = cache prefix [0, front) + the query's own draft row
    const __nv_bfloat16* newk; const __nv_bfloat16* newv; int new_ld; int row_len;
};

__global__ void __launch_bounds__(THREADS) attn_tc_kernel(Params p) {
    const int h = blockIdx.x, g = blockIdx.y, mt = blockIdx.z;
    pdl_launch_dependents();
    // the descriptor table, the live count and the cache rows older than this iteration were written before the
    // iteration's first kernel (a fully serialised launch): they may be read ahead of the programmatic dependency
    const int4 dsc = p.desc[g];
    const int dynLk = p.lk_dev ? *p.lk_dev : 0;
    const int n_live = p.n_groups_dev ? *p.n_groups_dev : 0x7fffffff;
    if (g >= n_live || mt * ROWS >= p.Lq) return;
    extern __shared__ uint8_t atc_raw[];
    const uint32_t raw = smem_u32(atc_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;           // 128-byte swizzle atoms need 1024-byte alignment
    uint8_t* gen = atc_raw + (base - raw);
    uint32_t* kmask = reinterpret_cast<uint32_t*>(gen + OFF_BIAS);   // bit j of word c: key 32 c + j is visible
    const uint32_t bar_s = base + OFF_BIAS + 64, bar_o = bar_s + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + OFF_BIAS + 64 + 16);
    const int t = threadIdx.x, warp = t >> 5;

    const int kvg = dsc.x;
    const int Lk = p.spec ? min(dsc.y, KMAX) : min(min(dynLk, dsc.w), KMAX);
    const int Nk = (Lk + 15) & ~15;                          // MMA N / K extent (keys Lk .. Nk-1 are zero rows with bias -inf)
    const uint32_t cols = Nk + 32 <= 64 ? 64u : (Nk + 32 <= 128 ? 128u : 256u);
    if (warp == 0) {
        if (t == 0) {
            mbar_init(bar_s, 1);
            mbar_init(bar_o, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- staging of the shared keys ------------------------------------------------------------------------------
    // cross-attention inside a replayed graph: the batch's source length (rows per group of the K buffer and of the key
    // tokens) is read on the device so that one captured graph serves batches of any length
    const long long k_group_stride = p.lk_dev ? (long long)dynLk * p.k_ld : p.k_group_stride;
    const int key_tok_stride = p.lk_dev ? dynLk : p.key_tok_stride;
    const __nv_bfloat16* kbase = p.k + (long long)kvg * k_group_stride + h * HD;
    const __nv_bfloat16* vtbase = p.vt + (long long)kvg * p.vt_group_stride + (long long)(h * HD) * p.vt_pitch;
    const int* key_tok = p.key_tok ? p.key_tok + (long long)kvg * key_tok_stride : nullptr;
    const int RL = p.spec ? p.row_len : 1;
    // rows appended to the cache by THIS iteration's first kernel (at most one draft row's worth) wait for the dependency
    const int n_pre = p.spec ? max(0, Lk - RL) & ~7 : Lk;
    auto stage_k = [&](int j0, int j1) {                     // K rows [j0, j1): 4 chunks of 16 bytes each
        for (int idx = j0 * 4 + t; idx < j1 * 4; idx += THREADS) {
            const int j = idx >> 2, c = idx & 3;
            cp_async16(base + OFF_K + j * 128 + ((c ^ (j & 7)) << 4), kbase + (long long)j * p.k_ld + c * 8);
        }
    };
    auto stage_vt = [&](int c0, int c1) {                    // key chunks (8 keys = 16 bytes) [c0, c1) of all 32 dims
        const int nch = c1 - c0;
        for (int idx = t; idx < nch * HD; idx += THREADS) {
            const int n = idx / nch, kc = c0 + idx % nch;
            cp_async16(base + OFF_VT + (kc >> 3) * 4096 + n * 128 + (((kc & 7) ^ (n & 7)) << 4), vtbase + (long long)n * p.vt_pitch + kc * 8);
        }
    };
    stage_k(0, n_pre);
    stage_vt(0, n_pre >> 3);
    for (int j = Lk + (t >> 2); j < Nk; j += THREADS / 4) {  // zero rows behind the last key (finite operands for the MMA)
        const int c = t & 3;
        *reinterpret_cast<uint4*>(gen + OFF_K + j * 128 + ((c ^ (j & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    }
    pdl_wait();
    stage_k(n_pre, Lk);
    stage_vt(n_pre >> 3, Nk >> 3);
    for (int j0 = 0; j0 < 256; j0 += THREADS) {           // visible-key mask, one ballot per 32 keys
        const int j = j0 + t;
        const unsigned m = __ballot_sync(0xffffffffu, j < Lk && !(key_tok && key_tok[j] == p.pad_id));
        if ((t & 31) == 0) kmask[j >> 5] = m;
    }
    // private keys: the K / V rows of the draft rows this block's queries belong to (first row: start of the draft row that
    // contains the block's first query), staged once per CTA; a query then reads its <= 16 rows from shared memory
    const int pr0 = p.spec ? ((mt * ROWS) / RL) * RL : 0;
    const int n_pr = p.spec ? min(p.Lq, mt * ROWS + ROWS) - pr0 : 0;
    for (int idx = t; idx < n_pr * 8; idx += THREADS) {
        const int j = idx >> 3, c = idx & 7;
        const long long row = (long long)g * p.Lq + pr0 + j;
        if (c < 4) cp_async16(base + OFF_PK + j * PPITCH + c * 16, p.newk + row * p.new_ld + h * HD + c * 8);
        else cp_async16(base + OFF_PV + j * PPITCH + (c - 4) * 16, p.newv + row * p.new_ld + h * HD + (c - 4) * 8);
    }
    {   // Q tile: rows mt*128 .. +127 (rows behind Lq are zero)
        const long long row0 = (long long)g * p.Lq + mt * ROWS;
        for (int idx = t; idx < ROWS * 4; idx += THREADS) {
            const int r = idx >> 2, c = idx & 3;
            const uint32_t dst = base + OFF_Q + r * 128 + ((c ^ (r & 7)) << 4);
            if (mt * ROWS + r < p.Lq) cp_async16(dst, p.q + (row0 + r) * p.q_ld + h * HD + c * 8);
            else *reinterpret_cast<uint4*>(gen + OFF_Q + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_wait_all();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_o = tmem_base + cols - 32;

    if (t == 0 && Nk > 0) {          // S = Q K^T
        const uint32_t idesc = idesc_bf16(ROWS, Nk);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
            mma_ss(tmem_base, desc_sw128(base + OFF_Q + kk * 32), desc_sw128(base + OFF_K + kk * 32), idesc, kk);
        mma_commit(bar_s);
    }

    // ---- private keys (own draft row) on the SIMT pipes while the tensor core works -------------------------------
    const int r = mt * ROWS + t;
    const bool row_live = r < p.Lq;
    float sp[RLMAX];
    int n_priv = 0;
    float m_run = -INFINITY;
    const long long grow0 = (long long)g * p.Lq;
    if (p.spec && row_live) {
        const int n0 = (r / RL) * RL, i = r - n0;
        n_priv = i + 1;
        float qf[HD];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(gen + OFF_Q + t * 128 + ((c ^ (t & 7)) << 4));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                qf[c * 8 + 2 * e] = __uint_as_float(w[e] << 16);
                qf[c * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
            }
        }
        const bool first_masked = dsc.z == p.pad_id;
        const uint8_t* pk = gen + OFF_PK + (n0 - pr0) * PPITCH;
#pragma unroll
        for (int j = 0; j < RLMAX; ++j) {
            sp[j] = -INFINITY;
            if (j < RL) {                                   // uniform bound: every lane walks the whole draft row, masks by position
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 u = *reinterpret_cast<const uint4*>(pk + j * PPITCH + c * 16);
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc = fmaf(qf[c * 8 + 2 * e], __uint_as_float(w[e] << 16), acc);
                        acc = fmaf(qf[c * 8 + 2 * e + 1], __uint_as_float(w[e] & 0xffff0000u), acc);
                    }
                }
                if (j < n_priv && !(j == 0 && first_masked)) {
                    sp[j] = acc * p.scale_log2e;
                    m_run = fmaxf(m_run, sp[j]);
                }
            }
        }
    }

    // ---- softmax over the shared keys, straight from tensor memory ---------------------------------------------------
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float l_run = 0.f;
    if (Nk > 0) {
        mbar_wait(bar_s, 0);
        fence_after();
        // pass 1: row maximum; the load of the next 32 columns is in flight while the current ones are reduced
        uint32_t sa[32], sb[32];
        tmem_ld32_nowait(tmem_base + lane_base, sa);
#pragma unroll 1
        for (int c0 = 0; c0 < Nk; c0 += 64) {
            tmem_ld_wait();
            if (c0 + 32 < Nk) tmem_ld32_nowait(tmem_base + lane_base + c0 + 32, sb);
            {
                const uint32_t mk = kmask[c0 >> 5];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if ((mk >> j) & 1u) m_run = fmaxf(m_run, __uint_as_float(sa[j]) * p.scale_log2e);
            }
            if (c0 + 32 < Nk) {
                tmem_ld_wait();
                if (c0 + 64 < Nk) tmem_ld32_nowait(tmem_base + lane_base + c0 + 64, sa);
                const uint32_t mk = kmask[(c0 >> 5) + 1];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if ((mk >> j) & 1u) m_run = fmaxf(m_run, __uint_as_float(sb[j]) * p.scale_log2e);
            }
        }
    }
    const float m_use = m_run == -INFINITY ? 0.f : m_run;
    if (Nk > 0) {
        // pass 2: probabilities, rounded to bf16, back into the (already consumed) score columns
        uint32_t sa[32], sb[32];
        tmem_ld32_nowait(tmem_base + lane_base, sa);
        auto emit = [&](const uint32_t (&sv)[32], int c0) {
            const uint32_t mk = kmask[c0 >> 5];
            uint32_t pw[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float e0 = ((mk >> j) & 1u) ? fast_exp2(fmaf(__uint_as_float(sv[j]), p.scale_log2e, -m_use)) : 0.f;
                const float e1 = ((mk >> (j + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(sv[j + 1]), p.scale_log2e, -m_use)) : 0.f;
                l_run += e0 + e1;
                pw[j >> 1] = pack_bf16(e0, e1);
            }
            tmem_st16(tmem_base + lane_base + (c0 >> 1), pw);
        };
#pragma unroll 1
        for (int c0 = 0; c0 < Nk; c0 += 64) {
            tmem_ld_wait();
            if (c0 + 32 < Nk) tmem_ld32_nowait(tmem_base + lane_base + c0 + 32, sb);
            emit(sa, c0);
            if (c0 + 32 < Nk) {
                tmem_ld_wait();
                // the store of P chunk c0/32 + 1 lands in columns [c0/2 + 16, c0/2 + 32), all below c0 + 64: the load ahead is safe
                if (c0 + 64 < Nk) tmem_ld32_nowait(tmem_base + lane_base + c0 + 64, sa);
                emit(sb, c0 + 32);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    if (t == 0 && Nk > 0) {          // O = P V
        fence_after();
        const uint32_t idesc = idesc_bf16(ROWS, HD);
        for (int kk = 0; kk < Nk / 16; ++kk)
            mma_ts(tm_o, tmem_base + kk * 8, desc_sw128(base + OFF_VT + (kk >> 2) * 4096 + (kk & 3) * 32), idesc, kk);
        mma_commit(bar_o);
    }

    // ---- private part of P V (SIMT) ------------------------------------------------------------------------------------
    float of[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) of[d] = 0.f;
    if (p.spec && row_live) {
        const int n0 = (r / RL) * RL;
        const uint8_t* pv = gen + OFF_PV + (n0 - pr0) * PPITCH;
#pragma unroll
        for (int j = 0; j < RLMAX; ++j) {
            if (j < n_priv) {                               // rows behind the query's own position may lie outside the staged range
                const float e = fast_exp2(sp[j] - m_use);     // sp = -inf (masked) -> 0
                l_run += e;
                const float eb = bf16_round(e);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 u = *reinterpret_cast<const uint4*>(pv + j * PPITCH + c * 16);
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2) {
                        of[c * 8 + 2 * e2] = fmaf(eb, __uint_as_float(w[e2] << 16), of[c * 8 + 2 * e2]);
                        of[c * 8 + 2 * e2 + 1] = fmaf(eb, __uint_as_float(w[e2] & 0xffff0000u), of[c * 8 + 2 * e2 + 1]);
                    }
                }
            }
        }
    }
    if (Nk > 0) {
        mbar_wait(bar_o, 0);
        fence_after();
        uint32_t ov[32];
        tmem_ld32(tm_o + lane_base, ov);
#pragma unroll
        for (int d = 0; d < HD; ++d) of[d] += __uint_as_float(ov[d]);
    }
    if (row_live) {
        const float nanv = __int_as_float(0x7fc00000);
        const float inv = 1.0f / l_run;
        uint4* op = reinterpret_cast<uint4*>(p.out + (grow0 + r) * p.out_ld + h * HD);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 u;
            if (l_run == 0.f) {
                u.x = u.y = u.z = u.w = pack_bf16(nanv, nanv);
            } else {
                u.x = pack_bf16(of[c * 8 + 0] * inv, of[c * 8 + 1] * inv);
                u.y = pack_bf16(of[c * 8 + 2] * inv, of[c * 8 + 3] * inv);
                u.z = pack_bf16(of[c * 8 + 4] * inv, of[c * 8 + 5] * inv);
                u.w = pack_bf16(of[c * 8 + 6] * inv, of[c * 8 + 7] * inv);
            }
            op[c] = u;
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
}

// out[(g * E + c) * pitch + j] = in[(g * L + j) * ld + c]  (V of the cached source memory -> transposed cache; once per batch)
__global__ void transpose_v_kernel(const __nv_bfloat16* __restrict__ in, int ld, int L, int E, __nv_bfloat16* __restrict__ out, int pitch) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int g = blockIdx.z, j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int j = j0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (j < L && c < E) ? in[((long long)g * L + j) * ld + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, j = j0 + threadIdx.x;
        if (c < E && j < pitch) out[((long long)g * E + c) * pitch + j] = tile[threadIdx.x][i];
    }
}

}  // namespace atc

bool attention_tc_supported(int head_dim, int Lq, int max_shared_keys, int row_len) {
    // OPT-IN (TTB_ATTN_TC=1), read once per generate(): measured on B200 inside the greedy loop (bs 32, 253 query rows and
    // <= 212 keys per (query, head)) the kernel is correct (tests/test_gpu_bench_configs.py) but slower than the mma.sync
    // kernel -- 25 / 19 us per self / cross launch against 13.5 / 13.2 us: at head_dim 32 the two products are ~1 % of the
    // launch, the rest is a chain of dependent latencies (stage -> barrier -> MMA -> mbarrier -> two tensor-memory passes ->
    // barrier -> MMA -> mbarrier -> store) that four warps per CTA and 2-3 CTAs per SM (tensor-memory and shared-memory
    // bound) cannot hide, whereas the mma.sync kernel keeps 16 independent warps per SM busy.
    const char* v = getenv("TTB_ATTN_TC");
    const bool on = v && v[0] == '1';
    return on && head_dim == atc::HD && Lq >= 1 && max_shared_keys <= atc::KMAX && row_len <= atc::RLMAX;
}

void launch_transpose_v(const __nv_bfloat16* in, int ld, int groups, int L, int E, __nv_bfloat16* out, int pitch, cudaStream_t s) {
    if (groups <= 0 || L <= 0) return;
    dim3 grid((pitch + 31) / 32, (E + 31) / 32, groups);
    atc::transpose_v_kernel<<<grid, dim3(32, 8), 0, s>>>(in, ld, L, E, out, pitch);
}

static int launch_tc(const atc::Params& p, int heads, int n_groups_max, cudaStream_t s) {
    if (int rc = ensure_dyn_smem(atc::attn_tc_kernel, atc::SMEM)) return rc;
    dim3 grid(heads, n_groups_max, (p.Lq + atc::ROWS - 1) / atc::ROWS);
    launch_pdl(atc::attn_tc_kernel, grid, dim3(atc::THREADS), (size_t)atc::SMEM, s, p);
    return 0;
}

int launch_cross_attention_tc(const __nv_bfloat16* q, int q_ld, const __nv_bfloat16* k, int k_ld, long long k_group_stride,
                              const __nv_bfloat16* vt, long long vt_group_stride, int vt_pitch, __nv_bfloat16* out, int out_ld,
                              int n_groups_max, const int* n_groups_dev, int Lq, const int* key_tok, int key_tok_stride, int pad_id,
                              const int* lk_dev, int heads, const int4* desc, cudaStream_t s) {
    if (n_groups_max <= 0 || Lq <= 0) return 0;
    atc::Params p{};
    p.q = q; p.q_ld = q_ld; p.k = k; p.k_ld = k_ld; p.k_group_stride = k_group_stride;
    p.vt = vt; p.vt_group_stride = vt_group_stride; p.vt_pitch = vt_pitch; p.out = out; p.out_ld = out_ld;
    p.n_groups_dev = n_groups_dev; p.Lq = Lq; p.desc = desc; p.key_tok = key_tok; p.key_tok_stride = key_tok_stride; p.pad_id = pad_id;
    p.lk_dev = lk_dev; p.scale_log2e = 1.4426950408889634f / sqrtf((float)atc::HD); p.spec = 0; p.row_len = 1;
    return launch_tc(p, heads, n_groups_max, s);
}

int launch_spec_self_attention_tc(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* kcache, long long k_query_stride, int k_ld,
                                  const __nv_bfloat16* vtcache, long long vt_query_stride, int vt_pitch, __nv_bfloat16* out, int out_ld,
                                  int B_max, const int* n_active_dev, const int* gen, int gen_ld, int pad_id, int N, int D, int heads,
                                  const int4* desc, cudaStream_t s) {
    if (B_max <= 0) return 0;
    const int E = heads * atc::HD;
    atc::Params p{};
    p.q = qkv; p.q_ld = qkv_ld; p.k = kcache; p.k_ld = k_ld; p.k_group_stride = k_query_stride;
    p.vt = vtcache; p.vt_group_stride = vt_query_stride; p.vt_pitch = vt_pitch; p.out = out; p.out_ld = out_ld;
    p.n_groups_dev = n_active_dev; p.Lq = N * (D + 1); p.desc = desc; p.key_tok = gen; p.key_tok_stride = gen_ld; p.pad_id = pad_id;
    p.lk_dev = nullptr; p.scale_log2e = 1.4426950408889634f / sqrtf((float)atc::HD); p.spec = 1;
    p.newk = qkv + E; p.newv = qkv + 2 * E; p.new_ld = qkv_ld; p.row_len = D + 1;
    return launch_tc(p, heads, B_max, s);
}

}  // namespace ttb
