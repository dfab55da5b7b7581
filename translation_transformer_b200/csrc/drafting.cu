// Draft construction on device (mirror of /root/reference/src/utils/drafting.py:5-65).
// One block per source row: the row (virtually right-padded to N + D - 1 tokens) is staged in
// shared memory with a prefix sum of its service-token flags, the number of windows free of
// EOS/PAD is counted, N window offsets are spread over them with the reference's float32
// arithmetic (`steps * ((take_from - 1) / max(N - 1, 1))`, truncated), and the windows are
// gathered with EOS/PAD replaced.
#include "kernels.cuh"

namespace ttb {

__global__ void make_drafts_kernel(const int* __restrict__ src, int src_ld, int L, int Lp, int D, int N,
                                   int eos, int pad, int replace, int* __restrict__ out) {
    extern __shared__ int sm[];
    int* tok = sm;             // [Lp]
    int* pre = sm + Lp;        // [Lp + 1] exclusive prefix of service flags
    __shared__ int n_clean_s;
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < Lp; j += blockDim.x) tok[j] = j < L ? src[(long long)b * src_ld + j] : pad;
    if (threadIdx.x == 0) n_clean_s = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        pre[0] = 0;
        for (int j = 0; j < Lp; ++j) {
            run += (tok[j] == eos || tok[j] == pad) ? 1 : 0;
            pre[j + 1] = run;
        }
    }
    __syncthreads();
    const int n_win = Lp - D + 1;
    int local = 0;
    for (int w = threadIdx.x; w < n_win; w += blockDim.x) local += (pre[w + D] - pre[w]) == 0 ? 1 : 0;
    if (local) atomicAdd(&n_clean_s, local);
    __syncthreads();
    const int take_from = max(n_clean_s, N);
    const float step = __fdiv_rn((float)(take_from - 1), (float)max(N - 1, 1));
    for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
        const int n = idx / D, d = idx % D;
        const int start = (int)__fmul_rn((float)n, step);
        int t = tok[start + d];
        if (t == eos || t == pad) t = replace;
        out[((long long)b * N + n) * D + d] = t;
    }
}

void launch_make_drafts(const int* src, int src_ld, int B, int L, int Deff, int N, int eos, int pad, int replace,
                        int* out, cudaStream_t s) {
    if (B <= 0) return;
    const int Lp = max(L, N + Deff - 1);
    size_t smem = (size_t)(2 * Lp + 1) * sizeof(int);
    make_drafts_kernel<<<B, 256, smem, s>>>(src, src_ld, L, Lp, Deff, N, eos, pad, replace, out);
}

}  // namespace ttb
