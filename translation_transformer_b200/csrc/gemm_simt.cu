// fp32 SIMT GEMM: C[M,N] = A[M,K] * W[N,K]^T + bias, optional ReLU.  This is the exact-precision
// path (precision = fp32); the bf16 path uses the tcgen05 kernel in gemm_tcgen05.cu.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, operands staged K-major in shared memory.
#include "kernels.cuh"

namespace ttb {

constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

template <typename OutT>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                const float* __restrict__ bias, OutT* __restrict__ C, int ldc,
                RowCount rows, int N, int K, int relu) {
    const int M = rows.live();
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    if (m0 >= M) return;
    __shared__ float As[GB_K][GB_M + 4];
    __shared__ float Ws[GB_K][GB_N + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;   // loader: row 0..63, k offset 0,4,8,12
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += GB_K) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
        if (m0 + lrow < M) a = *reinterpret_cast<const float4*>(A + (long long)(m0 + lrow) * lda + k0 + lk);
        if (n0 + lrow < N) w = *reinterpret_cast<const float4*>(W + (long long)(n0 + lrow) * K + k0 + lk);
        As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
        Ws[lk + 0][lrow] = w.x; Ws[lk + 1][lrow] = w.y; Ws[lk + 2][lrow] = w.z; Ws[lk + 3][lrow] = w.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GB_K; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[(long long)m * ldc + n] = from_f32<OutT>(v);
        }
    }
}

template <typename OutT>
void launch_gemm_f32(const float* A, int lda, const float* W, const float* bias, OutT* C, int ldc,
                     RowCount rows, int N, int K, bool relu, cudaStream_t s) {
    if (rows.max_rows <= 0 || N <= 0) return;
    dim3 grid((N + GB_N - 1) / GB_N, (rows.max_rows + GB_M - 1) / GB_M);
    gemm_f32_kernel<OutT><<<grid, 256, 0, s>>>(A, lda, W, bias, C, ldc, rows, N, K, relu ? 1 : 0);
}
template void launch_gemm_f32<float>(const float*, int, const float*, const float*, float*, int, RowCount, int, int, bool, cudaStream_t);
template void launch_gemm_f32<__nv_bfloat16>(const float*, int, const float*, const float*, __nv_bfloat16*, int, RowCount, int, int, bool, cudaStream_t);

}  // namespace ttb
