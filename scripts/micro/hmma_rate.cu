// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu ; run on the GPU box: ./scripts/micro/hmma_rate
// Micro-benchmark: throughput and latency of the legacy mma.sync.m16n8k16 (bf16) path on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(float* out, long long* cyc, int iters) {
    uint32_t a[4] = {threadIdx.x, 2, 3, 4};
    float d[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma16816(d[c], a, 0x3F803F80u + c, 0x3F803F80u);
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CHAINS>
void run(int warps_per_sm) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CHAINS><<<148, warps_per_sm * 32>>>(out, cyc, iters);
    k<CHAINS><<<148, warps_per_sm * 32>>>(out, cyc, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_mma_warp = (double)h / (iters * CHAINS);
    double per_sm = per_mma_warp / warps_per_sm;   // cycles per mma per SM
    printf("chains %d warps/SM %2d: %.2f cycles per mma per warp, %.2f cycles per mma per SM, %.0f dense bf16 FMA/clk/SM\n", CHAINS, warps_per_sm,
           per_mma_warp, per_sm, 2048.0 / per_sm);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(1); run<1>(4); run<1>(16);
    run<4>(1); run<4>(4); run<4>(8); run<4>(16);
    run<8>(4); run<8>(16);
    return 0;
}
