// placeholder replaced below
#include "kernels.cuh"
namespace ttb {
template <typename OutT>
int launch_gemm_bf16_tc(const __nv_bfloat16*, int, const __nv_bfloat16*, const float*, OutT*, int, RowCount, int, int, bool, cudaStream_t) {
    set_last_error("bf16 tcgen05 GEMM not built yet");
    return 3;
}
template int launch_gemm_bf16_tc<float>(const __nv_bfloat16*, int, const __nv_bfloat16*, const float*, float*, int, RowCount, int, int, bool, cudaStream_t);
template int launch_gemm_bf16_tc<__nv_bfloat16>(const __nv_bfloat16*, int, const __nv_bfloat16*, const float*, __nv_bfloat16*, int, RowCount, int, int, bool, cudaStream_t);
}
