// Common device/host helpers for libttb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace ttb {

// ---- error plumbing (C ABI returns int, message kept per thread) -------------------------
void set_last_error(const std::string& msg);

#define TTB_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ttb::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + \
                                " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");    \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

#define TTB_CHECK(cond, msg)                                                          \
    do {                                                                              \
        if (!(cond)) {                                                                \
            ttb::set_last_error(std::string(msg) + " (" + __FILE__ + ":" +            \
                                std::to_string(__LINE__) + ")");                      \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

// ---- storage-type conversion -----------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int kNumSMs = 148;  // B200

// Control words of the device-resident decoding loops (int32 each).
enum Ctrl : int {
    CTRL_N_ACTIVE = 0,   // live queries (rows of the compact token matrix = n_active * rows_per_query)
    CTRL_WIDTH = 1,      // width of the reference's token matrix for the coming iteration (Wn)
    CTRL_DONE = 2,       // 1 -> loop finished (all queries done, width limit, or error)
    CTRL_ERROR = 3,      // 0 ok, 1 draft splice out of bounds, 2 finished-row shape mismatch
    CTRL_ITERS = 4,      // decoder calls made
    CTRL_ACCEPTED = 5,   // accepted draft tokens (sum over queries and iterations)
    CTRL_TOKENS = 6,     // tokens produced (accepted + bonus)
    CTRL_PREV_WIDTH = 7, // width before growing (W)
    CTRL_N_SEL = 8,      // rows recorded in GreedyState::sel by the last accept (live queries before retirement)
    CTRL_N_LEFT = 9,     // live queries left when DONE was raised (N_ACTIVE is zeroed then so that every
                         // row-counted kernel of an already-enqueued iteration exits immediately)
    CTRL_LS = 10,        // source length of the batch (read by the cross-attention kernels so that the
                         // captured decoding graph does not depend on it)
    CTRL_COUNT = 16
};

// Control words of the speculative beam search (read back by the host once per iteration).
enum BeamCtrl : int {
    BC_NLIVE_ROWS = 0,    // (candidate, draft) rows of unfinished candidates = decoder batch of the iteration
    BC_ERROR = 1,         // 0 ok, 3 draft slots not contiguous (PAD inside a hypothesis), 4 fewer leaves than n_best
    BC_ALL_FINISHED = 2,  // every new candidate contains EOS
    BC_EMPTY_COLS = 3,    // min over new candidates of the number of PAD columns
    BC_ACCEPTED = 4,      // accepted draft tokens of the surviving hypotheses (reference accepted_tokens_num)
    BC_PRODUCED = 5,      // reference produced_non_pad_tokens
    BC_COUNT = 8
};

}  // namespace ttb
