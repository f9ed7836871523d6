// Shared device helpers for the b200rl kernels (sm_100a).
//
// Conventions used by every kernel in this directory
//   * trajectory tensors are row-major (T, C): time is the slow axis, the batch "column" c is contiguous, so a
//     warp reads 128 contiguous bytes of one time-step and a CTA owns a tile of columns for all T;
//   * loss heads reduce warp -> CTA -> grid inside the same kernel: every CTA stores its partial sums to the
//     caller-provided workspace and the last CTA to arrive (ticket from one atomic) adds them up in a fixed
//     order in fp64 -- deterministic, no second launch, no host sync;
//   * expressions whose rounding must match the reference's separate torch ops use __fmul_rn/__fadd_rn so ptxas
//     cannot contract them into FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define B200RL_OK 0
#ifndef B200RL_ERR_ARG
#define B200RL_ERR_ARG (-1)
#define B200RL_ERR_WORKSPACE (-2)
#endif

// workspace layout (floats): [0,16) control words, [16, ...) per-CTA partial sums
#define WS_CTRL_WORDS 16
#define WS_MIN_BYTES (1u << 20)

namespace b200rl {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// streaming loads / stores: every tensor on this path is touched once per launch
__device__ __forceinline__ float ldg_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ldg_stream4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) { __stcs(p, v); }

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

// Grid-wide sum of K per-thread partials.  After the call, in the LAST CTA to finish (returns true there, false
// elsewhere) thread 0 holds the grid totals in tot[0..K).  `ws` must hold WS_CTRL_WORDS + gridDim.x*K floats,
// control word `slot` must be zero on entry and is reset to zero on exit (stream-ordered reuse).
template <int K, int NT>
__device__ __forceinline__ bool grid_sum(float (&v)[K], double (&tot)[K], float* ws, int slot) {
    __shared__ float s_part[K][NT / 32];
    __shared__ double s_tot[K][NT / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float r = warp_sum(v[k]);
        if (lane == 0) s_part[k][wid] = r;
    }
    __syncthreads();
    float* part = ws + WS_CTRL_WORDS;
    unsigned int* ctrl = reinterpret_cast<unsigned int*>(ws);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float r = 0.f;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) r += s_part[k][w];
            part[(size_t)blockIdx.x * K + k] = r;
        }
        __threadfence();
        unsigned int ticket = atomicAdd(&ctrl[slot], 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += NT) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += (double)__ldcg(&part[(size_t)b * K + k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double r = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) s_tot[k][wid] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = 0.0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) r += s_tot[k][w];
            tot[k] = r;
        }
        ctrl[slot] = 0u;
    }
    return true;
}

// log-softmax statistics of one row of n logits read through `ld(j)`; L cooperating lanes (1 or 32) stride the row.
// Returns lse = logsumexp(z) and entropy H = -sum_j p_j log p_j with p_j = exp(z_j - lse)  (torch Categorical:
// logits - logsumexp, entropy = -(logits * softmax).sum(-1)).
template <int L, class Ld>
__device__ __forceinline__ void row_lse_entropy(Ld ld, int n, int lane, float& lse, float& ent) {
    float m = -INFINITY;
    for (int j = lane; j < n; j += L) m = fmaxf(m, ld(j));
    if (L == 32) m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < n; j += L) s += expf(ld(j) - m);
    if (L == 32) s = warp_sum(s);
    lse = m + logf(s);
    float h = 0.f;
    for (int j = lane; j < n; j += L) {
        float lp = ld(j) - lse;
        h += expf(lp) * fmaxf(lp, -3.402823466e38f);
    }
    if (L == 32) h = warp_sum(h);
    ent = -h;
}

template <int L, class Ld>
__device__ __forceinline__ float row_lse(Ld ld, int n, int lane) {
    float m = -INFINITY;
    for (int j = lane; j < n; j += L) m = fmaxf(m, ld(j));
    if (L == 32) m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < n; j += L) s += expf(ld(j) - m);
    if (L == 32) s = warp_sum(s);
    return m + logf(s);
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace b200rl
