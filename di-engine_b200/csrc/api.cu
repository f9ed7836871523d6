// Library-level entry points of the C ABI declared in include/b200rl.h.
#include "../../include/b200rl.h"
#include "common.cuh"

extern "C" int b200rl_version(void) { return 100; }
extern "C" int b200rl_built_for_sm(void) { return 100; }
extern "C" size_t b200rl_workspace_bytes(void) { return (size_t)WS_MIN_BYTES; }

// ---------------------------------------------------------------------------------------------------------------
// Bandwidth probe (tools/calib_copy.py): a plain persistent float4 copy with 8 independent 16-byte loads in flight per
// thread.  It calibrates what a well-formed streaming kernel reaches at the (small) byte counts of this path, where
// launch ramp-up and DRAM latency are a large share of the run time and the 2 GiB-copy peak is out of reach.
// ---------------------------------------------------------------------------------------------------------------
namespace b200rl {
__global__ void __launch_bounds__(256) probe_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                         long long n) {
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(src + i + k * stride);
#pragma unroll
        for (int k = 0; k < 8; ++k) __stcs(dst + i + k * stride, v[k]);
    }
    for (; i < n; i += stride) __stcs(dst + i, __ldcs(src + i));
}
}  // namespace b200rl

extern "C" int b200rl_probe_copy(const float* src, float* dst, long long n_floats, int ctas_per_sm, void* stream) {
    if (!src || !dst || n_floats < 0 || (n_floats & 3) || ctas_per_sm < 1) return B200RL_ERR_ARG;
    const long long n4 = n_floats / 4;
    long long grid = 148LL * ctas_per_sm;
    const long long need = (n4 + 255) / 256;
    if (grid > need) grid = need > 0 ? need : 1;
    (void)b200rl::launch_k(b200rl::probe_copy_kernel, (int)grid, 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(src),
                                                                          reinterpret_cast<float4*>(dst), n4);
    return (int)cudaGetLastError();
}
