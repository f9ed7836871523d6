// compute_q_retraces (ding/rl_utils/retrace.py:7-56): ACER's Retrace(lambda) targets, a reverse scan along T with two gathers
// per step.  No gradient (the reference computes it under torch.no_grad(), policy/acer.py:231-232).
//
//   Qret[T] = V[T];  tmp = V[T]
//   for t = T-1 .. 0:   Qret[t] = r[t] + (gamma * w[t]) * tmp
//                       tmp     = min(ratio[t, a_t], 1) * (Qret[t] - Q[t, a_t]) + V[t]
//
// thread = batch column; the operands of 8 time steps are requested together (the loads do not depend on the carry), then the
// dependent chain runs on registers.  Separate round-to-nearest multiplies and adds in the reference's order: bit-identical.
// Algorithmic traffic: 28 B + one sector each of the two gathered rows per transition.
#include "../../include/b200rl.h"
#include "common.cuh"

namespace b200rl {

constexpr int RT_U = 8;

__global__ void __launch_bounds__(128) retrace_kernel(const float* __restrict__ q, const float* __restrict__ v,
                                                      const float* __restrict__ reward, const long long* __restrict__ action,
                                                      const float* __restrict__ weight, const float* __restrict__ ratio,
                                                      long long T, long long B, long long N, float gamma,
                                                      float* __restrict__ out) {
    pdl_prologue();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float tmp = v[T * B + b];
    out[T * B + b] = tmp;
    long long t1 = T;
    while (t1 > 0) {
        const int n = t1 >= RT_U ? RT_U : (int)t1;
        float r[RT_U], gw[RT_U], c[RT_U], qa[RT_U], vv[RT_U];
#pragma unroll
        for (int k = 0; k < RT_U; ++k) {
            if (k < n) {
                const long long e = (t1 - 1 - k) * B + b;
                const long long a = action[e];
                r[k] = reward[e];
                gw[k] = fmul(gamma, weight[e]);
                c[k] = fminf(ratio[e * N + a], 1.0f);
                qa[k] = q[e * N + a];
                vv[k] = v[e];
            }
        }
#pragma unroll
        for (int k = 0; k < RT_U; ++k) {
            if (k < n) {
                const float qr = fadd(r[k], fmul(gw[k], tmp));
                out[(t1 - 1 - k) * B + b] = qr;
                tmp = fadd(fmul(c[k], fsub(qr, qa[k])), vv[k]);
            }
        }
        t1 -= n;
    }
}

}  // namespace b200rl

extern "C" int b200rl_q_retraces(const float* q_values, const float* v_pred, const float* rewards, const long long* actions,
                                 const float* weights, const float* ratio, long long T, long long B, long long N,
                                 double gamma, float* q_retraces, void* stream) {
    using namespace b200rl;
    if (T < 0 || B < 1 || N < 1 || !v_pred || !q_retraces) return B200RL_ERR_ARG;
    if (T > 0 && (!q_values || !rewards || !actions || !weights || !ratio)) return B200RL_ERR_ARG;
    (void)launch_k(retrace_kernel, div_up(B, 128), 128, 0, (cudaStream_t)stream, q_values, v_pred, rewards, actions, weights,
                   ratio, T, B, N, (float)gamma, q_retraces);
    return (int)cudaGetLastError();
}
