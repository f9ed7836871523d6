// GAE reverse scan -- replaces the Python `for t in reversed(range(T))` of ding/rl_utils/gae.py:65-69.
//
// Layout: value/next_value/adv are (T, C) row-major with C = B (or B*A for the multi-agent case, gae.py:56-59);
// reward/done/traj_flag are (T, C/A) and broadcast over the trailing A.
//
// One CTA owns TC consecutive columns for all T.
//   phase 1 (all warps, fully parallel, float4 coalesced): nv' = nv*(1-done)   (written back only where it
//            changed -- the reference mutates the caller's tensor, gae.py:61), delta = (r + g*nv') - v,
//            f = gl*(1-traj)  -> shared memory [t][col]
//   phase 2 (TC lanes of warp 0): A_t = delta_t + f_t*A_{t+1} walking shared memory backwards; separate
//            round-to-nearest mul and add in the reference's order, so the result is bit-identical to the torch loop.
//   phase 3 (all warps): adv tile -> HBM, coalesced.
// T longer than CHUNK rows is processed in CHUNK-row slabs from the end of the trajectory, the carry stays in the
// scan lanes' registers.  Algorithmic traffic: 5 reads + 1 write = 24 B per transition.
#include "../../include/b200rl.h"
#include "common.cuh"

namespace b200rl {

template <int TC, int NT, int CHUNK, bool VEC>
__global__ void __launch_bounds__(NT) gae_tile_kernel(
    const float* __restrict__ value, float* __restrict__ next_value, const float* __restrict__ reward,
    const float* __restrict__ done, const float* __restrict__ traj, float* __restrict__ adv, long long T,
    long long C, long long A, float gamma, float gl, int mask_inplace) {
    __shared__ __align__(16) float s_d[CHUNK][TC];
    __shared__ __align__(16) float s_f[CHUNK][TC];
    const long long c0 = (long long)blockIdx.x * TC;
    const long long Caux = C / A;
    float carry = 0.f;  // live in lanes [0,TC) of warp 0
    for (long long hi = T; hi > 0; hi -= CHUNK) {
        const long long lo = hi > CHUNK ? hi - CHUNK : 0;
        const int rows = (int)(hi - lo);
        if (VEC) {
            // TC/4 threads per row, each owns 4 consecutive columns (A == 1, C % 4 == 0, 16B-aligned bases).
            // U rows are loaded back-to-back before any of them is consumed: 5*U 16-byte requests in flight/thread.
            constexpr int TPR = TC / 4;
            constexpr int RPP = NT / TPR;  // rows per pass
            constexpr int U = 4;
            const int cq = (threadIdx.x % TPR) * 4;
            const long long c = c0 + cq;
            if (c < C) {
                for (int rb = threadIdx.x / TPR; rb < rows; rb += RPP * U) {
                    float4 v[U], nv[U], rw[U], dn[U], tf[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = rb + u * RPP;
                        if (r < rows) {
                            const long long off = (lo + r) * C + c;
                            v[u] = ldg_stream4(reinterpret_cast<const float4*>(value + off));
                            nv[u] = ldg_stream4(reinterpret_cast<const float4*>(next_value + off));
                            rw[u] = ldg_stream4(reinterpret_cast<const float4*>(reward + off));
                            dn[u] = done ? ldg_stream4(reinterpret_cast<const float4*>(done + off))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                            tf[u] = traj ? ldg_stream4(reinterpret_cast<const float4*>(traj + off)) : dn[u];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = rb + u * RPP;
                        if (r < rows) {
                            float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, nn[4] = {nv[u].x, nv[u].y, nv[u].z, nv[u].w};
                            float rr[4] = {rw[u].x, rw[u].y, rw[u].z, rw[u].w};
                            float dd[4] = {dn[u].x, dn[u].y, dn[u].z, dn[u].w};
                            float tt[4] = {tf[u].x, tf[u].y, tf[u].z, tf[u].w};
                            float de[4], fa[4];
                            bool changed = false;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (done) {
                                    changed |= (dd[k] != 0.f);
                                    nn[k] = fmul(nn[k], fsub(1.f, dd[k]));
                                }
                                de[k] = fsub(fadd(rr[k], fmul(gamma, nn[k])), vv[k]);
                                fa[k] = fmul(gl, fsub(1.f, tt[k]));
                            }
                            *reinterpret_cast<float4*>(&s_d[r][cq]) = make_float4(de[0], de[1], de[2], de[3]);
                            *reinterpret_cast<float4*>(&s_f[r][cq]) = make_float4(fa[0], fa[1], fa[2], fa[3]);
                            if (changed && mask_inplace)
                                *reinterpret_cast<float4*>(next_value + (lo + r) * C + c) =
                                    make_float4(nn[0], nn[1], nn[2], nn[3]);
                        }
                    }
                }
            }
        } else {
            for (int i = threadIdx.x; i < rows * TC; i += NT) {
                const int r = i / TC, cc = i % TC;
                const long long c = c0 + cc;
                if (c < C) {
                    const long long off = (lo + r) * C + c;
                    const long long aoff = (lo + r) * Caux + c / A;
                    float v = value[off], nv = next_value[off], rw = reward[aoff];
                    float dn = done ? done[aoff] : 0.f;
                    float tf = traj ? traj[aoff] : dn;
                    if (done) {
                        float m = fmul(nv, fsub(1.f, dn));
                        if (mask_inplace && dn != 0.f) next_value[off] = m;
                        nv = m;
                    }
                    s_d[r][cc] = fsub(fadd(rw, fmul(gamma, nv)), v);
                    s_f[r][cc] = fmul(gl, fsub(1.f, tf));
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < TC && c0 + threadIdx.x < C) {
            const int cc = threadIdx.x;
            // software-pipelined walk: the next 4 (delta, f) pairs are in registers before the dependent chain needs them
            int r = rows - 1;
            for (; r >= 3; r -= 4) {
                float d0 = s_d[r][cc], f0 = s_f[r][cc], d1 = s_d[r - 1][cc], f1 = s_f[r - 1][cc];
                float d2 = s_d[r - 2][cc], f2 = s_f[r - 2][cc], d3 = s_d[r - 3][cc], f3 = s_f[r - 3][cc];
                carry = fadd(d0, fmul(f0, carry));
                s_d[r][cc] = carry;
                carry = fadd(d1, fmul(f1, carry));
                s_d[r - 1][cc] = carry;
                carry = fadd(d2, fmul(f2, carry));
                s_d[r - 2][cc] = carry;
                carry = fadd(d3, fmul(f3, carry));
                s_d[r - 3][cc] = carry;
            }
            for (; r >= 0; --r) {
                carry = fadd(s_d[r][cc], fmul(s_f[r][cc], carry));
                s_d[r][cc] = carry;
            }
        }
        __syncthreads();
        if (VEC) {
            constexpr int TPR = TC / 4;
            constexpr int RPP = NT / TPR;
            const int cq = (threadIdx.x % TPR) * 4;
            const long long c = c0 + cq;
            if (c < C) {
                for (int r = threadIdx.x / TPR; r < rows; r += RPP) {
                    float4 o = *reinterpret_cast<const float4*>(&s_d[r][cq]);
                    stg_stream4(reinterpret_cast<float4*>(adv + (lo + r) * C + c), o);
                }
            }
        } else {
            for (int i = threadIdx.x; i < rows * TC; i += NT) {
                const int r = i / TC, cc = i % TC;
                if (c0 + cc < C) adv[(lo + r) * C + c0 + cc] = s_d[r][cc];
            }
        }
        if (lo > 0) __syncthreads();
    }
}

template <int TC, int NT, int CHUNK>
static int launch_gae(const float* value, float* next_value, const float* reward, const float* done,
                      const float* traj, float* adv, long long T, long long C, long long A, float gamma, float gl,
                      int mask_inplace, bool vec, cudaStream_t st) {
    int grid = div_up(C, TC);
    if (vec)
        gae_tile_kernel<TC, NT, CHUNK, true>
            <<<grid, NT, 0, st>>>(value, next_value, reward, done, traj, adv, T, C, A, gamma, gl, mask_inplace);
    else
        gae_tile_kernel<TC, NT, CHUNK, false>
            <<<grid, NT, 0, st>>>(value, next_value, reward, done, traj, adv, T, C, A, gamma, gl, mask_inplace);
    return (int)cudaGetLastError();
}

}  // namespace b200rl

extern "C" int b200rl_gae(const float* value, float* next_value, const float* reward, const float* done,
                          const float* traj_flag, float* adv, long long T, long long C, long long A, double gamma_d,
                          double lambda_d, int mask_next_value_inplace, void* stream) {
    using namespace b200rl;
    // python scalars reach torch as fp32(gamma) and fp32(gamma*lambda_) (product taken in double), gae.py:62-63
    const float gamma = (float)gamma_d, gamma_lambda = (float)(gamma_d * lambda_d);
    if (T < 0 || C < 0 || A < 1 || (C % A) != 0) return B200RL_ERR_ARG;
    if (T == 0 || C == 0) return B200RL_OK;
    if (!value || !next_value || !reward || !adv) return B200RL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    bool vec = (A == 1) && (C % 4 == 0) && aligned16(value) && aligned16(next_value) && aligned16(reward) &&
               aligned16(adv) && (!done || aligned16(done)) && (!traj_flag || aligned16(traj_flag));
    // column-tile width: the widest tile that still gives every SM at least ~2 CTAs (148 SMs)
    if (C >= 32 * 296)
        return launch_gae<32, 256, 128>(value, next_value, reward, done, traj_flag, adv, T, C, A, gamma,
                                        gamma_lambda, mask_next_value_inplace, vec, st);
    if (C >= 16 * 296)
        return launch_gae<16, 128, 128>(value, next_value, reward, done, traj_flag, adv, T, C, A, gamma,
                                        gamma_lambda, mask_next_value_inplace, vec, st);
    return launch_gae<8, 64, 128>(value, next_value, reward, done, traj_flag, adv, T, C, A, gamma, gamma_lambda,
                                  mask_next_value_inplace, vec, st);
}
