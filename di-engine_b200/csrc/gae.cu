// GAE reverse scan -- replaces the Python `for t in reversed(range(T))` of ding/rl_utils/gae.py:65-69.
//
// Layout: value/next_value/adv are (T, C) row-major with C = B (or B*A for the multi-agent case, gae.py:56-59);
// reward/done/traj_flag are (T, C/A) and broadcast over the trailing A.
//
// One CTA owns TC consecutive columns for all T (see gae_ws_kernel):
//   loader warps (fully parallel, float4 coalesced): nv' = nv*(1-done) (written back only where it changed -- the
//            reference mutates the caller's tensor, gae.py:61), delta = (r + g*nv') - v, f = gl*(1-traj) -> shared memory
//   scan warp (lane = column): A_t = delta_t + f_t*A_{t+1}; separate round-to-nearest mul and add in the reference's
//            order, so the result is bit-identical to the torch loop; adv is stored straight from the scan registers.
// T is processed in 128-row slabs from the end of the trajectory, each slab in four 32-row chunks handed from the
// loaders to the scan warp through named barriers; the carry stays in the scan lanes' registers.
// Algorithmic traffic: 5 reads + 1 write = 24 B per transition.
#include <stdlib.h>

#include "../../include/b200rl.h"
#include "common.cuh"

namespace b200rl {

// named barriers with immediate ids (a register id would make ptxas reserve all 16 hardware barriers per CTA)
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_arrive() {
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int COUNT>
__device__ __forceinline__ void chunk_arrive(int k) {
    if (k == 0) named_bar_arrive<1, COUNT>();
    else if (k == 1) named_bar_arrive<2, COUNT>();
    else if (k == 2) named_bar_arrive<3, COUNT>();
    else named_bar_arrive<4, COUNT>();
}
template <int COUNT>
__device__ __forceinline__ void chunk_wait(int k) {
    if (k == 0) named_bar_sync<1, COUNT>();
    else if (k == 1) named_bar_sync<2, COUNT>();
    else if (k == 2) named_bar_sync<3, COUNT>();
    else named_bar_sync<4, COUNT>();
}

constexpr int GAE_CH = 32;      // rows per chunk (one named barrier per chunk)
constexpr int GAE_NCHUNK = 4;   // chunks per slab of T
constexpr int GAE_SLAB = GAE_CH * GAE_NCHUNK;

// Warp-specialised tile kernel.  A CTA owns TC columns; warp 0 is the scan warp, the other TC/4 warps are loaders.
// Loaders issue every 16-byte load of a 128-row slab up front (20 in flight per thread), newest rows first, then turn
// each 32-row chunk into (delta, f) pairs in shared memory and signal the chunk's named barrier.  The scan warp walks
// the chunks as they arrive -- the sequential part of chunk k overlaps the HBM latency of chunks k+1.. -- and stores
// adv straight from registers.
template <int TC, bool VEC>
__global__ void __launch_bounds__((TC / 4 + 1) * 32) gae_ws_kernel(
    const float* __restrict__ value, float* __restrict__ next_value, const float* __restrict__ reward,
    const float* __restrict__ done, const float* __restrict__ traj, float* __restrict__ adv, long long T,
    long long C, long long A, float gamma, float gl, int mask_inplace) {
    pdl_prologue();
    constexpr int NL = (TC / 4) * 32;  // loader threads
    constexpr int NTHREADS = NL + 32;
    __shared__ __align__(16) float s_d[GAE_NCHUNK][GAE_CH][TC];
    __shared__ __align__(16) float s_f[GAE_NCHUNK][GAE_CH][TC];
    const long long c0 = (long long)blockIdx.x * TC;
    const long long Caux = C / A;
    const bool is_scan = threadIdx.x < 32;
    const int ltid = threadIdx.x - 32;  // loader thread index
    float carry = 0.f;                  // scan lanes
    for (long long hi = T; hi > 0; hi -= GAE_SLAB) {
        const long long lo = hi > GAE_SLAB ? hi - GAE_SLAB : 0;
        const int rows = (int)(hi - lo);
        // chunk k covers slab rows [rlo_k, rhi_k), k = 0 is the newest (processed first)
        if (!is_scan) {
            if (VEC) {
                constexpr int TPR = TC / 4;  // NL / TPR == 32 rows per pass == one chunk
                const int cq = (ltid % TPR) * 4;
                const int rr = ltid / TPR;   // row inside the chunk, counted from the chunk's top (newest) row
                const long long c = c0 + cq;
                const bool col_ok = c < C;
                float4 v[GAE_NCHUNK], nv[GAE_NCHUNK], rw[GAE_NCHUNK], dn[GAE_NCHUNK], tf[GAE_NCHUNK];
#pragma unroll
                for (int k = 0; k < GAE_NCHUNK; ++k) {
                    const int r = rows - 1 - k * GAE_CH - rr;  // slab row of this thread in chunk k
                    if (r >= 0 && col_ok) {
                        const long long off = (lo + r) * C + c;
                        v[k] = ldg_stream4(reinterpret_cast<const float4*>(value + off));
                        nv[k] = ldg_stream4(reinterpret_cast<const float4*>(next_value + off));
                        rw[k] = ldg_stream4(reinterpret_cast<const float4*>(reward + off));
                        dn[k] = done ? ldg_stream4(reinterpret_cast<const float4*>(done + off))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                        tf[k] = traj ? ldg_stream4(reinterpret_cast<const float4*>(traj + off)) : dn[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < GAE_NCHUNK; ++k) {
                    const int r = rows - 1 - k * GAE_CH - rr;
                    if (r >= 0 && col_ok) {
                        float vv[4] = {v[k].x, v[k].y, v[k].z, v[k].w}, nn[4] = {nv[k].x, nv[k].y, nv[k].z, nv[k].w};
                        float rw4[4] = {rw[k].x, rw[k].y, rw[k].z, rw[k].w};
                        float dd[4] = {dn[k].x, dn[k].y, dn[k].z, dn[k].w};
                        float tt[4] = {tf[k].x, tf[k].y, tf[k].z, tf[k].w};
                        float de[4], fa[4];
                        bool changed = false;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (done) {
                                changed |= (dd[q] != 0.f);
                                nn[q] = fmul(nn[q], fsub(1.f, dd[q]));
                            }
                            de[q] = fsub(fadd(rw4[q], fmul(gamma, nn[q])), vv[q]);
                            fa[q] = fmul(gl, fsub(1.f, tt[q]));
                        }
                        *reinterpret_cast<float4*>(&s_d[k][rr][cq]) = make_float4(de[0], de[1], de[2], de[3]);
                        *reinterpret_cast<float4*>(&s_f[k][rr][cq]) = make_float4(fa[0], fa[1], fa[2], fa[3]);
                        if (changed && mask_inplace)
                            *reinterpret_cast<float4*>(next_value + (lo + r) * C + c) =
                                make_float4(nn[0], nn[1], nn[2], nn[3]);
                    }
                    chunk_arrive<NTHREADS>(k);
                }
            } else {
#pragma unroll
                for (int k = 0; k < GAE_NCHUNK; ++k) {
                    const int rtop = rows - 1 - k * GAE_CH;  // newest slab row of chunk k
                    for (int i = ltid; i < GAE_CH * TC; i += NL) {
                        const int rr = i / TC, cc = i % TC;
                        const int r = rtop - rr;
                        const long long c = c0 + cc;
                        if (r >= 0 && c < C) {
                            const long long off = (lo + r) * C + c;
                            const long long aoff = (lo + r) * Caux + c / A;
                            float nvv = next_value[off];
                            const float dnn = done ? done[aoff] : 0.f;
                            const float tff = traj ? traj[aoff] : dnn;
                            if (done) {
                                const float mm = fmul(nvv, fsub(1.f, dnn));
                                if (mask_inplace && dnn != 0.f) next_value[off] = mm;
                                nvv = mm;
                            }
                            s_d[k][rr][cc] = fsub(fadd(reward[aoff], fmul(gamma, nvv)), value[off]);
                            s_f[k][rr][cc] = fmul(gl, fsub(1.f, tff));
                        }
                    }
                    chunk_arrive<NTHREADS>(k);
                }
            }
        } else {
            const int cc = threadIdx.x;
            const bool lane_ok = cc < TC && c0 + cc < C;
            float* out = adv + lo * C + c0 + cc;
#pragma unroll 1
            for (int k = 0; k < GAE_NCHUNK; ++k) {
                chunk_wait<NTHREADS>(k);
                const int rtop = rows - 1 - k * GAE_CH;
                if (lane_ok && rtop >= 0) {
                    if (rtop >= GAE_CH - 1) {  // full chunk: registers first, then the dependent chain
                        float d[GAE_CH], f[GAE_CH];
#pragma unroll
                        for (int j = 0; j < GAE_CH; ++j) {
                            d[j] = s_d[k][j][cc];
                            f[j] = s_f[k][j][cc];
                        }
#pragma unroll
                        for (int j = 0; j < GAE_CH; ++j) {
                            carry = fadd(d[j], fmul(f[j], carry));
                            out[(long long)(rtop - j) * C] = carry;
                        }
                    } else {
                        for (int j = 0; j <= rtop; ++j) {
                            carry = fadd(s_d[k][j][cc], fmul(s_f[k][j][cc], carry));
                            out[(long long)(rtop - j) * C] = carry;
                        }
                    }
                }
            }
        }
        if (lo > 0) __syncthreads();  // the next slab reuses the chunk buffers
    }
}

template <int TC>
static int launch_gae(const float* value, float* next_value, const float* reward, const float* done,
                      const float* traj, float* adv, long long T, long long C, long long A, float gamma, float gl,
                      int mask_inplace, bool vec, cudaStream_t st) {
    const int grid = div_up(C, TC);
    constexpr int NT = (TC / 4 + 1) * 32;
    if (vec)
        (void)launch_k(gae_ws_kernel<TC, true>, grid, NT, 0, st, value, next_value, reward, done, traj, adv, T, C, A, gamma, gl,
                                                      mask_inplace);
    else
        (void)launch_k(gae_ws_kernel<TC, false>, grid, NT, 0, st, value, next_value, reward, done, traj, adv, T, C, A, gamma, gl,
                                                       mask_inplace);
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
    // column-tile width: the widest tile that still gives every SM at least ~2 CTAs (148 SMs);
    // B200RL_GAE_TC=8|16|32 overrides the choice (tuning experiments)
    static int forced_tc = -1;
    if (forced_tc < 0) {
        const char* e = getenv("B200RL_GAE_TC");
        forced_tc = e ? atoi(e) : 0;
    }
    if (forced_tc == 32 || (forced_tc == 0 && C >= 32 * 296))
        return launch_gae<32>(value, next_value, reward, done, traj_flag, adv, T, C, A, gamma, gamma_lambda,
                              mask_next_value_inplace, vec, st);
    if (forced_tc == 16 || (forced_tc == 0 && C >= 16 * 296))
        return launch_gae<16>(value, next_value, reward, done, traj_flag, adv, T, C, A, gamma, gamma_lambda,
                              mask_next_value_inplace, vec, st);
    return launch_gae<8>(value, next_value, reward, done, traj_flag, adv, T, C, A, gamma, gamma_lambda,
                         mask_next_value_inplace, vec, st);
}
