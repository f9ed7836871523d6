// n-step TD heads of ding/rl_utils/td.py:
//   q_nstep_td_error (:649-719) / q_nstep_td_error_with_rescale (:810-867)  -> qntd_fwd / qntd_bwd
//   dist_nstep_td_error (C51 projection, :413-523)                           -> dntd_fwd / dntd_bwd
//   generalized_lambda_returns / multistep_forward_view (:1574-1651), td_lambda_error (:1539-1571) and the
//   UPGO return (upgo.py:46-68)                                              -> lambda_returns (+ fused TD(lambda) head)
// Each reference call is ~20-35 tiny torch launches plus an autograd backward; here it is one forward and one
// backward launch.  These batches are small (B ~ 32..512): latency, not bandwidth, is what the kernels minimise.
#include <math.h>

#include "../../include/b200rl.h"
#include "common.cuh"
#include "ppo_math.cuh"  // lg2f_ / rcpf_ / kLn2

namespace b200rl {

// ---------------------------------------------------------------------------------------------------------------
// value rescaling h / h^-1 (value_rescale.py:4-34), evaluated in the reference's operation order
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
__device__ __forceinline__ float value_h(float x, float eps) {
    return fadd(fmul(sgn(x), fsub(__fsqrt_rn(fadd(fabsf(x), 1.f)), 1.f)), fmul(eps, x));
}
__device__ __forceinline__ float value_h_inv(float x, float eps, float four_eps, float two_eps) {
    float t = fadd(fadd(fabsf(x), 1.f), eps);
    float u = __fdiv_rn(fsub(__fsqrt_rn(fadd(1.f, fmul(four_eps, t))), 1.f), two_eps);
    return fmul(sgn(x), fsub(fmul(u, u), 1.f));
}

// elementwise criteria (reduction='none'): value and d/d(input)
__device__ __forceinline__ float criterion_eval(int kind, float param, float x, float y, float& dx) {
    const float d = x - y, ad = fabsf(d);
    switch (kind) {
        case 0: dx = 2.f * d; return d * d;                                        // nn.MSELoss
        case 1: dx = sgn(d); return ad;                                            // nn.L1Loss
        case 2:                                                                    // nn.SmoothL1Loss(beta)
            if (ad < param) { dx = d / param; return 0.5f * d * d / param; }
            dx = sgn(d); return ad - 0.5f * param;
        default:                                                                   // nn.HuberLoss(delta)
            if (ad <= param) { dx = d; return 0.5f * d * d; }
            dx = param * sgn(d); return param * (ad - 0.5f * param);
    }
}

struct QntdArgs {
    const float* q;            // (S, G, N)
    const float* next_q;       // (S, G, N)
    const long long* action;   // (S, G)
    const long long* next_action;
    const float* reward;       // (nstep, S), (S) when cum_reward; sequence form (S = Tseq*Bcol): (Tseq, nstep, Bcol)
    const float* done;         // (S)
    const float* weight;       // (S) or null
    const float* value_gamma;  // null, or pointer with stride 0 (0-dim) / 1 (S)
    long long value_gamma_stride;
    const float* gamma_ps;     // NGU per-sample gamma (Bcol) or null
    long long S;               // samples
    long long Bcol;            // == S, or the batch width of the sequence form (sample s = t*Bcol + b)
    int G;                     // rows per sample: 1, the agent dim of the multi-agent form or the BDQ branches
    int N;
    int nstep;
    float gamma;
    float gamma_pow_n;
    int cum_reward;
    int rescale;
    float eps, four_eps, two_eps;
    int criterion;
    float crit_param;
    int group_mean;    // td_error_per_sample (S) = mean over the G rows (bdq_nstep_td_error, td.py:788) instead of (S, G)
    double loss_div;   // loss = sum(w*td) / loss_div
    float prio_max_w, prio_mean_w, prio_div;  // sequence form: priority[b] = max_w*max_t|td| + mean_w*sum_t|td|/prio_div
    float* loss;
    float* td_err;
    float* dcrit;      // (S, G) d criterion / d q_sa (unweighted): what the backward launch needs
    float* target;     // nullable: the (detached) n-step target (S, G), for callers that apply their own criterion
    float* grad_unit;  // nullable (S, G, N): d loss / d q for a unit upstream gradient, written by the forward launch
    float* priority;   // nullable (Bcol): sequence form only
};

// One launch: n-step target, criterion, deterministic loss reduction, per-sample errors AND (grad_unit) the gradient of the
// loss for a unit upstream gradient -- the backward launch only has to verify that the upstream gradient was 1.
template <int NT>
__global__ void __launch_bounds__(NT) qntd_fwd_kernel(QntdArgs a, float* ws) {
    pdl_prologue();
    __shared__ float s_coef[NT];
    __shared__ int s_act[NT];
    const long long s0 = (long long)blockIdx.x * NT;
    const long long s = s0 + threadIdx.x;
    const float inv_div = (float)(1.0 / a.loss_div);
    float acc[1] = {0.f};
    s_coef[threadIdx.x] = 0.f;
    s_act[threadIdx.x] = -1;
    if (s < a.S) {
        const long long tq_ = s / a.Bcol, b = s - tq_ * a.Bcol;  // sequence step / batch column (tq_ = 0 when Bcol == S)
        // every per-sample operand is requested before any is used: a load inside the n-step loop is one L2 round trip per step
        const float* __restrict__ rw = a.cum_reward ? a.reward + s : a.reward + tq_ * a.nstep * a.Bcol + b;
        const size_t Bi = (size_t)a.Bcol;
        const int nrw = a.cum_reward ? 1 : a.nstep;
        float rwv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rwv[i] = i < nrw ? rw[i * Bi] : 0.f;
        const float dn = a.done[s];
        const float w = a.weight ? a.weight[s] : 1.f;
        const float vgl = a.value_gamma ? a.value_gamma[s * a.value_gamma_stride] : a.gamma_pow_n;
        const float g = a.gamma_ps ? a.gamma_ps[b] : a.gamma;
        const int act0 = (int)a.action[s * a.G], nact0 = (int)a.next_action[s * a.G];
        const float q0 = a.q[s * a.G * a.N + act0], nq0 = a.next_q[s * a.G * a.N + nact0];
        const float nd = fsub(1.f, dn);
        float ret = 0.f, vg;
        if (a.cum_reward) {
            ret = rwv[0];
            vg = vgl;
        } else {
            float rf = 1.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // td.py:261-264 / :277-281
                if (i < a.nstep) {
                    ret = fadd(ret, fmul(rwv[i], rf));
                    rf = fmul(g, rf);
                }
            }
            for (int i = 8; i < a.nstep; ++i) {
                ret = fadd(ret, fmul(rw[i * Bi], rf));
                rf = fmul(g, rf);
            }
            vg = a.gamma_ps ? rf : vgl;  // reward_factor[nstep] for the NGU list-gamma form
        }
        float td_sum = 0.f;
        for (int gi = 0; gi < a.G; ++gi) {
            const long long r = s * a.G + gi;
            const int act = gi == 0 ? act0 : (int)a.action[r];
            const float q_sa = gi == 0 ? q0 : a.q[r * a.N + act];
            float tq = gi == 0 ? nq0 : a.next_q[r * a.N + a.next_action[r]];
            if (a.rescale) tq = value_h_inv(tq, a.eps, a.four_eps, a.two_eps);
            float target = fadd(ret, fmul(fmul(vg, tq), nd));  // td.py:266 / :273 / :282 / :712-715
            if (a.rescale) target = value_h(target, a.eps);
            float dx;
            const float td = criterion_eval(a.criterion, a.crit_param, q_sa, target, dx);
            if (a.group_mean) td_sum += td;
            else a.td_err[r] = td;
            if (a.target) a.target[r] = target;
            a.dcrit[r] = dx;
            acc[0] += td * w;
            if (a.grad_unit) {
                const float coef = w * dx * inv_div;
                if (a.G == 1) {
                    s_coef[threadIdx.x] = coef;
                    s_act[threadIdx.x] = act;
                } else {
                    float* gq = a.grad_unit + r * a.N;
                    for (int j = 0; j < a.N; ++j) gq[j] = (j == act) ? coef : 0.f;
                }
            }
        }
        if (a.group_mean) a.td_err[s] = td_sum / (float)a.G;
    }
    if (a.grad_unit && a.G == 1) {  // the block's NT gradient rows are contiguous: coalesced stores
        __syncthreads();
        const long long rows = (a.S - s0) < NT ? (a.S - s0) : NT;
        float* gq = a.grad_unit + s0 * a.N;
        for (long long i = threadIdx.x; i < rows * a.N; i += NT) {
            const int rr = (int)(i / a.N), j = (int)(i - (long long)rr * a.N);
            gq[i] = (j == s_act[rr]) ? s_coef[rr] : 0.f;
        }
    }
    if (!a.priority && gridDim.x <= FX_MAX_GRID) {  // one atomic round trip (none for a one-CTA grid)
        grid_sum_fx<1, NT>(acc, ws, [&](int, double t) { a.loss[0] = (float)(t / a.loss_div); });
        return;
    }
    double tot[1];
    const bool last = grid_sum<1, NT>(acc, tot, ws, 0);
    if (last && threadIdx.x == 0) a.loss[0] = (float)(tot[0] / a.loss_div);
    if (last && a.priority) {
        // sequence form (ding/policy/r2d2.py:367-369): the last CTA sees every per-step error (published before the tickets)
        const long long Tq = a.S / a.Bcol;
        for (long long b = threadIdx.x; b < a.Bcol; b += NT) {
            float mx = 0.f, sm = 0.f;
            for (long long t = 0; t < Tq; ++t) {
                const float e = fabsf(__ldcg(a.td_err + t * a.Bcol + b));
                mx = (t == 0) ? e : fmaxf(mx, e);
                sm += e;
            }
            a.priority[b] = a.prio_max_w * mx + a.prio_mean_w * (sm / a.prio_div);
        }
    }
}

// grad_q[r, j] = 1[j == a_r] * (g_loss * w_s / loss_div + g_td) * dcrit[r].  skip_if_unit: the forward launch already wrote
// grad_q for g_loss == 1 and no per-sample upstream gradient -- verify on the device and leave at once if that held.
__global__ void qntd_bwd_kernel(const float* __restrict__ dcrit, const float* __restrict__ weight,
                                const long long* __restrict__ action, const float* __restrict__ g_loss,
                                const float* __restrict__ g_td, long long S, int G, int N, int group_mean, float inv_div,
                                int skip_if_unit, float* __restrict__ grad_q) {
    pdl_prologue();
    const float g = g_loss ? *g_loss : 0.f;
    if (skip_if_unit && !g_td && g == 1.f) return;
    const long long n = S * G * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / N;
        const int j = (int)(i - r * N);
        float out = 0.f;
        if (j == (int)action[r]) {
            const long long s = r / G;
            float c = g * (weight ? weight[s] : 1.f) * inv_div;
            if (g_td) c += group_mean ? g_td[s] / (float)G : g_td[r];
            out = c * dcrit[r];
        }
        grad_q[i] = out;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C51 categorical projection: one warp per (sample, agent) row, lanes stride the atoms, the projected distribution is
// accumulated with shared-memory atomics (the index_add_ of td.py:510-511) and kept for the backward pass.
// ---------------------------------------------------------------------------------------------------------------
struct DntdArgs {
    const float* dist;       // (R, N, n_atom)
    const float* next_dist;  // (R, N, n_atom)
    const long long* act;    // (R)
    const long long* next_act;
    const float* reward;  // (nstep, B)
    const float* done;    // (B)
    const float* weight;  // null, or stride 0 / 1 over R
    long long weight_stride;
    const float* value_gamma;  // null or stride 0 / 1 over B
    long long value_gamma_stride;
    const float* support;  // (n_atom), torch.linspace(v_min, v_max, n_atom) computed by the caller
    long long R;
    long long A;  // rows per batch entry (1 single agent)
    long long B;
    int N;
    int n_atom;
    int nstep;
    float gamma;
    float gamma_pow_n;
    float v_min, v_max, delta_z;
    float* loss;
    float* td_err;  // (R)
    float* proj;    // (R, n_atom)
    int* bad_flag;  // set to 1 if any selected dist entry is <= 0 (td.py:513)
    float* grad_unit;  // nullable (R, N, n_atom): d loss / d dist for a unit upstream gradient
};

// NJ = atoms per lane (ceil(n_atom / 32), 2 for C51): every global operand of a row -- the chosen rows of dist / next_n_dist,
// the support, the n-step rewards -- is requested in ONE batch right after the two action indices have arrived, and lives in
// registers from then on (the first build reloaded dist[j] in every iteration of the gradient loop: 12 dependent L2 round
// trips per row, 10.8 us at config C; profiles/r02_ncu_small_kernels.md).
template <int NT, int NJ>
__global__ void __launch_bounds__(NT) dntd_fwd_kernel(DntdArgs a, float* ws) {
    pdl_prologue();
    extern __shared__ float s_proj[];  // [NT/32][n_atom]
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (NT / 32) + wid;
    const int na = a.n_atom;
    float* pj = s_proj + wid * na;
    float acc[1] = {0.f};
    // a row's work is one serial chain in one warp (~700 instructions): 32-bit index arithmetic, MUFU log / reciprocal where
    // the reference's bits do not depend on them (the bin positions keep the IEEE division of td.py:500)
    if (r < a.R) {
        const long long b = r / a.A;
        const int sel = (int)a.act[r], nsel = (int)a.next_act[r];
        const float* __restrict__ nd = a.next_dist + (r * a.N + nsel) * na;
        const float* __restrict__ dd = a.dist + (r * a.N + sel) * na;
        float p_[NJ], d_[NJ], z_[NJ];
#pragma unroll
        for (int u = 0; u < NJ; ++u) {
            const int j = lane + 32 * u;
            const bool ok = j < na;
            p_[u] = ok ? nd[j] : 0.f;
            d_[u] = ok ? dd[j] : 1.f;
            z_[u] = ok ? a.support[j] : 0.f;
            if (ok) pj[j] = 0.f;
        }
        // every per-sample scalar is requested before any is used (a load inside the n-step loop costs one L2 round trip per
        // step: the first build spent 6 serialised round trips here, ~4 of its 8 us)
        const float* __restrict__ rw = a.reward + b;
        const size_t Bi = (size_t)a.B;
        float rwv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rwv[i] = i < a.nstep ? rw[i * Bi] : 0.f;
        const float dn = a.done[b];
        const float vg = a.value_gamma ? a.value_gamma[b * a.value_gamma_stride] : a.gamma_pow_n;
        const float w = a.weight ? a.weight[r * a.weight_stride] : 1.f;
        float rf = 1.f, ret = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // matmul(reward_factor, reward), td.py:453-456
            if (i < a.nstep) {
                ret = fadd(ret, fmul(rf, rwv[i]));
                rf = fmul(a.gamma, rf);
            }
        }
        for (int i = 8; i < a.nstep; ++i) {
            ret = fadd(ret, fmul(rf, rw[i * Bi]));
            rf = fmul(a.gamma, rf);
        }
        const float scale = fmul(fsub(1.f, dn), vg);  // (1-done) * gamma**n, td.py:492-498
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NJ; ++u) {
            const int j = lane + 32 * u;
            int key = -1 - lane;  // lanes past n_atom: unique keys, zero contributions
            float clo = 0.f, chi = 0.f;
            if (j < na) {
                float tz = fadd(ret, fmul(scale, z_[u]));
                tz = fminf(fmaxf(tz, a.v_min), a.v_max);
                const float pos = __fdiv_rn(fsub(tz, a.v_min), a.delta_z);  // td.py:500
                float lo = floorf(pos), hi = ceilf(pos);
                if (hi > 0.f && lo == hi) lo -= 1.f;                          // td.py:504
                if (lo < (float)(na - 1) && lo == hi) hi += 1.f;              // td.py:505 (now hi == lo + 1 always)
                key = (int)lo;
                clo = fmul(p_[u], fsub(hi, pos));
                chi = fmul(p_[u], fsub(pos, lo));
            }
            // Tz is monotone in the atom index, so lanes that hit the same bin are CONTIGUOUS (a terminal sample, done = 1, or a
            // clamped tail sends all of them to one bin): a segmented warp scan adds them up and only the last lane of a
            // segment touches shared memory -- the 32-way CAS loop of the first build cost up to ~3 us per such row
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int k2 = __shfl_up_sync(0xffffffffu, key, d);
                const float l2 = __shfl_up_sync(0xffffffffu, clo, d), h2 = __shfl_up_sync(0xffffffffu, chi, d);
                if (lane >= d && k2 == key) {
                    clo += l2;
                    chi += h2;
                }
            }
            const int knext = __shfl_down_sync(0xffffffffu, key, 1);
            if (key >= 0 && (lane == 31 || knext != key)) {
                atomicAdd(&pj[key], clo);
                atomicAdd(&pj[key + 1], chi);
            }
        }
        __syncwarp();
        float td = 0.f;
        bool bad = false;
        float m_[NJ];
        float* __restrict__ prow = a.proj + r * na;
#pragma unroll
        for (int u = 0; u < NJ; ++u) {
            const int j = lane + 32 * u;
            m_[u] = 0.f;
            if (j < na) {
                m_[u] = pj[j];
                bad |= !(d_[u] > 0.f);
                td = fmaf(lg2f_(d_[u]) * kLn2, m_[u], td);
                prow[j] = m_[u];
            }
        }
        td = -warp_sum(td);
        if (a.bad_flag && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.bad_flag, 1);
        if (lane == 0) {
            a.td_err[r] = td;
            acc[0] = td * w;
        }
        if (a.grad_unit) {  // dense (N, n_atom) gradient block of the row: non-zero on the chosen action only
            const float c = -w / (float)a.R;
            float* __restrict__ gr = a.grad_unit + r * a.N * na;
            float g_[NJ];
#pragma unroll
            for (int u = 0; u < NJ; ++u) g_[u] = c * m_[u] * rcpf_(d_[u]);
            const int Ni = a.N;
            for (int n = 0; n < Ni; ++n) {
#pragma unroll
                for (int u = 0; u < NJ; ++u) {
                    const int j = lane + 32 * u;
                    if (j < na) gr[n * na + j] = (n == sel) ? g_[u] : 0.f;
                }
            }
        }
    }
    if (gridDim.x <= FX_MAX_GRID) {
        grid_sum_fx<1, NT>(acc, ws, [&](int, double t) { a.loss[0] = (float)(t / (double)a.R); });
        return;
    }
    double tot[1];
    if (grid_sum<1, NT>(acc, tot, ws, 0) && threadIdx.x == 0) a.loss[0] = (float)(tot[0] / (double)a.R);
}

__global__ void dntd_bwd_kernel(const float* __restrict__ dist, const long long* __restrict__ act,
                                const float* __restrict__ proj, const float* __restrict__ weight,
                                long long weight_stride, const float* __restrict__ g_loss,
                                const float* __restrict__ g_td, long long R, int N, int n_atom, int skip_if_unit,
                                float* __restrict__ grad_dist) {
    pdl_prologue();
    // skip_if_unit: the forward launch already wrote grad_dist for a unit upstream gradient -- verify and leave
    if (skip_if_unit && !g_td && g_loss && *g_loss == 1.f) return;
    const long long per_row = (long long)N * n_atom;
    const float g = g_loss ? *g_loss : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * per_row;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / per_row;
        const int rem = (int)(i - r * per_row);
        const int n = rem / n_atom, j = rem - n * n_atom;
        float out = 0.f;
        if (n == (int)act[r]) {
            const float w = weight ? weight[r * weight_stride] : 1.f;
            float c = g * w / (float)R;
            if (g_td) c += g_td[r];  // the unweighted per-sample error carries gradient too (td.py:519)
            out = -c * proj[r * n_atom + j] / dist[i];
        }
        grad_dist[i] = out;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// lambda-return scan along T on a (T, B) tile -- same three-phase tile scheme as gae.cu.
//   G_{T-1} = r + ((1-d)*gamma)*V_T ;  G_t = r_t + (1-d_t)*(disc_t*G_{t+1} + (gamma_t - disc_t)*V_{t+1}),  disc = gamma*lambda
// MODE 0: gamma/lambda scalars or (T,B) tensors, optional done (generalized_lambda_returns, td.py:1574-1651)
// MODE 1: UPGO: gamma = 1, lambda_t = [r_{t+1} + V_{t+2} >= V_{t+1}], last = 1 (upgo.py:66-68)
// HEAD 1: fused TD(lambda) loss head: loss = 0.5*mean(w*(G - V_t)^2) and the saved gradient dV (td.py:1570)
// ---------------------------------------------------------------------------------------------------------------
struct LamArgs {
    const float* value;   // (T+1, B)
    const float* reward;  // (T, B)
    const float* gammas;  // nullable (T, B)
    const float* lambdas; // nullable (T, B)
    const float* done;    // nullable (T, B)
    const float* weight;  // HEAD: nullable (T, B)
    float gamma, lambda;
    long long T, B;
    float* ret;     // (T, B) output (nullable when HEAD)
    float* loss;    // HEAD
    float* dvalue;  // HEAD: (T+1, B) saved d loss / d value for unit upstream gradient
};

template <int TC, int NT, int CHUNK, int MODE, int HEAD>
__global__ void __launch_bounds__(NT) lambda_scan_kernel(LamArgs a, float* ws) {
    pdl_prologue();
    // Chunks of CHUNK time steps, newest first, double-buffered: while the TC scan lanes run the dependent chain of chunk k
    // out of shared memory, every thread already has the global loads of chunk k+1 in flight (U elements per thread in
    // registers); they are turned into the scan's operands and written to the other buffer once the scan is done.  (Without
    // the overlap T = 1024, B = 64 took 46 us: eight times load latency + scan + store in a row.)
    constexpr int U = CHUNK * TC / NT;
    static_assert(U * NT == CHUNK * TC && U >= 1, "one pass per chunk");
    __shared__ float s_r[2][CHUNK][TC];
    __shared__ float s_m[2][CHUNK][TC];
    __shared__ float s_disc[2][CHUNK][TC];
    __shared__ float s_c[2][CHUNK][TC];
    // HEAD: V_t and the weight of the chunk, fetched with the scan operands one chunk ahead (single-buffered: a thread writes
    // the slots of chunk k+1 only after it has consumed the same slots of chunk k)
    __shared__ float s_hv[HEAD ? CHUNK : 1][TC];
    __shared__ float s_hw[HEAD ? CHUNK : 1][TC];
    const long long c0 = (long long)blockIdx.x * TC;
    const long long T = a.T, B = a.B;
    struct Raw {
        float rw[U], vn[U], g[U], l[U], dn[U], r2[U], v2[U], hv[U], hw[U];
        bool ok[U];
    };
    // all global loads of a chunk are issued before any of them is consumed
    auto fetch = [&](long long lo, int rows, Raw& x) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = threadIdx.x + u * NT;
            const int r = i / TC, cc = i % TC;
            const long long c = c0 + cc, t = lo + r;
            x.ok[u] = i < rows * TC && c < B;
            x.rw[u] = x.vn[u] = x.dn[u] = x.r2[u] = x.v2[u] = 0.f;
            x.g[u] = a.gamma;
            x.l[u] = a.lambda;
            if (x.ok[u]) {
                const long long off = t * B + c;
                x.rw[u] = a.reward[off];
                x.vn[u] = a.value[off + B];  // V_{t+1}
                if (MODE == 1) {
                    if (t < T - 1) {
                        x.r2[u] = a.reward[off + B];
                        x.v2[u] = a.value[off + 2 * B];
                    }
                } else {
                    if (a.gammas) x.g[u] = a.gammas[off];
                    if (a.lambdas) x.l[u] = a.lambdas[off];
                }
                if (a.done) x.dn[u] = a.done[off];
                if (HEAD == 1) {
                    x.hv[u] = a.value[off];
                    x.hw[u] = a.weight ? a.weight[off] : 1.f;
                }
            }
        }
    };
    auto commit = [&](int buf, long long lo, const Raw& x) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!x.ok[u]) continue;
            const int i = threadIdx.x + u * NT;
            const int r = i / TC, cc = i % TC;
            const long long t = lo + r;
            float gg = x.g[u], ll = x.l[u];
            if (MODE == 1) {
                gg = 1.f;
                ll = 1.f;
                if (t < T - 1) ll = (fadd(x.r2[u], x.v2[u]) >= x.vn[u]) ? 1.f : 0.f;
            }
            const float m = a.done ? fsub(1.f, x.dn[u]) : 1.f;
            const float disc = fmul(gg, ll);
            if (HEAD == 1) {
                s_hv[r][cc] = x.hv[u];
                s_hw[r][cc] = x.hw[u];
            }
            s_r[buf][r][cc] = x.rw[u];
            s_m[buf][r][cc] = m;
            if (t == T - 1) {
                // closed form of the last row kept in s_c; disc = 0 so the carry (0) is ignored exactly
                s_disc[buf][r][cc] = 0.f;
                s_c[buf][r][cc] = fmul(fmul(m, gg), x.vn[u]);
                s_m[buf][r][cc] = 1.f;
            } else {
                s_disc[buf][r][cc] = disc;
                s_c[buf][r][cc] = fmul(fsub(gg, disc), x.vn[u]);
            }
        }
    };
    float carry = 0.f;
    float acc[1] = {0.f};
    {
        const long long lo = T > CHUNK ? T - CHUNK : 0;
        Raw x;
        fetch(lo, (int)(T - lo), x);
        commit(0, lo, x);
    }
    __syncthreads();
    int buf = 0;
    for (long long hi = T; hi > 0; hi -= CHUNK, buf ^= 1) {
        const long long lo = hi > CHUNK ? hi - CHUNK : 0;
        const int rows = (int)(hi - lo);
        const bool have_next = lo > 0;
        const long long nlo = lo > CHUNK ? lo - CHUNK : 0;
        Raw nx;
        if (have_next) fetch(nlo, (int)(lo - nlo), nx);
        if (threadIdx.x < TC && c0 + threadIdx.x < B) {
            const int cc = threadIdx.x;
            // 16 rows at a time: the shared-memory operands of the next 16 steps are in registers before the dependent
            // chain needs them, so each step costs only its 4 dependent fp32 operations
            constexpr int UU = 16;
            int r = rows - 1;
            for (; r >= UU - 1; r -= UU) {
                float rr[UU], mm[UU], dd[UU], cq[UU];
#pragma unroll
                for (int j = 0; j < UU; ++j) {
                    rr[j] = s_r[buf][r - j][cc];
                    mm[j] = s_m[buf][r - j][cc];
                    dd[j] = s_disc[buf][r - j][cc];
                    cq[j] = s_c[buf][r - j][cc];
                }
#pragma unroll
                for (int j = 0; j < UU; ++j) {
                    carry = fadd(rr[j], fmul(mm[j], fadd(fmul(dd[j], carry), cq[j])));
                    s_r[buf][r - j][cc] = carry;
                }
            }
            for (; r >= 0; --r) {
                carry = fadd(s_r[buf][r][cc], fmul(s_m[buf][r][cc], fadd(fmul(s_disc[buf][r][cc], carry), s_c[buf][r][cc])));
                s_r[buf][r][cc] = carry;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = threadIdx.x + u * NT;
            const int r = i / TC, cc = i % TC;
            if (i >= rows * TC || c0 + cc >= B) continue;
            const long long off = (lo + r) * B + c0 + cc;
            const float gret = s_r[buf][r][cc];
            if (a.ret) a.ret[off] = gret;
            if (HEAD == 1) {
                const float w = s_hw[r][cc];
                const float d = gret - s_hv[r][cc];
                acc[0] += w * d * d;
                a.dvalue[off] = -w * d / (float)(T * B);  // 0.5 * w * 2 * (V - G) / count
            }
        }
        if (have_next) commit(buf ^ 1, nlo, nx);
        __syncthreads();
    }
    if (HEAD == 1) {
        // last value row receives no gradient (value[:-1], td.py:1570)
        for (int cc = threadIdx.x; cc < TC; cc += NT)
            if (c0 + cc < B) a.dvalue[T * B + c0 + cc] = 0.f;
        double tot[1];
        if (grid_sum<1, NT>(acc, tot, ws, 0) && threadIdx.x == 0)
            a.loss[0] = (float)(0.5 * tot[0] / ((double)T * (double)B));
    }
}

// Backward of the lambda-return scan (generalized_lambda_returns is differentiable in the reference, td.py:1574-1651; MBSAC
// and Dreamer back-propagate an actor loss through it): the transposed recurrence runs FORWARD in time,
//   a_t = gG_t + (1-d_{t-1})*disc_{t-1}*a_{t-1};  dr_t = a_t;  dV_{t+1} = a_t*(1-d_t)*(gamma_t - disc_t)  (last row: gamma_t)
// and, when the (T, B) gamma / lambda tensors want gradients, dgamma_t = a_t*(1-d_t)*(lambda_t*G_{t+1} + (1-lambda_t)*V_{t+1}),
// dlambda_t = a_t*(1-d_t)*gamma_t*(G_{t+1} - V_{t+1}).  Thread = column; every load of a step is independent of the carry.
struct LamBwdArgs {
    const float* g_ret;    // (T, B) upstream gradient
    const float* value;    // (T+1, B)
    const float* reward;   // (T, B)  (upgo mode: drives lambda)
    const float* ret;      // (T, B) forward result (needed for dgamma / dlambda only)
    const float* gammas;   // nullable
    const float* lambdas;  // nullable
    const float* done;     // nullable
    float gamma, lambda;
    int upgo_mode;
    long long T, B;
    float* g_value;    // (T+1, B)
    float* g_reward;   // nullable (T, B)
    float* g_gammas;   // nullable (T, B)
    float* g_lambdas;  // nullable (T, B)
};

template <int NT>
__global__ void __launch_bounds__(NT) lambda_returns_bwd_kernel(LamBwdArgs a) {
    pdl_prologue();
    const long long c = (long long)blockIdx.x * NT + threadIdx.x;
    if (c >= a.B) return;
    const long long T = a.T, B = a.B;
    a.g_value[c] = 0.f;
    float adj = 0.f, coef = 0.f;
    constexpr int U = 8;
    for (long long t0 = 0; t0 < T; t0 += U) {
        float g[U], gg[U], ll[U], m[U], vn[U], gn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {  // all loads of U steps first
            const long long t = t0 + u;
            g[u] = 0.f; gg[u] = a.gamma; ll[u] = a.lambda; m[u] = 1.f; vn[u] = 0.f; gn[u] = 0.f;
            if (t < T) {
                const long long off = t * B + c;
                g[u] = a.g_ret[off];
                if (a.upgo_mode) {
                    gg[u] = 1.f;
                    ll[u] = 1.f;
                    if (t < T - 1) ll[u] = (fadd(a.reward[off + B], a.value[off + 2 * B]) >= a.value[off + B]) ? 1.f : 0.f;
                } else {
                    if (a.gammas) gg[u] = a.gammas[off];
                    if (a.lambdas) ll[u] = a.lambdas[off];
                }
                if (a.done) m[u] = 1.f - a.done[off];
                if (a.g_gammas || a.g_lambdas) {
                    vn[u] = a.value[off + B];
                    gn[u] = (t < T - 1) ? a.ret[off + B] : 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long t = t0 + u;
            if (t >= T) break;
            const long long off = t * B + c;
            adj = fmaf(coef, adj, g[u]);
            if (a.g_reward) a.g_reward[off] = adj;
            const float am = adj * m[u];
            if (t == T - 1) {
                a.g_value[off + B] = am * gg[u];
                if (a.g_gammas) a.g_gammas[off] = am * vn[u];
                if (a.g_lambdas) a.g_lambdas[off] = 0.f;
            } else {
                const float disc = gg[u] * ll[u];
                a.g_value[off + B] = am * (gg[u] - disc);
                if (a.g_gammas) a.g_gammas[off] = am * (ll[u] * gn[u] + (1.f - ll[u]) * vn[u]);
                if (a.g_lambdas) a.g_lambdas[off] = am * gg[u] * (gn[u] - vn[u]);
                coef = m[u] * disc;
            }
        }
    }
}

// out[i] = (*g) * in[i]  -- backward of the heads that saved their unit-upstream gradient in the forward pass
__global__ void scale_kernel(const float* __restrict__ g, const float* __restrict__ in, float* __restrict__ out,
                             long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (*g) * in[i];
}

}  // namespace b200rl

using namespace b200rl;

static double qntd_loss_div(long long S, long long G, long long seq_len) {
    if (seq_len <= 0) return (double)S * (double)G;  // (td * weight).mean() over every row
    // sequence form: sum_t mean_b(...) / (len + 1e-8) -- a python float that torch narrows to fp32 (r2d2.py:364)
    return (double)(S / seq_len) * (double)G * (double)(float)((double)seq_len + 1e-8);
}

extern "C" int b200rl_qntd_fwd(const float* q, const float* next_n_q, const long long* action,
                               const long long* next_n_action, const float* reward, const float* done,
                               const float* weight, const float* value_gamma, long long value_gamma_stride,
                               const float* gamma_per_sample, long long S, long long G, long long N, int nstep,
                               double gamma, int cum_reward, int rescale, double rescale_eps, int criterion,
                               double criterion_param, int group_mean, long long seq_len, double priority_mix,
                               float* loss, float* td_error_per_sample, float* dcrit_saved, float* target_out,
                               float* grad_q_unit, float* priority_out, float* workspace, size_t workspace_bytes,
                               void* stream) {
    if (S <= 0 || G < 1 || N < 1 || nstep < 1 || !q || !next_n_q || !action || !next_n_action || !reward || !done ||
        !loss || !td_error_per_sample || !dcrit_saved || !workspace)
        return B200RL_ERR_ARG;
    if (criterion < 0 || criterion > 3) return B200RL_ERR_ARG;
    if (seq_len < 0 || (seq_len > 0 && (S % seq_len != 0 || cum_reward)) || (priority_out && seq_len <= 0) ||
        (priority_out && (G != 1 || group_mean)))
        return B200RL_ERR_ARG;
    QntdArgs a{};
    a.q = q; a.next_q = next_n_q; a.action = action; a.next_action = next_n_action; a.reward = reward; a.done = done;
    a.weight = weight; a.value_gamma = value_gamma; a.value_gamma_stride = value_gamma_stride;
    a.gamma_ps = gamma_per_sample; a.S = S; a.Bcol = seq_len > 0 ? S / seq_len : S; a.G = (int)G; a.N = (int)N;
    a.nstep = nstep; a.gamma = (float)gamma;
    a.gamma_pow_n = (float)pow(gamma, (double)nstep);  // python's `gamma ** nstep` (libm pow in double), then fp32
    a.cum_reward = cum_reward; a.rescale = rescale; a.eps = (float)rescale_eps;
    a.four_eps = (float)(4.0 * rescale_eps); a.two_eps = (float)(2.0 * rescale_eps);
    a.criterion = criterion; a.crit_param = (float)criterion_param; a.group_mean = group_mean;
    a.loss_div = qntd_loss_div(S, G, seq_len);
    a.prio_max_w = (float)priority_mix; a.prio_mean_w = (float)(1.0 - priority_mix);
    a.prio_div = (float)((double)seq_len + 1e-8);
    a.loss = loss; a.td_err = td_error_per_sample; a.dcrit = dcrit_saved; a.target = target_out;
    a.grad_unit = grad_q_unit; a.priority = priority_out;
    if (workspace_bytes < WS_MIN_BYTES) return B200RL_ERR_WORKSPACE;
    if (S <= 1024 && !priority_out) {  // the usual replay-buffer batch: ONE CTA, no grid reduction at all
        const int nt = S <= 256 ? 256 : (S <= 512 ? 512 : 1024);
        if (nt == 256) (void)launch_k(qntd_fwd_kernel<256>, 1, 256, 0, (cudaStream_t)stream, a, workspace);
        else if (nt == 512) (void)launch_k(qntd_fwd_kernel<512>, 1, 512, 0, (cudaStream_t)stream, a, workspace);
        else (void)launch_k(qntd_fwd_kernel<1024>, 1, 1024, 0, (cudaStream_t)stream, a, workspace);
        return (int)cudaGetLastError();
    }
    constexpr int NT = 128;
    const int grid = div_up(S, NT);
    if ((size_t)(WS_CTRL_WORDS + grid) > WS_PARTIAL_LIMIT_WORDS) return B200RL_ERR_WORKSPACE;
    (void)launch_k(qntd_fwd_kernel<NT>, grid, NT, 0, (cudaStream_t)stream, a, workspace);
    return (int)cudaGetLastError();
}

extern "C" int b200rl_qntd_bwd(const float* dcrit_saved, const float* weight, const long long* action,
                               const float* g_loss, const float* g_td, long long S, long long G, long long N,
                               int group_mean, long long seq_len, int skip_if_unit, float* grad_q, void* stream) {
    if (S <= 0 || G < 1 || N < 1 || !dcrit_saved || !action || !grad_q) return B200RL_ERR_ARG;
    long long grid = div_up(S * G * N, 256);
    if (grid > 148 * 8) grid = 148 * 8;  // grid-stride; the verification launch normally returns at once
    const float inv_div = (float)(1.0 / qntd_loss_div(S, G, seq_len));
    (void)launch_k(qntd_bwd_kernel, (int)grid, 256, 0, (cudaStream_t)stream, dcrit_saved, weight, action, g_loss, g_td, S,
                   (int)G, (int)N, group_mean, inv_div, skip_if_unit, grad_q);
    return (int)cudaGetLastError();
}

extern "C" int b200rl_dntd_fwd(const float* dist, const float* next_n_dist, const long long* act,
                               const long long* next_n_act, const float* reward, const float* done,
                               const float* weight, long long weight_stride, const float* value_gamma,
                               long long value_gamma_stride, const float* support, long long B, long long A,
                               long long N, int n_atom, int nstep, double gamma, double v_min, double v_max,
                               float* loss, float* td_error_per_sample, float* proj_saved, int* bad_flag,
                               float* grad_dist_unit, float* workspace, size_t workspace_bytes, void* stream) {
    if (B <= 0 || A < 1 || N < 1 || n_atom < 2 || nstep < 1 || !dist || !next_n_dist || !act || !next_n_act ||
        !reward || !done || !support || !loss || !td_error_per_sample || !proj_saved || !workspace)
        return B200RL_ERR_ARG;
    DntdArgs a{};
    a.dist = dist; a.next_dist = next_n_dist; a.act = act; a.next_act = next_n_act; a.reward = reward; a.done = done;
    a.weight = weight; a.weight_stride = weight_stride; a.value_gamma = value_gamma;
    a.value_gamma_stride = value_gamma_stride; a.support = support; a.R = B * A; a.A = A; a.B = B; a.N = (int)N;
    a.n_atom = n_atom; a.nstep = nstep; a.gamma = (float)gamma; a.gamma_pow_n = (float)pow(gamma, (double)nstep);
    a.v_min = (float)v_min; a.v_max = (float)v_max; a.delta_z = (float)((v_max - v_min) / (double)(n_atom - 1));
    a.loss = loss; a.td_err = td_error_per_sample; a.proj = proj_saved; a.bad_flag = bad_flag;
    a.grad_unit = grad_dist_unit;
    if (workspace_bytes < WS_MIN_BYTES) return B200RL_ERR_WORKSPACE;
    if (n_atom > 256) return B200RL_ERR_ARG;  // 8 atoms per lane in registers
    cudaStream_t st = (cudaStream_t)stream;
    if (a.R <= 8 * 511) {  // 8 rows per CTA: few enough CTAs for the one-round-trip reduction
        constexpr int NT = 256;
        const size_t sm = (size_t)(NT / 32) * n_atom * sizeof(float);
        const int grid = div_up(a.R, NT / 32);
        if (n_atom <= 64) (void)launch_k(dntd_fwd_kernel<NT, 2>, grid, NT, sm, st, a, workspace);
        else (void)launch_k(dntd_fwd_kernel<NT, 8>, grid, NT, sm, st, a, workspace);
        return (int)cudaGetLastError();
    }
    constexpr int NT = 128;
    const int grid = div_up(a.R, NT / 32);
    if ((size_t)(WS_CTRL_WORDS + grid) > WS_PARTIAL_LIMIT_WORDS) return B200RL_ERR_WORKSPACE;
    const size_t sm = (size_t)(NT / 32) * n_atom * sizeof(float);
    if (n_atom <= 64) (void)launch_k(dntd_fwd_kernel<NT, 2>, grid, NT, sm, st, a, workspace);
    else (void)launch_k(dntd_fwd_kernel<NT, 8>, grid, NT, sm, st, a, workspace);
    return (int)cudaGetLastError();
}

extern "C" int b200rl_dntd_bwd(const float* dist, const long long* act, const float* proj_saved, const float* weight,
                               long long weight_stride, const float* g_loss, const float* g_td, long long R,
                               long long N, int n_atom, int skip_if_unit, float* grad_dist, void* stream) {
    if (R <= 0 || N < 1 || n_atom < 2 || !dist || !act || !proj_saved || !grad_dist) return B200RL_ERR_ARG;
    long long grid = div_up(R * N * n_atom, 256);
    if (grid > 148 * 8) grid = 148 * 8;  // grid-stride; the verification launch normally returns at once
    (void)launch_k(dntd_bwd_kernel, (int)grid, 256, 0, (cudaStream_t)stream, dist, act, proj_saved, weight, weight_stride, g_loss, g_td,
                   R, (int)N, n_atom, skip_if_unit, grad_dist);
    return (int)cudaGetLastError();
}

template <int MODE, int HEAD>
static int launch_lambda(const LamArgs& a, float* ws, cudaStream_t st) {
    if (a.B >= 16 * 296) {
        (void)launch_k(lambda_scan_kernel<16, 256, 64, MODE, HEAD>, div_up(a.B, 16), 256, 0, st, a, ws);
    } else {
        (void)launch_k(lambda_scan_kernel<8, 256, 128, MODE, HEAD>, div_up(a.B, 8), 256, 0, st, a, ws);
    }
    return (int)cudaGetLastError();
}

extern "C" int b200rl_lambda_returns(const float* value, const float* reward, const float* gammas, double gamma,
                                     const float* lambdas, double lambda_, const float* done, int upgo_mode,
                                     long long T, long long B, float* ret, void* stream) {
    if (T <= 0 || B <= 0 || !value || !reward || !ret) return B200RL_ERR_ARG;
    LamArgs a{};
    a.value = value; a.reward = reward; a.gammas = gammas; a.lambdas = lambdas; a.done = done;
    a.gamma = (float)gamma; a.lambda = (float)lambda_; a.T = T; a.B = B; a.ret = ret;
    return upgo_mode ? launch_lambda<1, 0>(a, nullptr, (cudaStream_t)stream)
                     : launch_lambda<0, 0>(a, nullptr, (cudaStream_t)stream);
}

extern "C" int b200rl_lambda_returns_bwd(const float* g_ret, const float* value, const float* reward, const float* ret,
                                         const float* gammas, double gamma, const float* lambdas, double lambda_,
                                         const float* done, int upgo_mode, long long T, long long B, float* grad_value,
                                         float* grad_reward, float* grad_gammas, float* grad_lambdas, void* stream) {
    if (T <= 0 || B <= 0 || !g_ret || !value || !grad_value) return B200RL_ERR_ARG;
    if ((upgo_mode && !reward) || ((grad_gammas || grad_lambdas) && !ret)) return B200RL_ERR_ARG;
    LamBwdArgs a{};
    a.g_ret = g_ret; a.value = value; a.reward = reward; a.ret = ret; a.gammas = gammas; a.lambdas = lambdas;
    a.done = done; a.gamma = (float)gamma; a.lambda = (float)lambda_; a.upgo_mode = upgo_mode; a.T = T; a.B = B;
    a.g_value = grad_value; a.g_reward = grad_reward; a.g_gammas = grad_gammas; a.g_lambdas = grad_lambdas;
    constexpr int NT = 64;
    (void)launch_k(lambda_returns_bwd_kernel<NT>, div_up(B, NT), NT, 0, (cudaStream_t)stream, a);
    return (int)cudaGetLastError();
}

extern "C" int b200rl_td_lambda_fwd(const float* value, const float* reward, const float* weight, double gamma,
                                    double lambda_, long long T, long long B, float* loss, float* dvalue_saved,
                                    float* workspace, size_t workspace_bytes, void* stream) {
    if (T <= 0 || B <= 0 || !value || !reward || !loss || !dvalue_saved || !workspace) return B200RL_ERR_ARG;
    LamArgs a{};
    a.value = value; a.reward = reward; a.weight = weight; a.gamma = (float)gamma; a.lambda = (float)lambda_;
    a.T = T; a.B = B; a.loss = loss; a.dvalue = dvalue_saved;
    if (!ws_partials_fit((long long)(div_up(B, 8)), workspace_bytes)) return B200RL_ERR_WORKSPACE;
    return launch_lambda<0, 1>(a, workspace, (cudaStream_t)stream);
}

extern "C" int b200rl_scale(const float* g, const float* in, float* out, long long n, void* stream) {
    if (n < 0 || !g || (n > 0 && (!in || !out))) return B200RL_ERR_ARG;
    if (n == 0) return B200RL_OK;
    (void)launch_k(scale_kernel, div_up(n, 256), 256, 0, (cudaStream_t)stream, g, in, out, n);
    return (int)cudaGetLastError();
}
