// PPO clipped surrogate + value + entropy (+ optional KL-to-pretrained) loss, forward and backward.
// Replaces ppo_error / ppo_policy_error / ppo_value_error of ding/rl_utils/ppo.py:77-275 (~40 torch kernels forward
// plus the autograd backward, and two .item() host syncs) with ONE forward and ONE backward kernel.
//
// Shapes: S samples, G "agent" rows per sample (G == 1 except the multi-agent case ppo.py:199-200,206-207),
// N logits per row.  logit_* are (S*G, N) row-major, action (S*G) int64, value_new/value_old/adv/return_/weight (S).
//
// Row mapping
//   STAGED (G == 1, N <= 32, 16B-aligned): a CTA of NT threads stages NT consecutive rows (NT*N floats, contiguous
//     in HBM) into shared memory with coalesced float4 loads, then thread i owns row i -- every byte of the logit
//     tensors crosses HBM once, fully coalesced.  Backward writes its gradient tile through shared memory the same way.
//   DIRECT L=1: one thread per sample, rows read straight from global (multi-agent / unaligned fallback).
//   DIRECT L=32: one warp per sample, lanes stride the row (large N, e.g. token vocabularies).
//
// Forward output: out[0..5] = policy_loss, value_loss, entropy_loss, kl_div, approx_kl, clipfrac (device floats, the
// caller decides when to read them -- no host sync in here).  Backward recomputes the softmax statistics from the
// inputs (cheaper than saving O(rows*N) state) and takes the four upstream gradients as device pointers.
#include "../../include/b200rl.h"
#include "common.cuh"

namespace b200rl {

struct PpoArgs {
    const float* logit_new;
    const float* logit_old;
    const float* logit_pre;  // nullable
    const long long* action;
    const float* value_new;
    const float* value_old;
    const float* adv;
    const float* ret;
    const float* weight;  // nullable -> 1
    long long S;
    int G;
    int N;
    float clip;       // fp32(clip_ratio)
    float clip_lo;    // fp32(1 - clip_ratio), computed in double like the python scalar of the reference
    float clip_hi;    // fp32(1 + clip_ratio)
    float dual_clip;  // <= 0: disabled
    int use_value_clip;
    int kl_type;  // 1,2,3
    // backward only
    const float* g_policy;
    const float* g_value;
    const float* g_entropy;
    const float* g_kl;
    float* grad_logit;
    float* grad_value;
};

// d(selected surrogate)/d(ratio) with torch's tie rules: min/max split the gradient 0.5/0.5 on equality, clamp passes
// gradient on the closed interval (ppo.py:208-216).  Also returns the selected surrogate value.
__device__ __forceinline__ float surrogate(float ratio, float adv, float lo, float hi, float dual_clip,
                                           float& dsel_dratio) {
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float s1 = ratio * adv, s2 = rc * adv;
    const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
    float w1, w2;
    if (s1 < s2) { w1 = 1.f; w2 = 0.f; }
    else if (s1 > s2) { w1 = 0.f; w2 = 1.f; }
    else { w1 = 0.5f; w2 = 0.5f; }
    float sel = fminf(s1, s2);
    float d = adv * (w1 + w2 * in_range);
    if (dual_clip > 0.f && adv < 0.f) {
        const float floor_ = dual_clip * adv;
        if (sel < floor_) { sel = floor_; d = 0.f; }
        else if (sel == floor_) { d *= 0.5f; }
    }
    dsel_dratio = d;
    return sel;
}

// 0.5*w*max(e1,e2) pieces: returns max(e1,e2) and d max / d value_new (ppo.py:267-274)
__device__ __forceinline__ float value_term(float v, float v_old, float ret, float clip, int use_clip, float& dterm_dv) {
    const float r1 = ret - v;
    const float e1 = r1 * r1;
    if (!use_clip) { dterm_dv = -2.f * r1; return e1; }
    const float dv = v - v_old;
    const float vc = v_old + fminf(fmaxf(dv, -clip), clip);
    const float r2 = ret - vc;
    const float e2 = r2 * r2;
    const float pass = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
    const float d1 = -2.f * r1, d2 = -2.f * r2 * pass;
    if (e1 > e2) { dterm_dv = d1; return e1; }
    if (e1 < e2) { dterm_dv = d2; return e2; }
    dterm_dv = 0.5f * (d1 + d2);
    return e1;
}

__device__ __forceinline__ float kl_term(float log_ratio, int kl_type, float& dterm) {
    if (kl_type == 1) { dterm = 1.f; return log_ratio; }
    if (kl_type == 2) { dterm = log_ratio; return log_ratio * log_ratio / 2.f; }
    const float e = expf(-log_ratio);
    dterm = 1.f - e;
    return e - 1.f + log_ratio;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// MODE 0: STAGED (thread per row via smem), 1: DIRECT thread per sample, 2: DIRECT warp per sample
template <int NT, int MODE>
__global__ void __launch_bounds__(NT) ppo_fwd_kernel(PpoArgs a, float* out, float* ws) {
    extern __shared__ __align__(16) float smem[];
    constexpr int L = (MODE == 2) ? 32 : 1;
    const int lane = (MODE == 2) ? (threadIdx.x & 31) : 0;
    const int N = a.N, G = a.G;
    long long s;  // sample handled by this thread / warp
    const float *zn = nullptr, *zo = nullptr, *zp = nullptr;  // row 0 of this sample
    if (MODE == 0) {
        const long long row0 = (long long)blockIdx.x * NT;
        const long long nrows = min((long long)NT, a.S - row0);
        const int nflt = (int)nrows * N;
        float* s_new = smem;
        float* s_old = smem + NT * N;
        float* s_pre = smem + 2 * NT * N;
        const float4* gn = reinterpret_cast<const float4*>(a.logit_new + row0 * N);
        const float4* go = reinterpret_cast<const float4*>(a.logit_old + row0 * N);
        const float4* gp = a.logit_pre ? reinterpret_cast<const float4*>(a.logit_pre + row0 * N) : nullptr;
        const int nv4 = nflt >> 2;
        for (int i = threadIdx.x; i < nv4; i += NT) {
            reinterpret_cast<float4*>(s_new)[i] = ldg_stream4(gn + i);
            reinterpret_cast<float4*>(s_old)[i] = ldg_stream4(go + i);
            if (gp) reinterpret_cast<float4*>(s_pre)[i] = ldg_stream4(gp + i);
        }
        for (int i = (nv4 << 2) + threadIdx.x; i < nflt; i += NT) {
            s_new[i] = a.logit_new[row0 * N + i];
            s_old[i] = a.logit_old[row0 * N + i];
            if (gp) s_pre[i] = a.logit_pre[row0 * N + i];
        }
        __syncthreads();
        s = row0 + threadIdx.x;
        zn = s_new + threadIdx.x * N;
        zo = s_old + threadIdx.x * N;
        zp = a.logit_pre ? s_pre + threadIdx.x * N : nullptr;
    } else if (MODE == 1) {
        s = (long long)blockIdx.x * NT + threadIdx.x;
    } else {
        s = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    }
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // policy, value, entropy, kl, approx_kl, clipfrac
    if (s < a.S) {
        if (MODE != 0) {
            zn = a.logit_new + s * G * N;
            zo = a.logit_old + s * G * N;
            zp = a.logit_pre ? a.logit_pre + s * G * N : nullptr;
        }
        float ratio_sum = 0.f, ent_sum = 0.f, akl = 0.f, kl = 0.f;
        for (int g = 0; g < G; ++g) {
            const float* rn = zn + (size_t)g * N;
            const float* ro = zo + (size_t)g * N;
            const int act = (int)a.action[s * G + g];
            float lse_n, ent;
            row_lse_entropy<L>([&](int j) { return rn[j]; }, N, lane, lse_n, ent);
            const float lse_o = row_lse<L>([&](int j) { return ro[j]; }, N, lane);
            const float lp_n = rn[act] - lse_n;
            const float lp_o = ro[act] - lse_o;
            ratio_sum += expf(lp_n - lp_o);
            ent_sum += ent;
            akl += lp_o - lp_n;
            if (zp) {
                const float* rp = zp + (size_t)g * N;
                const float lse_p = row_lse<L>([&](int j) { return rp[j]; }, N, lane);
                float dummy;
                kl += kl_term(lp_n - (rp[act] - lse_p), a.kl_type, dummy);
            }
        }
        if (lane == 0) {
            const float w = a.weight ? a.weight[s] : 1.f;
            const float adv = a.adv[s];
            const float ratio = (G == 1) ? ratio_sum : ratio_sum / (float)G;
            const float ent = (G == 1) ? ent_sum : ent_sum / (float)G;
            float dsel;
            const float sel = surrogate(ratio, adv, a.clip_lo, a.clip_hi, a.dual_clip, dsel);
            acc[0] = -sel * w;
            float dterm;
            acc[1] = value_term(a.value_new[s], a.value_old[s], a.ret[s], a.clip, a.use_value_clip, dterm) * w;
            acc[2] = ent * w;
            acc[3] = kl;
            acc[4] = akl;
            acc[5] = (ratio > a.clip_hi || ratio < a.clip_lo) ? 1.f : 0.f;
        }
    }
    double tot[6];
    if (grid_sum<6, NT>(acc, tot, ws, 0) && threadIdx.x == 0) {
        const double inv_s = 1.0 / (double)a.S, inv_m = 1.0 / ((double)a.S * (double)G);
        out[0] = (float)(tot[0] * inv_s);
        out[1] = (float)(0.5 * tot[1] * inv_s);
        out[2] = (float)(tot[2] * inv_s);
        out[3] = a.logit_pre ? (float)(tot[3] * inv_m) : 0.f;
        out[4] = (float)(tot[4] * inv_m);
        out[5] = (float)(tot[5] * inv_s);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
template <int NT, int MODE>
__global__ void __launch_bounds__(NT) ppo_bwd_kernel(PpoArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int L = (MODE == 2) ? 32 : 1;
    const int lane = (MODE == 2) ? (threadIdx.x & 31) : 0;
    const int N = a.N, G = a.G;
    const float g_pol = a.g_policy ? *a.g_policy : 0.f;
    const float g_val = a.g_value ? *a.g_value : 0.f;
    const float g_ent = a.g_entropy ? *a.g_entropy : 0.f;
    const float g_kl = (a.g_kl && a.logit_pre) ? *a.g_kl : 0.f;
    const float inv_s = 1.f / (float)a.S;
    const float inv_m = 1.f / ((float)a.S * (float)G);
    long long s;
    long long row0 = 0;
    int nflt = 0;
    const float *zn = nullptr, *zo = nullptr, *zp = nullptr;
    float* gz = nullptr;  // where this sample's gradient rows go (smem tile or global)
    if (MODE == 0) {
        row0 = (long long)blockIdx.x * NT;
        const long long nrows = min((long long)NT, a.S - row0);
        nflt = (int)nrows * N;
        float* s_new = smem;  // overwritten in place by the gradient
        float* s_old = smem + NT * N;
        float* s_pre = smem + 2 * NT * N;
        const float4* gn = reinterpret_cast<const float4*>(a.logit_new + row0 * N);
        const float4* go = reinterpret_cast<const float4*>(a.logit_old + row0 * N);
        const float4* gp = a.logit_pre ? reinterpret_cast<const float4*>(a.logit_pre + row0 * N) : nullptr;
        const int nv4 = nflt >> 2;
        for (int i = threadIdx.x; i < nv4; i += NT) {
            reinterpret_cast<float4*>(s_new)[i] = ldg_stream4(gn + i);
            reinterpret_cast<float4*>(s_old)[i] = ldg_stream4(go + i);
            if (gp) reinterpret_cast<float4*>(s_pre)[i] = ldg_stream4(gp + i);
        }
        for (int i = (nv4 << 2) + threadIdx.x; i < nflt; i += NT) {
            s_new[i] = a.logit_new[row0 * N + i];
            s_old[i] = a.logit_old[row0 * N + i];
            if (gp) s_pre[i] = a.logit_pre[row0 * N + i];
        }
        __syncthreads();
        s = row0 + threadIdx.x;
        zn = s_new + threadIdx.x * N;
        zo = s_old + threadIdx.x * N;
        zp = a.logit_pre ? s_pre + threadIdx.x * N : nullptr;
        gz = s_new + threadIdx.x * N;
    } else if (MODE == 1) {
        s = (long long)blockIdx.x * NT + threadIdx.x;
    } else {
        s = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    }
    if (s < a.S) {
        if (MODE != 0) {
            zn = a.logit_new + s * G * N;
            zo = a.logit_old + s * G * N;
            zp = a.logit_pre ? a.logit_pre + s * G * N : nullptr;
            gz = a.grad_logit + s * G * N;
        }
        const float w = a.weight ? a.weight[s] : 1.f;
        const float adv = a.adv[s];
        // pass A (only when G > 1): the sample's mean ratio decides the clip branch for all of its rows
        float ratio_s = 0.f;
        if (G > 1) {
            for (int g = 0; g < G; ++g) {
                const float* rn = zn + (size_t)g * N;
                const float* ro = zo + (size_t)g * N;
                const int act = (int)a.action[s * G + g];
                const float lse_n = row_lse<L>([&](int j) { return rn[j]; }, N, lane);
                const float lse_o = row_lse<L>([&](int j) { return ro[j]; }, N, lane);
                ratio_s += expf((rn[act] - lse_n) - (ro[act] - lse_o));
            }
            ratio_s /= (float)G;
        }
        for (int g = 0; g < G; ++g) {
            const float* rn = zn + (size_t)g * N;
            const float* ro = zo + (size_t)g * N;
            float* gr = gz + (size_t)g * N;
            const int act = (int)a.action[s * G + g];
            float lse_n, ent;
            row_lse_entropy<L>([&](int j) { return rn[j]; }, N, lane, lse_n, ent);
            const float lse_o = row_lse<L>([&](int j) { return ro[j]; }, N, lane);
            const float lp_n = rn[act] - lse_n;
            const float ratio_g = expf(lp_n - (ro[act] - lse_o));
            if (G == 1) ratio_s = ratio_g;
            float dsel;
            surrogate(ratio_s, adv, a.clip_lo, a.clip_hi, a.dual_clip, dsel);
            // d policy_loss / d logp_new(row) = -(w/S) * dsel/dratio * ratio_g / G
            float c_act = g_pol * (-w * inv_s) * dsel * ratio_g / (float)G;
            if (zp) {
                const float* rp = zp + (size_t)g * N;
                const float lse_p = row_lse<L>([&](int j) { return rp[j]; }, N, lane);
                float dk;
                kl_term(lp_n - (rp[act] - lse_p), a.kl_type, dk);
                c_act += g_kl * dk * inv_m;
            }
            const float c_ent = g_ent * w * inv_m;  // d entropy_loss / d H(row)
            // grad z_j = c_act*(1[j==a] - p_j) - c_ent * p_j*(logp_j + H)
            for (int j = lane; j < N; j += L) {
                const float lp = rn[j] - lse_n;
                const float p = expf(lp);
                float gj = -c_act * p - c_ent * p * (lp + ent);
                if (j == act) gj += c_act;
                gr[j] = gj;  // MODE 0: in place over the staged logit (each j is read before it is written)
            }
        }
        if (lane == 0) {
            float dterm;
            value_term(a.value_new[s], a.value_old[s], a.ret[s], a.clip, a.use_value_clip, dterm);
            a.grad_value[s] = g_val * 0.5f * w * inv_s * dterm;
        }
    }
    if (MODE == 0) {
        __syncthreads();
        float4* out4 = reinterpret_cast<float4*>(a.grad_logit + row0 * N);
        const int nv4 = nflt >> 2;
        for (int i = threadIdx.x; i < nv4; i += NT) stg_stream4(out4 + i, reinterpret_cast<const float4*>(smem)[i]);
        for (int i = (nv4 << 2) + threadIdx.x; i < nflt; i += NT) a.grad_logit[row0 * N + i] = smem[i];
    }
}

static int pick_mode(const PpoArgs& a) {
    const bool al = aligned16(a.logit_new) && aligned16(a.logit_old) && (!a.logit_pre || aligned16(a.logit_pre)) &&
                    (!a.grad_logit || aligned16(a.grad_logit));
    if (a.G == 1 && a.N <= 32 && al) return 0;
    if (a.N <= 64) return 1;
    return 2;
}

}  // namespace b200rl

using namespace b200rl;

static int check_ppo(const PpoArgs& a) {
    if (a.S < 0 || a.G < 1 || a.N < 1) return B200RL_ERR_ARG;
    if (!a.logit_new || !a.logit_old || !a.action || !a.value_new || !a.value_old || !a.adv || !a.ret)
        return B200RL_ERR_ARG;
    if (a.kl_type < 1 || a.kl_type > 3) return B200RL_ERR_ARG;
    return B200RL_OK;
}

extern "C" int b200rl_ppo_fwd(const float* logit_new, const float* logit_old, const float* logit_pretrained,
                              const long long* action, const float* value_new, const float* value_old,
                              const float* adv, const float* return_, const float* weight, long long S, long long G,
                              long long N, double clip_ratio, int use_value_clip, double dual_clip, int kl_type,
                              float* out, float* workspace, size_t workspace_bytes, void* stream) {
    PpoArgs a{};
    a.logit_new = logit_new; a.logit_old = logit_old; a.logit_pre = logit_pretrained; a.action = action;
    a.value_new = value_new; a.value_old = value_old; a.adv = adv; a.ret = return_; a.weight = weight;
    a.S = S; a.G = (int)G; a.N = (int)N; a.clip = (float)clip_ratio; a.clip_lo = (float)(1.0 - clip_ratio);
    a.clip_hi = (float)(1.0 + clip_ratio); a.dual_clip = (float)dual_clip;
    a.use_value_clip = use_value_clip; a.kl_type = kl_type;
    int rc = check_ppo(a);
    if (rc != B200RL_OK || !out || !workspace) return rc != B200RL_OK ? rc : B200RL_ERR_ARG;
    if (S == 0) return B200RL_ERR_ARG;  // mean over an empty batch is undefined (reference returns nan)
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int NT = 128;
    const int mode = pick_mode(a);
    const int grid = mode == 2 ? div_up(S, NT / 32) : div_up(S, NT);
    if ((size_t)(WS_CTRL_WORDS + (size_t)grid * 6) * sizeof(float) > workspace_bytes) return B200RL_ERR_WORKSPACE;
    if (mode == 0) {
        const size_t sm = (size_t)(a.logit_pre ? 3 : 2) * NT * a.N * sizeof(float);
        ppo_fwd_kernel<NT, 0><<<grid, NT, sm, st>>>(a, out, workspace);
    } else if (mode == 1) {
        ppo_fwd_kernel<NT, 1><<<grid, NT, 0, st>>>(a, out, workspace);
    } else {
        ppo_fwd_kernel<NT, 2><<<grid, NT, 0, st>>>(a, out, workspace);
    }
    return (int)cudaGetLastError();
}

extern "C" int b200rl_ppo_bwd(const float* logit_new, const float* logit_old, const float* logit_pretrained,
                              const long long* action, const float* value_new, const float* value_old,
                              const float* adv, const float* return_, const float* weight, long long S, long long G,
                              long long N, double clip_ratio, int use_value_clip, double dual_clip, int kl_type,
                              const float* g_policy, const float* g_value, const float* g_entropy, const float* g_kl,
                              float* grad_logit_new, float* grad_value_new, void* stream) {
    PpoArgs a{};
    a.logit_new = logit_new; a.logit_old = logit_old; a.logit_pre = logit_pretrained; a.action = action;
    a.value_new = value_new; a.value_old = value_old; a.adv = adv; a.ret = return_; a.weight = weight;
    a.S = S; a.G = (int)G; a.N = (int)N; a.clip = (float)clip_ratio; a.clip_lo = (float)(1.0 - clip_ratio);
    a.clip_hi = (float)(1.0 + clip_ratio); a.dual_clip = (float)dual_clip;
    a.use_value_clip = use_value_clip; a.kl_type = kl_type;
    a.g_policy = g_policy; a.g_value = g_value; a.g_entropy = g_entropy; a.g_kl = g_kl;
    a.grad_logit = grad_logit_new; a.grad_value = grad_value_new;
    int rc = check_ppo(a);
    if (rc != B200RL_OK) return rc;
    if (!grad_logit_new || !grad_value_new) return B200RL_ERR_ARG;
    if (S == 0) return B200RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int NT = 128;
    const int mode = pick_mode(a);
    const int grid = mode == 2 ? div_up(S, NT / 32) : div_up(S, NT);
    if (mode == 0) {
        const size_t sm = (size_t)(a.logit_pre ? 3 : 2) * NT * a.N * sizeof(float);
        ppo_bwd_kernel<NT, 0><<<grid, NT, sm, st>>>(a);
    } else if (mode == 1) {
        ppo_bwd_kernel<NT, 1><<<grid, NT, 0, st>>>(a);
    } else {
        ppo_bwd_kernel<NT, 2><<<grid, NT, 0, st>>>(a);
    }
    return (int)cudaGetLastError();
}
