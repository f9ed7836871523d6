// PPO clipped surrogate + value + entropy (+ optional KL-to-pretrained) loss, forward and backward.
// Replaces ppo_error / ppo_policy_error / ppo_value_error of ding/rl_utils/ppo.py:77-275 (~40 torch kernels forward
// plus the autograd backward, and two .item() host syncs).
//
// Shapes: S samples, G "agent" rows per sample (G == 1 except the multi-agent case ppo.py:199-200,206-207),
// N logits per row.  logit_* are (S*G, N) row-major, action (S*G) int64, value_new/value_old/adv/return_/weight (S).
//
// Main path (G == 1, N <= 32, 16-byte aligned tensors): ppo_tile_kernel
//   * persistent grid (SM count x resident CTAs), each CTA walks tiles of 128 consecutive rows;
//   * every input of a tile -- the logit rows (128*N contiguous floats per tensor), the int64 actions and the four
//     per-sample scalars -- arrives in shared memory by TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that complete
//     on an mbarrier; a 3-stage ring keeps two tiles in flight per CTA while one is consumed, so HBM latency is hidden
//     without spending issue slots on address arithmetic;
//   * thread i owns row i of the tile; N is a template parameter (rows live in registers, loops fully unrolled) and the
//     softmax statistics use ex2/lg2 approximations (relative error ~1e-7, far inside the 1e-5 parity bar);
//   * gradient tiles leave through shared memory and TMA bulk stores (cp.async.bulk.global.shared::cta);
//   * loss partial sums stay in registers across tiles; one deterministic grid reduction per CTA at the end.
//   Three variants of the same pipeline: FWD (losses), BWD (gradients for given upstream gradients) and FWD_GRAD: the
//   forward pass also writes the gradients for the upstream gradients it is told to expect (they are constants of the
//   training loop: policy + c_v*value - c_e*entropy), so the batch crosses HBM once; the backward launch then only
//   verifies the expectation on the device and recomputes nothing unless it was wrong (exact for any upstream value).
// Fallback paths: DIRECT L=1 one thread per sample (multi-agent / unaligned / N in 33..64), DIRECT L=32 one warp per
// sample (large N, e.g. token vocabularies).
//
// Forward output: out[0..5] = policy_loss, value_loss, entropy_loss, kl_div, approx_kl, clipfrac (device floats, the
// caller decides when to read them -- no host sync in here).
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
    // upstream gradients (device scalars, nullable = 0): actual ones for BWD, expected ones for FWD_GRAD
    const float* g_policy;
    const float* g_value;
    const float* g_entropy;
    const float* g_kl;
    float* grad_logit;
    float* grad_value;
    // FWD_GRAD: the 4 upstream values the gradients were scaled with are recorded here;
    // BWD: when non-null and equal to the actual upstream values the launch is a no-op (gradients already written)
    float* g_used;
    float* g_hint;  // BWD: refreshed with the actual upstream values for the next forward pass (nullable)
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

// ===============================================================================================================
// main path: persistent TMA-pipelined tile kernel
// ===============================================================================================================
constexpr int PPO_R = 128;      // rows per tile == consumer threads per CTA
constexpr int PPO_THREADS = PPO_R + 32;  // + one producer warp
constexpr int PPO_STAGES = 3;   // input ring depth
constexpr int PPO_OUTBUFS = 2;  // gradient tile ring depth (per warp)
enum { PPO_FWD = 0, PPO_FWD_GRAD = 1, PPO_BWD = 2 };

// MUFU approximations with flush-to-zero (no denormal fix-up code around them): relative error ~2^-22
__device__ __forceinline__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kF32Min = -3.402823466e38f;

__device__ __forceinline__ float kl_term(float log_ratio, int kl_type, float& dterm) {
    if (kl_type == 1) { dterm = 1.f; return log_ratio; }
    if (kl_type == 2) { dterm = log_ratio; return log_ratio * log_ratio / 2.f; }
    const float e = ex2f_(-log_ratio * kLog2e);
    dterm = 1.f - e;
    return e - 1.f + log_ratio;
}

// one row of NC logits from shared memory into registers; 8/16-byte vector loads are bank-conflict free for the
// row strides that occur (e.g. 24 B rows read as 3 x float2)
template <int NC>
__device__ __forceinline__ void load_row(const float* src, float (&z)[NC]) {
    if (NC % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) {
            const float4 v = reinterpret_cast<const float4*>(src)[j];
            z[4 * j] = v.x; z[4 * j + 1] = v.y; z[4 * j + 2] = v.z; z[4 * j + 3] = v.w;
        }
    } else if (NC % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) {
            const float2 v = reinterpret_cast<const float2*>(src)[j];
            z[2 * j] = v.x; z[2 * j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) z[j] = src[j];
    }
}
template <int NC>
__device__ __forceinline__ void store_row(float* dst, const float (&g)[NC]) {
    if (NC % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NC / 4; ++j)
            reinterpret_cast<float4*>(dst)[j] = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
    } else if (NC % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) reinterpret_cast<float2*>(dst)[j] = make_float2(g[2 * j], g[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) dst[j] = g[j];
    }
}

// Softmax statistics of one row in the log2 domain: t_j = (z_j - max)*log2(e) <= 0, e_j = 2^t_j, s = sum e_j,
// u2 = sum e_j t_j.  Then logsumexp = max + ln2*log2(s), entropy = ln2*(log2(s) - u2/s), p_j = e_j/s.
struct RowStat {
    float m, l2s, inv, log_s, ent;
};

struct PpoTileLayout {
    int logit_bytes;  // one logit tile
    int off_old, off_pre, off_act, off_vn, off_vo, off_adv, off_ret, off_w;
    int stage_bytes;
    int tx_bytes;  // bytes TMA delivers per stage
};
__host__ __device__ inline PpoTileLayout ppo_layout(int N, bool has_pre, bool has_w) {
    PpoTileLayout L;
    L.logit_bytes = PPO_R * N * 4;
    int o = L.logit_bytes;
    L.off_old = o; o += L.logit_bytes;
    L.off_pre = o; if (has_pre) o += L.logit_bytes;
    L.off_act = o; o += PPO_R * 8;
    L.off_vn = o; o += PPO_R * 4;
    L.off_vo = o; o += PPO_R * 4;
    L.off_adv = o; o += PPO_R * 4;
    L.off_ret = o; o += PPO_R * 4;
    L.off_w = o; if (has_w) o += PPO_R * 4;
    L.stage_bytes = (o + 127) & ~127;
    L.tx_bytes = o;
    return L;
}

template <int NC, int WHAT>
__global__ void __launch_bounds__(PPO_THREADS) ppo_tile_kernel(PpoArgs a, float* out, float* ws) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool GRADS = (WHAT != PPO_FWD);
    constexpr bool LOSSES = (WHAT != PPO_BWD);
    const int N = NC ? NC : a.N;
    const int tid = threadIdx.x;
    const int wid = tid >> 5, lane = tid & 31;
    const bool is_producer = wid == PPO_R / 32;  // warp 4: TMA issue only
    const bool has_pre = a.logit_pre != nullptr, has_w = a.weight != nullptr;
    const PpoTileLayout L = ppo_layout(N, has_pre, has_w);
    const int warp_out_bytes = 32 * N * 4;  // one warp's gradient rows of a tile
    unsigned char* outbuf = smem + PPO_STAGES * L.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(outbuf + (GRADS ? PPO_OUTBUFS * L.logit_bytes : 0));
    uint64_t* empty = full + PPO_STAGES;

    float g_pol = 0.f, g_val = 0.f, g_ent = 0.f, g_kl = 0.f;
    if (GRADS) {
        g_pol = a.g_policy ? *a.g_policy : 0.f;
        g_val = a.g_value ? *a.g_value : 0.f;
        g_ent = a.g_entropy ? *a.g_entropy : 0.f;
        g_kl = (a.g_kl && has_pre) ? *a.g_kl : 0.f;
        if (WHAT == PPO_BWD) {
            if (a.g_hint && blockIdx.x == 0 && tid == 0) {
                a.g_hint[0] = g_pol; a.g_hint[1] = g_val; a.g_hint[2] = g_ent; a.g_hint[3] = a.g_kl ? *a.g_kl : 0.f;
            }
            if (a.g_used) {  // gradients were already produced by the forward pass for exactly these upstream values?
                const bool same = __float_as_uint(a.g_used[0]) == __float_as_uint(g_pol) &&
                                  __float_as_uint(a.g_used[1]) == __float_as_uint(g_val) &&
                                  __float_as_uint(a.g_used[2]) == __float_as_uint(g_ent) &&
                                  (!has_pre || __float_as_uint(a.g_used[3]) == __float_as_uint(g_kl));
                if (same) return;
            }
        } else if (a.g_used && blockIdx.x == 0 && tid == 0) {
            a.g_used[0] = g_pol; a.g_used[1] = g_val; a.g_used[2] = g_ent; a.g_used[3] = g_kl;
        }
    }
    const float inv_s = 1.f / (float)a.S;

    const long long n_full = a.S / PPO_R;
    const int tail_rows = (int)(a.S - n_full * PPO_R);
    const long long n_tiles = n_full + (tail_rows ? 1 : 0);
    const int my_n = (n_tiles > blockIdx.x) ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (tid == 0) {
        for (int s = 0; s < PPO_STAGES; ++s) {
            mbar_init(&full[s], 1);           // producer's expect_tx arrive + TMA byte count
            mbar_init(&empty[s], PPO_R / 32);  // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // policy, value, entropy, kl, approx_kl, clipfrac
    if (is_producer) {
        // ---- producer warp: keeps the stage ring full; never touches the data ---------------------------------------
        if (lane == 0) {
            for (int i = 0; i < my_n; ++i) {
                const long long t = blockIdx.x + (long long)i * gridDim.x;
                if (t >= n_full) break;  // the ragged last tile is read with plain loads by its consumers
                const int sg = i % PPO_STAGES;
                if (i >= PPO_STAGES) mbar_wait(&empty[sg], (uint32_t)(((i / PPO_STAGES) - 1) & 1));
                const long long row0 = t * PPO_R;
                unsigned char* st = smem + sg * L.stage_bytes;
                uint64_t* bar = &full[sg];
                mbar_expect_tx(bar, (uint32_t)L.tx_bytes);
                tma_load_1d(st, a.logit_new + row0 * N, L.logit_bytes, bar);
                tma_load_1d(st + L.off_old, a.logit_old + row0 * N, L.logit_bytes, bar);
                if (has_pre) tma_load_1d(st + L.off_pre, a.logit_pre + row0 * N, L.logit_bytes, bar);
                tma_load_1d(st + L.off_act, a.action + row0, PPO_R * 8, bar);
                tma_load_1d(st + L.off_vn, a.value_new + row0, PPO_R * 4, bar);
                tma_load_1d(st + L.off_vo, a.value_old + row0, PPO_R * 4, bar);
                tma_load_1d(st + L.off_adv, a.adv + row0, PPO_R * 4, bar);
                tma_load_1d(st + L.off_ret, a.ret + row0, PPO_R * 4, bar);
                if (has_w) tma_load_1d(st + L.off_w, a.weight + row0, PPO_R * 4, bar);
            }
        }
    } else {
    // ---- consumer warps: warp w owns rows [32w, 32w+32) of every tile; no CTA-wide barrier in this loop ------------
    for (int i = 0; i < my_n; ++i) {
        const long long t = blockIdx.x + (long long)i * gridDim.x;
        const long long row0 = t * PPO_R;
        const int sg = i % PPO_STAGES;
        unsigned char* st = smem + sg * L.stage_bytes;
        const bool full_tile = t < n_full;
        if (full_tile) {
            mbar_wait(&full[sg], (uint32_t)((i / PPO_STAGES) & 1));
        } else if (tid < tail_rows) {
            // ragged last tile: every thread fetches its own row into its own slots of the stage (no sharing)
            float* d0 = reinterpret_cast<float*>(st) + tid * N;
            float* d1 = reinterpret_cast<float*>(st + L.off_old) + tid * N;
            float* d2 = reinterpret_cast<float*>(st + L.off_pre) + tid * N;
            for (int k = 0; k < N; ++k) {
                d0[k] = a.logit_new[(row0 + tid) * N + k];
                d1[k] = a.logit_old[(row0 + tid) * N + k];
                if (has_pre) d2[k] = a.logit_pre[(row0 + tid) * N + k];
            }
            reinterpret_cast<long long*>(st + L.off_act)[tid] = a.action[row0 + tid];
            reinterpret_cast<float*>(st + L.off_vn)[tid] = a.value_new[row0 + tid];
            reinterpret_cast<float*>(st + L.off_vo)[tid] = a.value_old[row0 + tid];
            reinterpret_cast<float*>(st + L.off_adv)[tid] = a.adv[row0 + tid];
            reinterpret_cast<float*>(st + L.off_ret)[tid] = a.ret[row0 + tid];
            if (has_w) reinterpret_cast<float*>(st + L.off_w)[tid] = a.weight[row0 + tid];
        }
        // this warp's slice of the gradient-tile ring (2 buffers per warp inside the CTA's output area)
        float* gtile = reinterpret_cast<float*>(outbuf + (wid * 2 + (i & 1)) * warp_out_bytes) - wid * 32 * N;
        if (full_tile || tid < tail_rows) {
            const float* zn = reinterpret_cast<const float*>(st) + tid * N;
            const float* zo = reinterpret_cast<const float*>(st + L.off_old) + tid * N;
            const int act = (int)reinterpret_cast<const long long*>(st + L.off_act)[tid];
            const float v_new = reinterpret_cast<const float*>(st + L.off_vn)[tid];
            const float v_old = reinterpret_cast<const float*>(st + L.off_vo)[tid];
            const float adv = reinterpret_cast<const float*>(st + L.off_adv)[tid];
            const float ret = reinterpret_cast<const float*>(st + L.off_ret)[tid];
            const float w = has_w ? reinterpret_cast<const float*>(st + L.off_w)[tid] : 1.f;
            constexpr int NR = NC ? NC : 1;
            float tn[NR], en[NR];  // new-policy row: t_j and e_j (compile-time N only)
            float m = kF32Min, s = 0.f, u2 = 0.f;
            if (NC) {
                load_row<NR>(zn, tn);
#pragma unroll
                for (int j = 0; j < NR; ++j) m = fmaxf(m, tn[j]);
                const float m2 = m * kLog2e;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    tn[j] = fmaxf(fmaf(tn[j], kLog2e, -m2), kF32Min);  // clamp: Categorical.entropy's finfo.min
                    en[j] = ex2f_(tn[j]);
                    s += en[j];
                    u2 = fmaf(en[j], tn[j], u2);
                }
            } else {
                for (int j = 0; j < N; ++j) m = fmaxf(m, zn[j]);
                const float m2 = m * kLog2e;
                for (int j = 0; j < N; ++j) {
                    const float t = fmaxf(fmaf(zn[j], kLog2e, -m2), kF32Min);
                    const float e = ex2f_(t);
                    s += e;
                    u2 = fmaf(e, t, u2);
                }
            }
            const float l2s = lg2f_(s), inv_sum = rcpf_(s);
            const float log_s = l2s * kLn2;
            const float ent = (l2s - u2 * inv_sum) * kLn2;
            const float lp_n = (zn[act] - m) - log_s;
            // behaviour ("old") policy row: only logsumexp is needed
            float mo = kF32Min, so = 0.f;
            if (NC) {
                float to[NR];
                load_row<NR>(zo, to);
#pragma unroll
                for (int j = 0; j < NR; ++j) mo = fmaxf(mo, to[j]);
                const float mo2 = mo * kLog2e;
#pragma unroll
                for (int j = 0; j < NR; ++j) so += ex2f_(fmaf(to[j], kLog2e, -mo2));
            } else {
                for (int j = 0; j < N; ++j) mo = fmaxf(mo, zo[j]);
                const float mo2 = mo * kLog2e;
                for (int j = 0; j < N; ++j) so += ex2f_(fmaf(zo[j], kLog2e, -mo2));
            }
            const float lp_o = (zo[act] - mo) - lg2f_(so) * kLn2;
            const float ratio = ex2f_((lp_n - lp_o) * kLog2e);
            float dsel, dterm, dk = 0.f, klv = 0.f;
            const float sel = surrogate(ratio, adv, a.clip_lo, a.clip_hi, a.dual_clip, dsel);
            const float vt = value_term(v_new, v_old, ret, a.clip, a.use_value_clip, dterm);
            if (has_pre) {
                const float* zp = reinterpret_cast<const float*>(st + L.off_pre) + tid * N;
                float mp = kF32Min, sp = 0.f;
                for (int j = 0; j < N; ++j) mp = fmaxf(mp, zp[j]);
                const float mp2 = mp * kLog2e;
                for (int j = 0; j < N; ++j) sp += ex2f_(fmaf(zp[j], kLog2e, -mp2));
                klv = kl_term(lp_n - ((zp[act] - mp) - lg2f_(sp) * kLn2), a.kl_type, dk);
            }
            if (LOSSES) {
                acc[0] -= sel * w;
                acc[1] += vt * w;
                acc[2] += ent * w;
                acc[3] += klv;
                acc[4] += lp_o - lp_n;
                acc[5] += (ratio > a.clip_hi || ratio < a.clip_lo) ? 1.f : 0.f;
            }
            if (GRADS) {
                // d/dlogp(a): policy -(w/S)*dsel*ratio, kl dk/S;  d/dH: entropy w/S
                const float c_act = g_pol * (-w * inv_s) * dsel * ratio + g_kl * dk * inv_s;
                const float c_ent = g_ent * w * inv_s;
                // grad z_j = c_act*(1[j==a] - p_j) - c_ent*p_j*(logp_j + H),  logp_j = ln2*t_j - log_s
                //          = p_j*(k0 - k1*t_j) + 1[j==a]*c_act
                const float k0 = -c_act - c_ent * (ent - log_s), k1 = c_ent * kLn2;
                float* gr = full_tile ? gtile + tid * N : a.grad_logit + (row0 + tid) * N;
                if (NC) {
                    float gj[NR];
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        gj[j] = (en[j] * inv_sum) * fmaf(-k1, tn[j], k0);
                        if (j == act) gj[j] += c_act;
                    }
                    if (full_tile) {
                        store_row<NR>(gr, gj);
                    } else {
#pragma unroll
                        for (int j = 0; j < NR; ++j) gr[j] = gj[j];
                    }
                } else {
                    const float m2 = m * kLog2e;
                    for (int j = 0; j < N; ++j) {
                        const float t = fmaxf(fmaf(zn[j], kLog2e, -m2), kF32Min);
                        float g = (ex2f_(t) * inv_sum) * fmaf(-k1, t, k0);
                        if (j == act) g += c_act;
                        gr[j] = g;
                    }
                }
                a.grad_value[row0 + tid] = g_val * 0.5f * w * inv_s * dterm;
            }
        }
        if (full_tile) {
            if (GRADS) {
                // hand this warp's 32 gradient rows to the TMA store engine; keep at most one store reading smem
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_1d(a.grad_logit + (row0 + wid * 32) * N, gtile + wid * 32 * N, warp_out_bytes);
                    tma_store_commit();
                    tma_store_wait_read<1>();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sg]);  // stage may be refilled once all four warps have arrived
        }
    }
    if (GRADS && lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the bulk stores that read it
    }
    if (LOSSES) {
        double tot[6];
        if (grid_sum<6, PPO_THREADS>(acc, tot, ws, 0) && tid == 0) {
            const double is = 1.0 / (double)a.S;
            out[0] = (float)(tot[0] * is);
            out[1] = (float)(0.5 * tot[1] * is);
            out[2] = (float)(tot[2] * is);
            out[3] = has_pre ? (float)(tot[3] * is) : 0.f;
            out[4] = (float)(tot[4] * is);
            out[5] = (float)(tot[5] * is);
        }
    }
}

// ===============================================================================================================
// fallback paths (multi-agent rows, N > 32, unaligned tensors): direct global loads
// ===============================================================================================================
// MODE 1: one thread per sample, 2: one warp per sample
template <int NT, int MODE>
__global__ void __launch_bounds__(NT) ppo_fwd_kernel(PpoArgs a, float* out, float* ws) {
    pdl_prologue();
    constexpr int L = (MODE == 2) ? 32 : 1;
    const int lane = (MODE == 2) ? (threadIdx.x & 31) : 0;
    const int N = a.N, G = a.G;
    const long long s = (MODE == 1) ? (long long)blockIdx.x * NT + threadIdx.x
                                    : (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // policy, value, entropy, kl, approx_kl, clipfrac
    if (s < a.S) {
        const float* zn = a.logit_new + s * G * N;
        const float* zo = a.logit_old + s * G * N;
        const float* zp = a.logit_pre ? a.logit_pre + s * G * N : nullptr;
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

template <int NT, int MODE>
__global__ void __launch_bounds__(NT) ppo_bwd_kernel(PpoArgs a) {
    pdl_prologue();
    constexpr int L = (MODE == 2) ? 32 : 1;
    const int lane = (MODE == 2) ? (threadIdx.x & 31) : 0;
    const int N = a.N, G = a.G;
    const float g_pol = a.g_policy ? *a.g_policy : 0.f;
    const float g_val = a.g_value ? *a.g_value : 0.f;
    const float g_ent = a.g_entropy ? *a.g_entropy : 0.f;
    const float g_kl = (a.g_kl && a.logit_pre) ? *a.g_kl : 0.f;
    const float inv_s = 1.f / (float)a.S;
    const float inv_m = 1.f / ((float)a.S * (float)G);
    const long long s = (MODE == 1) ? (long long)blockIdx.x * NT + threadIdx.x
                                    : (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (s >= a.S) return;
    const float* zn = a.logit_new + s * G * N;
    const float* zo = a.logit_old + s * G * N;
    const float* zp = a.logit_pre ? a.logit_pre + s * G * N : nullptr;
    float* gz = a.grad_logit + s * G * N;
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
            gr[j] = gj;
        }
    }
    if (lane == 0) {
        float dterm;
        value_term(a.value_new[s], a.value_old[s], a.ret[s], a.clip, a.use_value_clip, dterm);
        a.grad_value[s] = g_val * 0.5f * w * inv_s * dterm;
    }
}

static bool tile_path_ok(const PpoArgs& a) {
    const bool al = aligned16(a.logit_new) && aligned16(a.logit_old) && (!a.logit_pre || aligned16(a.logit_pre)) &&
                    aligned16(a.action) && aligned16(a.value_new) && aligned16(a.value_old) && aligned16(a.adv) &&
                    aligned16(a.ret) && (!a.weight || aligned16(a.weight)) &&
                    (!a.grad_logit || aligned16(a.grad_logit));
    return a.G == 1 && a.N <= 32 && al;
}

// launch geometry of the persistent kernel: SM count x resident CTAs per SM for this instantiation / smem size
template <int NC, int WHAT>
static int launch_tile(const PpoArgs& a, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    const PpoTileLayout L = ppo_layout(a.N, a.logit_pre != nullptr, a.weight != nullptr);
    const size_t smem = (size_t)PPO_STAGES * L.stage_bytes + (WHAT != PPO_FWD ? (size_t)PPO_OUTBUFS * L.logit_bytes : 0) +
                        2 * PPO_STAGES * sizeof(uint64_t);
    auto kern = ppo_tile_kernel<NC, WHAT>;
    static int sm_count = 0;
    static size_t smem_set = 0;
    cudaError_t e;
    if (sm_count == 0) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    }
    if (smem > 48 * 1024 && smem > smem_set) {
        if (smem > 227 * 1024) return B200RL_ERR_ARG;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return (int)e;
        smem_set = smem;
    }
    static size_t occ_smem = (size_t)-1;
    static int per_sm = 0;
    if (occ_smem != smem) {
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PPO_THREADS, smem)) != cudaSuccess)
            return (int)e;
        if (per_sm > 6) per_sm = 6;
        occ_smem = smem;
    }
    if (per_sm < 1) return B200RL_ERR_ARG;
    const long long n_tiles = (a.S + PPO_R - 1) / PPO_R;
    // the verification launch that follows a fused forward normally exits at once: keep its grid to one CTA per SM
    long long grid = (long long)sm_count * ((WHAT == PPO_BWD && a.g_used) ? 1 : per_sm);
    if (grid > n_tiles) grid = n_tiles;
    if (WHAT != PPO_BWD && (size_t)(WS_CTRL_WORDS + grid * 6) * sizeof(float) > ws_bytes) return B200RL_ERR_WORKSPACE;
    (void)launch_k(kern, (int)grid, PPO_THREADS, smem, st, a, out, ws);
    return (int)cudaGetLastError();
}

template <int WHAT>
static int dispatch_tile(const PpoArgs& a, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    switch (a.N) {
#define B200RL_CASE(n) case n: return launch_tile<n, WHAT>(a, out, ws, ws_bytes, st);
        B200RL_CASE(2) B200RL_CASE(3) B200RL_CASE(4) B200RL_CASE(5) B200RL_CASE(6) B200RL_CASE(7) B200RL_CASE(8)
        B200RL_CASE(9) B200RL_CASE(10) B200RL_CASE(12) B200RL_CASE(14) B200RL_CASE(16) B200RL_CASE(18)
#undef B200RL_CASE
        default: return launch_tile<0, WHAT>(a, out, ws, ws_bytes, st);
    }
}

}  // namespace b200rl

using namespace b200rl;

static int fill_args(PpoArgs& a, const float* logit_new, const float* logit_old, const float* logit_pretrained,
                     const long long* action, const float* value_new, const float* value_old, const float* adv,
                     const float* return_, const float* weight, long long S, long long G, long long N,
                     double clip_ratio, int use_value_clip, double dual_clip, int kl_type) {
    a.logit_new = logit_new; a.logit_old = logit_old; a.logit_pre = logit_pretrained; a.action = action;
    a.value_new = value_new; a.value_old = value_old; a.adv = adv; a.ret = return_; a.weight = weight;
    a.S = S; a.G = (int)G; a.N = (int)N; a.clip = (float)clip_ratio; a.clip_lo = (float)(1.0 - clip_ratio);
    a.clip_hi = (float)(1.0 + clip_ratio); a.dual_clip = (float)dual_clip;
    a.use_value_clip = use_value_clip; a.kl_type = kl_type;
    if (S < 0 || G < 1 || N < 1) return B200RL_ERR_ARG;
    if (!logit_new || !logit_old || !action || !value_new || !value_old || !adv || !return_) return B200RL_ERR_ARG;
    if (kl_type < 1 || kl_type > 3) return B200RL_ERR_ARG;
    return B200RL_OK;
}

extern "C" int b200rl_ppo_fwd(const float* logit_new, const float* logit_old, const float* logit_pretrained,
                              const long long* action, const float* value_new, const float* value_old,
                              const float* adv, const float* return_, const float* weight, long long S, long long G,
                              long long N, double clip_ratio, int use_value_clip, double dual_clip, int kl_type,
                              float* out, float* workspace, size_t workspace_bytes, void* stream) {
    PpoArgs a{};
    int rc = fill_args(a, logit_new, logit_old, logit_pretrained, action, value_new, value_old, adv, return_, weight,
                       S, G, N, clip_ratio, use_value_clip, dual_clip, kl_type);
    if (rc != B200RL_OK || !out || !workspace) return rc != B200RL_OK ? rc : B200RL_ERR_ARG;
    if (S == 0) return B200RL_ERR_ARG;  // mean over an empty batch is undefined (reference returns nan)
    cudaStream_t st = (cudaStream_t)stream;
    if (tile_path_ok(a)) return dispatch_tile<PPO_FWD>(a, out, workspace, workspace_bytes, st);
    constexpr int NT = 128;
    const bool warp = a.N > 64;
    const int grid = warp ? div_up(S, NT / 32) : div_up(S, NT);
    if ((size_t)(WS_CTRL_WORDS + (size_t)grid * 6) * sizeof(float) > workspace_bytes) return B200RL_ERR_WORKSPACE;
    if (warp) (void)launch_k(ppo_fwd_kernel<NT, 2>, grid, NT, 0, st, a, out, workspace);
    else (void)launch_k(ppo_fwd_kernel<NT, 1>, grid, NT, 0, st, a, out, workspace);
    return (int)cudaGetLastError();
}

extern "C" int b200rl_ppo_fwd_grad(const float* logit_new, const float* logit_old, const float* logit_pretrained,
                                   const long long* action, const float* value_new, const float* value_old,
                                   const float* adv, const float* return_, const float* weight, long long S,
                                   long long G, long long N, double clip_ratio, int use_value_clip, double dual_clip,
                                   int kl_type, const float* g_expected, float* g_used, float* out,
                                   float* grad_logit_new, float* grad_value_new, float* workspace,
                                   size_t workspace_bytes, void* stream) {
    PpoArgs a{};
    int rc = fill_args(a, logit_new, logit_old, logit_pretrained, action, value_new, value_old, adv, return_, weight,
                       S, G, N, clip_ratio, use_value_clip, dual_clip, kl_type);
    if (rc != B200RL_OK) return rc;
    if (!out || !workspace || !g_expected || !g_used || !grad_logit_new || !grad_value_new || S == 0)
        return B200RL_ERR_ARG;
    a.g_policy = g_expected; a.g_value = g_expected + 1; a.g_entropy = g_expected + 2; a.g_kl = g_expected + 3;
    a.g_used = g_used; a.grad_logit = grad_logit_new; a.grad_value = grad_value_new;
    if (!tile_path_ok(a)) return B200RL_ERR_ARG;  // callers probe with b200rl_ppo_fused_supported first
    return dispatch_tile<PPO_FWD_GRAD>(a, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int b200rl_ppo_fused_supported(const float* logit_new, const float* logit_old,
                                          const float* logit_pretrained, const long long* action,
                                          const float* value_new, const float* value_old, const float* adv,
                                          const float* return_, const float* weight, const float* grad_logit_new,
                                          long long G, long long N) {
    PpoArgs a{};
    a.logit_new = logit_new; a.logit_old = logit_old; a.logit_pre = logit_pretrained; a.action = action;
    a.value_new = value_new; a.value_old = value_old; a.adv = adv; a.ret = return_; a.weight = weight;
    a.grad_logit = const_cast<float*>(grad_logit_new); a.G = (int)G; a.N = (int)N;
    return tile_path_ok(a) ? 1 : 0;
}

extern "C" int b200rl_ppo_bwd(const float* logit_new, const float* logit_old, const float* logit_pretrained,
                              const long long* action, const float* value_new, const float* value_old,
                              const float* adv, const float* return_, const float* weight, long long S, long long G,
                              long long N, double clip_ratio, int use_value_clip, double dual_clip, int kl_type,
                              const float* g_policy, const float* g_value, const float* g_entropy, const float* g_kl,
                              const float* g_used, float* g_hint, float* grad_logit_new, float* grad_value_new,
                              void* stream) {
    PpoArgs a{};
    int rc = fill_args(a, logit_new, logit_old, logit_pretrained, action, value_new, value_old, adv, return_, weight,
                       S, G, N, clip_ratio, use_value_clip, dual_clip, kl_type);
    if (rc != B200RL_OK) return rc;
    a.g_policy = g_policy; a.g_value = g_value; a.g_entropy = g_entropy; a.g_kl = g_kl;
    a.g_used = const_cast<float*>(g_used); a.g_hint = g_hint;
    a.grad_logit = grad_logit_new; a.grad_value = grad_value_new;
    if (!grad_logit_new || !grad_value_new) return B200RL_ERR_ARG;
    if (S == 0) return B200RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (tile_path_ok(a)) return dispatch_tile<PPO_BWD>(a, nullptr, nullptr, 0, st);
    if (g_used) return B200RL_ERR_ARG;  // the fused forward only exists on the tile path
    constexpr int NT = 128;
    if (a.N > 64) (void)launch_k(ppo_bwd_kernel<NT, 2>, div_up(S, NT / 32), NT, 0, st, a);
    else (void)launch_k(ppo_bwd_kernel<NT, 1>, div_up(S, NT), NT, 0, st, a);
    return (int)cudaGetLastError();
}
