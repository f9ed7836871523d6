// Column-tile learner step: GAE scan + ppo_error forward (+ gradients) in ONE launch with NO cross-CTA dependency.
//
// The recurrence of gae (ding/rl_utils/gae.py:65-69) runs along T only and ppo_error (ppo.py:77-140) is pointwise per
// transition, so a CTA that owns TC batch columns for ALL T can do both: it needs nothing from any other CTA.  In the
// time-major layout the transitions (t, c0..c0+TC) of one time step are contiguous in every tensor (TC*4 B in the (T, B)
// tensors, TC*8 B of actions, TC*N*4 B of logits), so a column tile is a stack of T such segments.
//
// Per CTA (256 threads, two CTAs per SM), for every column tile it owns (static stride over the grid) and every slab of
// 4*R time steps, newest slab first (R = 512/TC time steps per chunk, 512 transitions per chunk):
//   1. all threads load the slab's five GAE inputs with 16-byte loads (10 in flight per thread), mask next_value in
//      place where done != 0 (gae.py:61), form delta and f in the reference's operation order and park them in shared memory;
//   2. warp 0 (lane = column) runs the sequential scan A = delta + f*A with separate round-to-nearest mul and add
//      (bit-identical to the torch loop); the carry stays in its registers across slabs;
//   3. all threads write the slab's advantages to HBM (16-byte coalesced) -- and keep using them from shared memory;
//   4. the slab's PPO chunks stream through a two-stage shared-memory ring filled by 16-byte cp.async (LDGSTS) copies
//      issued by all threads one chunk ahead of the consumer; a thread computes two transitions per chunk
//      (ppo_row_compute_to, the same row code as ppo.cu), writes the gradient row IN PLACE over the logit_new row it
//      consumed, and every warp copies its 32-transition gradient block to HBM with 16-byte coalesced stores.
// The ring is filled across slab and tile boundaries, so the next tile's logits are in flight while this tile finishes.
// Loss partial sums stay in registers; each CTA stores six partials and finalize_sums_kernel adds them in a fixed order
// (deterministic: the tile -> CTA assignment is static).
//
// Algorithmic traffic: 24 B (GAE) + 104 B (ppo_error forward + gradients, N = 6) = 128 B per transition, each byte once.
#include "../../include/b200rl.h"
#include "fused_args.cuh"

namespace b200rl {

constexpr int CT_THREADS = 256;
constexpr int CT_ITEMS = 2 * CT_THREADS;  // transitions per chunk
constexpr int CT_STAGES = 2;
constexpr int CT_SLAB_CHUNKS = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// stage layout for CT_ITEMS transitions (no adv slot: the advantages live in the slab's scan buffer)
__host__ __device__ inline PpoTileLayout col_layout(int N, bool has_pre, bool has_w) {
    PpoTileLayout L;
    L.logit_bytes = CT_ITEMS * N * 4;
    int o = L.logit_bytes;
    L.off_old = o; o += L.logit_bytes;
    L.off_pre = o; if (has_pre) o += L.logit_bytes;
    L.off_act = o; o += CT_ITEMS * 8;
    L.off_vn = o; o += CT_ITEMS * 4;
    L.off_vo = o; o += CT_ITEMS * 4;
    L.off_adv = 0;
    L.off_ret = o; o += CT_ITEMS * 4;
    L.off_w = o; if (has_w) o += CT_ITEMS * 4;
    L.stage_bytes = (o + 127) & ~127;
    L.tx_bytes = o;
    return L;
}

// position of a CTA in its work list: column tile, slab (hi = exclusive top time step), chunk inside the slab
struct ColItem {
    long long tile;
    long long hi;
    int q;
};

template <int NC, bool GRADS, int TC>
__global__ void __launch_bounds__(CT_THREADS, 2) gae_ppo_col_kernel(FusedArgs f, float* ws) {
    pdl_prologue();
    constexpr int R = CT_ITEMS / TC;           // time steps per chunk
    constexpr int SLAB = R * CT_SLAB_CHUNKS;   // time steps per slab
    constexpr int TPR = TC / 4;                // threads per time step in the 16-byte passes over (T, B) tensors
    constexpr int RPP = CT_THREADS / TPR;      // time steps per such pass; SLAB == 2 * RPP
    static_assert(SLAB == 2 * RPP, "two GAE load passes per slab");
    extern __shared__ __align__(128) unsigned char smem[];
    const PpoArgs& a = f.p;
    const int N = NC ? NC : a.N;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const bool has_pre = a.logit_pre != nullptr, has_w = a.weight != nullptr;
    const PpoTileLayout L = col_layout(N, has_pre, has_w);
    auto s_d = reinterpret_cast<float (*)[TC]>(smem + CT_STAGES * L.stage_bytes);  // [SLAB][TC]: delta, then adv
    auto s_f = s_d + SLAB;                                                          // [SLAB][TC]
    const long long T = f.T, B = f.B;
    const long long n_tiles = (B + TC - 1) / TC;

    PpoUpstream up{0.f, 0.f, 0.f, 0.f, 1.f / (float)a.S};
    if (GRADS) {
        up.g_pol = a.g_policy ? *a.g_policy : 0.f;
        up.g_val = a.g_value ? *a.g_value : 0.f;
        up.g_ent = a.g_entropy ? *a.g_entropy : 0.f;
        up.g_kl = (a.g_kl && has_pre) ? *a.g_kl : 0.f;
        if (a.g_used && blockIdx.x == 0 && tid == 0) {
            a.g_used[0] = up.g_pol; a.g_used[1] = up.g_val; a.g_used[2] = up.g_ent; a.g_used[3] = up.g_kl;
        }
    }

    auto item_valid = [&](const ColItem& it) { return it.tile < n_tiles; };
    auto item_next = [&](ColItem& it) {
        const long long rows = it.hi < SLAB ? it.hi : SLAB;
        if ((long long)(++it.q) * R >= rows) {
            it.q = 0;
            it.hi -= SLAB;
            if (it.hi <= 0) {
                it.hi = T;
                it.tile += gridDim.x;
            }
        }
    };
    // 16-byte cp.async copies of one chunk into ring stage `sg`: for every tensor, R segments (one per time step) of the
    // tile's W columns
    auto issue = [&](const ColItem& it, int sg) {
        const long long c0 = it.tile * TC;
        const int W = (int)((B - c0) < TC ? (B - c0) : TC);
        const long long lo = it.hi > SLAB ? it.hi - SLAB : 0;
        const int rows = (int)(it.hi - lo);
        const int rbase = rows - (it.q + 1) * R;  // slab row of the chunk's oldest time step (may be < 0: ragged)
        unsigned char* st = smem + sg * L.stage_bytes;
        auto rows_of = [&](unsigned char* dst, const void* src, int elem_bytes) {
            // P 16-byte pieces per time step in shared memory, Pv of them valid
            const int P = TC * elem_bytes / 16, Pv = W * elem_bytes / 16;
            const unsigned char* g = reinterpret_cast<const unsigned char*>(src);
            for (int i = tid; i < R * P; i += CT_THREADS) {
                const int jj = i / P, o = i - jj * P;
                const int r = rbase + jj;
                if (r >= 0 && o < Pv)
                    cp_async16(dst + (size_t)i * 16, g + ((lo + r) * B + c0) * elem_bytes + (size_t)o * 16);
            }
        };
        rows_of(st, a.logit_new, N * 4);
        rows_of(st + L.off_old, a.logit_old, N * 4);
        if (has_pre) rows_of(st + L.off_pre, a.logit_pre, N * 4);
        rows_of(st + L.off_act, a.action, 8);
        rows_of(st + L.off_vn, a.value_new, 4);
        rows_of(st + L.off_vo, a.value_old, 4);
        rows_of(st + L.off_ret, a.ret, 4);
        if (has_w) rows_of(st + L.off_w, a.weight, 4);
    };

    ColItem cur{(long long)blockIdx.x, T, 0};
    ColItem pf = cur;
    float carry = 0.f;  // scan lanes of warp 0
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    // ---- GAE inputs of the first slab go out first, then the ring prologue ---------------------------------------------
    float4 gv[2], gn[2], gr[2], gd[2], gt[2];
    auto gae_issue = [&](const ColItem& it) {
        const long long c0 = it.tile * TC;
        const long long lo = it.hi > SLAB ? it.hi - SLAB : 0;
        const int rows = (int)(it.hi - lo);
        const int cq = (tid % TPR) * 4;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int r = p * RPP + tid / TPR;
            if (r < rows && c0 + cq < B) {
                const long long off = (lo + r) * B + c0 + cq;
                gv[p] = ldg_stream4(reinterpret_cast<const float4*>(f.value + off));
                gn[p] = ldg_stream4(reinterpret_cast<const float4*>(f.next_value + off));
                gr[p] = ldg_stream4(reinterpret_cast<const float4*>(f.reward + off));
                gd[p] = f.done ? ldg_stream4(reinterpret_cast<const float4*>(f.done + off))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
                gt[p] = f.traj ? ldg_stream4(reinterpret_cast<const float4*>(f.traj + off)) : gd[p];
            }
        }
    };
    if (item_valid(cur)) gae_issue(cur);
#pragma unroll
    for (int s = 0; s < CT_STAGES; ++s) {
        if (item_valid(pf)) {
            issue(pf, s);
            item_next(pf);
        }
        cp_async_commit();
    }

    for (int i = 0; item_valid(cur); ++i) {
        const int sg = i % CT_STAGES;
        const long long c0 = cur.tile * TC;
        const long long lo = cur.hi > SLAB ? cur.hi - SLAB : 0;
        const int rows = (int)(cur.hi - lo);
        if (cur.q == 0) {
            // ---- slab: delta / f -> shared memory, scan, advantages -> HBM ----------------------------------------------
            if (i != 0) gae_issue(cur);
            if (cur.hi == T) carry = 0.f;
            const int cq = (tid % TPR) * 4;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int r = p * RPP + tid / TPR;
                if (r < rows && c0 + cq < B) {
                    float vv[4] = {gv[p].x, gv[p].y, gv[p].z, gv[p].w}, nn[4] = {gn[p].x, gn[p].y, gn[p].z, gn[p].w};
                    float rw[4] = {gr[p].x, gr[p].y, gr[p].z, gr[p].w}, dd[4] = {gd[p].x, gd[p].y, gd[p].z, gd[p].w};
                    float tt[4] = {gt[p].x, gt[p].y, gt[p].z, gt[p].w};
                    float de[4], fa[4];
                    bool changed = false;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (f.done) {
                            changed |= (dd[k] != 0.f);
                            nn[k] = fmul(nn[k], fsub(1.f, dd[k]));
                        }
                        de[k] = fsub(fadd(rw[k], fmul(f.gamma, nn[k])), vv[k]);
                        fa[k] = fmul(f.gl, fsub(1.f, tt[k]));
                    }
                    *reinterpret_cast<float4*>(&s_d[r][cq]) = make_float4(de[0], de[1], de[2], de[3]);
                    *reinterpret_cast<float4*>(&s_f[r][cq]) = make_float4(fa[0], fa[1], fa[2], fa[3]);
                    if (changed && f.mask_inplace)
                        *reinterpret_cast<float4*>(f.next_value + (lo + r) * B + c0 + cq) =
                            make_float4(nn[0], nn[1], nn[2], nn[3]);
                }
            }
            __syncthreads();
            if (wid == 0 && lane < TC && c0 + lane < B) {
                int r = rows - 1;
                constexpr int U = 32;  // operands of 32 steps in registers before the dependent chain needs them
                for (; r >= U - 1; r -= U) {
                    float d[U], g[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        d[j] = s_d[r - j][lane];
                        g[j] = s_f[r - j][lane];
                    }
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        carry = fadd(d[j], fmul(g[j], carry));
                        s_d[r - j][lane] = carry;
                    }
                }
                for (; r >= 0; --r) {
                    carry = fadd(s_d[r][lane], fmul(s_f[r][lane], carry));
                    s_d[r][lane] = carry;
                }
            }
            __syncthreads();
            for (int k = tid; k < rows * TPR; k += CT_THREADS) {
                const int r = k / TPR, c4 = (k % TPR) * 4;
                if (c0 + c4 < B)
                    stg_stream4(reinterpret_cast<float4*>(a.adv_out + (lo + r) * B + c0 + c4),
                                *reinterpret_cast<const float4*>(&s_d[r][c4]));
            }
        }

        // ---- PPO chunk from ring stage sg ----------------------------------------------------------------------------
        cp_async_wait<CT_STAGES - 1>();
        __syncthreads();
        unsigned char* st = smem + sg * L.stage_bytes;
        const int rbase = rows - (cur.q + 1) * R;
        const int W = (int)((B - c0) < TC ? (B - c0) : TC);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = h * CT_THREADS + tid;
            const int jj = idx / TC, c = idx % TC;
            const int r = rbase + jj;
            if (r >= 0 && c < W) {
                float* grow = GRADS ? reinterpret_cast<float*>(st) + idx * N : nullptr;
                float* gval = GRADS ? reinterpret_cast<float*>(st + L.off_vn) + idx : nullptr;
                ppo_row_compute_to<NC, true, GRADS>(a, L, st, idx, N, s_d[r][c], grow, gval, up, acc);
            }
            if (GRADS) {
                // this warp's 32 transitions = 32/TC whole time-step segments: 8N float4 of logit gradients, 8 of value
                __syncwarp();
                const int idx0 = h * CT_THREADS + wid * 32;
                const float* src = reinterpret_cast<const float*>(st) + (size_t)idx0 * N;
                const int seg = TC * N;  // floats per time-step segment in shared memory
                for (int k = lane; k < 8 * N; k += 32) {
                    const int e = k * 4;
                    const int tr = e / seg, eo = e - tr * seg;
                    const int rr = rbase + idx0 / TC + tr;
                    if (rr >= 0 && eo < W * N)
                        stg_stream4(reinterpret_cast<float4*>(a.grad_logit + ((lo + rr) * B + c0) * N + eo),
                                    *reinterpret_cast<const float4*>(src + e));
                }
                if (lane < 8) {
                    const int e = lane * 4;
                    const int tr = e / TC, co = e % TC;
                    const int rr = rbase + idx0 / TC + tr;
                    if (rr >= 0 && co < W)
                        stg_stream4(reinterpret_cast<float4*>(a.grad_value + (lo + rr) * B + c0 + co),
                                    *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(st + L.off_vn) +
                                                                     idx0 + e));
                }
            }
        }
        __syncthreads();  // every read of stage sg is done: refill it with the chunk CT_STAGES ahead
        if (item_valid(pf)) {
            issue(pf, sg);
            item_next(pf);
        }
        cp_async_commit();
        item_next(cur);
    }
    cp_async_wait<0>();
    grid_store_partials<6, CT_THREADS>(acc, ws);  // summed by finalize_sums_kernel
}

bool coltile_ok(const FusedArgs& f) {
    const PpoArgs& a = f.p;
    const bool al = aligned16(a.logit_new) && aligned16(a.logit_old) && (!a.logit_pre || aligned16(a.logit_pre)) &&
                    aligned16(a.action) && aligned16(a.value_new) && aligned16(a.value_old) && aligned16(a.ret) &&
                    (!a.weight || aligned16(a.weight)) && (!a.grad_logit || aligned16(a.grad_logit)) &&
                    (!a.grad_value || aligned16(a.grad_value)) && aligned16(f.value) && aligned16(f.next_value) &&
                    aligned16(f.reward) && aligned16(a.adv) && (!f.done || aligned16(f.done)) &&
                    (!f.traj || aligned16(f.traj));
    if (!(al && a.N >= 1 && a.N <= 32 && f.T >= 1 && f.B >= 4 && (f.B % 4) == 0 && f.T * f.B == a.S)) return false;
    const PpoTileLayout L = col_layout(a.N, a.logit_pre != nullptr, a.weight != nullptr);
    return (size_t)CT_STAGES * L.stage_bytes + 2 * CT_ITEMS * CT_SLAB_CHUNKS * sizeof(float) <= 227 * 1024;
}

template <int NC, bool GRADS, int TC>
static int launch_col(const FusedArgs& f, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    const PpoArgs& a = f.p;
    const PpoTileLayout L = col_layout(a.N, a.logit_pre != nullptr, a.weight != nullptr);
    const size_t smem = (size_t)CT_STAGES * L.stage_bytes + 2 * CT_ITEMS * CT_SLAB_CHUNKS * sizeof(float);
    auto kern = gae_ppo_col_kernel<NC, GRADS, TC>;
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
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CT_THREADS, smem)) != cudaSuccess)
            return (int)e;
        occ_smem = smem;
    }
    if (per_sm < 1) return B200RL_ERR_ARG;
    const long long n_tiles = (f.B + TC - 1) / TC;
    long long grid = (long long)sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    if (ws_bytes < WS_MIN_BYTES || !ws_partials_fit((long long)(grid * 6), ws_bytes))
        return B200RL_ERR_WORKSPACE;
    (void)launch_k(kern, (int)grid, CT_THREADS, smem, st, f, ws);
    FinalizeArgs fa{};
    const double is = 1.0 / (double)a.S;
    fa.scale[0] = is; fa.scale[1] = 0.5 * is; fa.scale[2] = is; fa.scale[3] = a.logit_pre ? is : 0.0;
    fa.scale[4] = is; fa.scale[5] = is;
    fa.k = 6; fa.n_blocks = (int)grid;
    (void)launch_finalize(ws, out, fa, st);
    return (int)cudaGetLastError();
}

template <bool GRADS>
static int dispatch_col(const FusedArgs& f, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    // tile width: B200RL_COL_TC = 8 | 16 | 32 (tuning experiments, N = 6 only); default 16
    static int tc = -1;
    if (tc < 0) {
        const char* e = getenv("B200RL_COL_TC");
        tc = e ? atoi(e) : 16;
    }
    switch (f.p.N) {
        case 6:
            if (tc == 8) return launch_col<6, GRADS, 8>(f, out, ws, ws_bytes, st);
            if (tc == 32) return launch_col<6, GRADS, 32>(f, out, ws, ws_bytes, st);
            return launch_col<6, GRADS, 16>(f, out, ws, ws_bytes, st);
#define B200RL_CASE(n) \
    case n:            \
        return launch_col<n, GRADS, 16>(f, out, ws, ws_bytes, st);
        B200RL_CASE(2) B200RL_CASE(3) B200RL_CASE(4) B200RL_CASE(5) B200RL_CASE(7) B200RL_CASE(8) B200RL_CASE(9)
        B200RL_CASE(10) B200RL_CASE(12) B200RL_CASE(14) B200RL_CASE(16) B200RL_CASE(18)
#undef B200RL_CASE
        default:
            return launch_col<0, GRADS, 16>(f, out, ws, ws_bytes, st);
    }
}

int launch_coltile(const FusedArgs& f, bool grads, int variant, float* out, float* ws, size_t ws_bytes,
                   cudaStream_t st) {
    if ((variant == 0 || variant == 3) && colws_ok(f)) return launch_colws(f, grads, out, ws, ws_bytes, st);
    if (f.x_mailboxes) return B200RL_ERR_ARG;  // the fused exchange exists in colws.cu only
    if (variant == 2 && coltma_ok(f)) return launch_coltma(f, grads, out, ws, ws_bytes, st);
    return grads ? dispatch_col<true>(f, out, ws, ws_bytes, st) : dispatch_col<false>(f, out, ws, ws_bytes, st);
}

}  // namespace b200rl
