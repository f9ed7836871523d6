// Column-tile learner step, TMA edition: gae -> ppo_error (+ gradients) in ONE launch, every byte moved by 2-D tensor-map
// TMA copies (cp.async.bulk.tensor -> SASS UTMALDG / UTMASTG), so the SM's issue slots are spent on the row math only.
// (coltile.cu is the same algorithm with per-thread 16-byte cp.async copies; it took 19.7 warp instructions per
// transition, half of them address arithmetic of the copies, and was issue-bound at 54 % issue-active.)
//
// Decomposition as in coltile.cu: a CTA owns TC = 16 batch columns for ALL T (the recurrence of gae.py:65-69 runs along T
// only; ppo.py:77-140 is pointwise), so no CTA ever waits for another.  In the time-major layout the tile is a 2-D box of
// every tensor: (T, B) tensors are [T][B] with a box of R x 16 elements, logits are [T][B*N] with a box of R x 16N.
//
// Warp roles (10 warps, two CTAs per SM):
//   warp 8, lane 0  producer: keeps a two-deep ring of GAE input chunks (5 x R x 16 fp32) and a two-deep ring of PPO
//                   chunks (logit_new | logit_old | action | value_new | value_old | return_ [| weight | logit_pre] for
//                   R x 16 = 512 transitions) full; when the consumers have finished a chunk it stores the chunk's gradient
//                   boxes (written in place over logit_new / value_new) with TMA and refills the stage.
//   warp 9          scanner: per chunk, newest first: delta / f from the raw inputs in the reference's operation order, the
//                   in-place next_value mask (gae.py:61), then the sequential scan A = delta + f*A (lane = column,
//                   separate round-to-nearest mul and add: bit-identical to the torch loop); publishes the chunk's
//                   advantages in shared memory (mbarrier) and TMA-stores the slab (4 chunks) when it is complete.
//   warps 0..7      consumers: wait for "PPO chunk landed" and "advantages of this chunk ready", compute two transitions
//                   per thread (ppo_row_compute_to: the row code of ppo.cu), gradients in place, arrive on the chunk's
//                   "done" barrier.  No CTA-wide barrier anywhere in the loop.
// Rings run across slab and tile boundaries (static tile -> CTA assignment, so the loss partial sums are deterministic).
// Ragged edges (T % R, B % 16) are out-of-bound box coordinates: TMA zero-fills loads and clips stores.
//
// Algorithmic traffic: 24 B (GAE) + 104 B (ppo_error forward + gradients, N = 6) = 128 B per transition, each byte once.
#include <cuda.h>
#include <string.h>

#include "../../include/b200rl.h"
#include "fused_args.cuh"

namespace b200rl {

constexpr int TM_CW = 8;                       // consumer warps
constexpr int TM_CT = TM_CW * 32;              // consumer threads
constexpr int TM_THREADS = TM_CT + 64;         // + producer warp + scanner warp
constexpr int TM_ITEMS = 2 * TM_CT;            // transitions per chunk
constexpr int TM_STAGES = 2;
constexpr int TM_SLAB_CHUNKS = 4;
constexpr int TM_TC = 16;                      // columns per tile
constexpr int TM_R = TM_ITEMS / TM_TC;         // time steps per chunk (32)
constexpr int TM_SLAB = TM_R * TM_SLAB_CHUNKS; // time steps per slab (128)
constexpr int TM_RAW_ARR = TM_R * TM_TC * 4;   // bytes of one raw GAE array chunk
constexpr int TM_RAW_BYTES = 5 * TM_RAW_ARR;   // value | next_value | reward | done | traj_flag

struct ColMaps {
    CUtensorMap ln, lo, lp, act, vn, vo, ret, w;  // PPO chunk loads
    CUtensorMap gl, gv;                           // gradient stores
    CUtensorMap value, nv, reward, done, traj;    // GAE chunk loads
    CUtensorMap adv;                              // advantage slab store
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(m), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, int x, int y, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(m), "r"(x),
                 "r"(y), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

// timeline instrumentation for tuning (B200RL_FUSED_TRACE=1): 32 globaltimer stamps per CTA at workspace word 65536
// (tools/trace_col.py): 0 start | consumers, chunk j < 6: 1+3j landed, 2+3j advantages ready, 3+3j computed |
// 19 scanner published its first chunk | producer, chunk j < 6: 20+2j saw "done", 21+2j refilled the stage
#define TM_TRACE(slot)                                                                                   \
    do {                                                                                                 \
        if (f.trace) reinterpret_cast<unsigned long long*>(ws + 65536)[blockIdx.x * 32 + (slot)] = gtimer(); \
    } while (0)

struct TmItem {
    long long tile;
    long long hi;  // exclusive top time step of the slab
    int q;         // chunk inside the slab, 0 = newest
};

template <int NC, bool GRADS>
__global__ void __launch_bounds__(TM_THREADS, 2)
gae_ppo_tma_kernel(const __grid_constant__ ColMaps m, FusedArgs f, float* ws) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem[];
    const PpoArgs& a = f.p;
    const int N = NC ? NC : a.N;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const bool has_pre = a.logit_pre != nullptr, has_w = a.weight != nullptr;
    // stage layout: same fields as ppo.cu's (no adv slot)
    PpoTileLayout L;
    {
        L.logit_bytes = TM_ITEMS * N * 4;
        int o = L.logit_bytes;
        L.off_old = o; o += L.logit_bytes;
        L.off_pre = o; if (has_pre) o += L.logit_bytes;
        L.off_act = o; o += TM_ITEMS * 8;
        L.off_vn = o; o += TM_ITEMS * 4;
        L.off_vo = o; o += TM_ITEMS * 4;
        L.off_adv = 0;
        L.off_ret = o; o += TM_ITEMS * 4;
        L.off_w = o; if (has_w) o += TM_ITEMS * 4;
        L.stage_bytes = (o + 127) & ~127;
        L.tx_bytes = o;
    }
    unsigned char* raw = smem + TM_STAGES * L.stage_bytes;                               // [2][5][R][TC]
    auto advb = reinterpret_cast<float (*)[TM_SLAB][TM_TC]>(raw + 2 * TM_RAW_BYTES);    // [2][SLAB][TC]
    auto fbuf = reinterpret_cast<float (*)[TM_TC]>(reinterpret_cast<unsigned char*>(advb) + 2 * TM_SLAB * TM_TC * 4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(fbuf) + TM_R * TM_TC * 4);
    uint64_t* full = bars;                   // [2]  PPO chunk landed (tx)
    uint64_t* done = bars + 2;               // [2]  consumers finished the chunk (TM_CW arrivals)
    uint64_t* graw_full = bars + 4;          // [2]  GAE raw chunk landed (tx)
    uint64_t* adv_ready = bars + 6;          // [2][4] advantages of (slab parity, chunk) are in shared memory

    const long long T = f.T, B = f.B;
    const long long n_tiles = (B + TM_TC - 1) / TM_TC;
    const bool has_done = f.done != nullptr, has_traj = f.traj != nullptr;

    if (tid == 0) {
        TM_TRACE(0);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], TM_CW);
            mbar_init(&graw_full[s], 1);
        }
        for (int s = 0; s < 8; ++s) mbar_init(&adv_ready[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto item_valid = [&](const TmItem& it) { return it.tile < n_tiles; };
    auto item_next = [&](TmItem& it) {
        const long long rows = it.hi < TM_SLAB ? it.hi : TM_SLAB;
        if ((long long)(++it.q) * TM_R >= rows) {
            it.q = 0;
            it.hi -= TM_SLAB;
            if (it.hi <= 0) {
                it.hi = T;
                it.tile += gridDim.x;
            }
        }
    };
    auto last_in_slab = [&](const TmItem& it) {
        const long long rows = it.hi < TM_SLAB ? it.hi : TM_SLAB;
        return (long long)(it.q + 1) * TM_R >= rows;
    };
    const TmItem first{(long long)blockIdx.x, T, 0};
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    if (wid == TM_CW) {
        // =============================================== producer =========================================================
        if (lane == 0) {
            prefetch_tmap(&m.ln); prefetch_tmap(&m.lo); prefetch_tmap(&m.act); prefetch_tmap(&m.vn);
            prefetch_tmap(&m.vo); prefetch_tmap(&m.ret); prefetch_tmap(&m.value); prefetch_tmap(&m.nv);
            prefetch_tmap(&m.reward);
            const uint32_t raw_tx = (uint32_t)((3 + (has_done ? 1 : 0) + (has_traj ? 1 : 0)) * TM_RAW_ARR);
            auto issue_raw = [&](const TmItem& it, int slot) {
                const int x = (int)(it.tile * TM_TC), y = (int)(it.hi - (long long)(it.q + 1) * TM_R);
                unsigned char* dst = raw + slot * TM_RAW_BYTES;
                uint64_t* bar = &graw_full[slot];
                mbar_expect_tx(bar, raw_tx);
                tma_load_2d(dst, &m.value, x, y, bar);
                tma_load_2d(dst + TM_RAW_ARR, &m.nv, x, y, bar);
                tma_load_2d(dst + 2 * TM_RAW_ARR, &m.reward, x, y, bar);
                if (has_done) tma_load_2d(dst + 3 * TM_RAW_ARR, &m.done, x, y, bar);
                if (has_traj) tma_load_2d(dst + 4 * TM_RAW_ARR, &m.traj, x, y, bar);
            };
            auto issue_ppo = [&](const TmItem& it, int sg) {
                const int c0 = (int)(it.tile * TM_TC), y = (int)(it.hi - (long long)(it.q + 1) * TM_R);
                unsigned char* st = smem + sg * L.stage_bytes;
                uint64_t* bar = &full[sg];
                mbar_expect_tx(bar, (uint32_t)L.tx_bytes);
                tma_load_2d(st, &m.ln, c0 * N, y, bar);
                tma_load_2d(st + L.off_old, &m.lo, c0 * N, y, bar);
                if (has_pre) tma_load_2d(st + L.off_pre, &m.lp, c0 * N, y, bar);
                tma_load_2d(st + L.off_act, &m.act, c0, y, bar);
                tma_load_2d(st + L.off_vn, &m.vn, c0, y, bar);
                tma_load_2d(st + L.off_vo, &m.vo, c0, y, bar);
                tma_load_2d(st + L.off_ret, &m.ret, c0, y, bar);
                if (has_w) tma_load_2d(st + L.off_w, &m.w, c0, y, bar);
            };
            TmItem pf = first;   // next chunk to load
            TmItem cur = first;  // chunk whose completion is awaited next
            for (int s = 0; s < TM_STAGES; ++s) {
                if (item_valid(pf)) {
                    issue_raw(pf, s);
                    issue_ppo(pf, s);
                    item_next(pf);
                }
            }
            for (int j = 0; item_valid(cur); ++j) {
                const int sg = j % TM_STAGES;
                mbar_wait(&done[sg], (uint32_t)((j / TM_STAGES) & 1));
                if (j < 6) TM_TRACE(20 + 2 * j);
                // consumers are through chunk j => the scanner consumed raw chunk j long ago: both slots are free
                if (item_valid(pf)) issue_raw(pf, sg);
                if (GRADS) {
                    const int c0 = (int)(cur.tile * TM_TC), y = (int)(cur.hi - (long long)(cur.q + 1) * TM_R);
                    unsigned char* st = smem + sg * L.stage_bytes;
                    tma_store_2d(&m.gl, c0 * N, y, st);
                    tma_store_2d(&m.gv, c0, y, st + L.off_vn);
                    tma_store_commit();
                    tma_store_wait_read<0>();  // the stage may be overwritten once the store engine has read it
                }
                if (item_valid(pf)) {
                    issue_ppo(pf, sg);
                    item_next(pf);
                }
                if (j < 6) TM_TRACE(21 + 2 * j);
                item_next(cur);
            }
            tma_store_wait_all<0>();
        }
    } else if (wid == TM_CW + 1) {
        // =============================================== scanner ==========================================================
        TmItem it = first;
        float carry = 0.f;
        int slab = 0;
        for (int j = 0; item_valid(it); ++j) {
            const long long c0 = it.tile * TM_TC;
            const long long t0 = it.hi - (long long)(it.q + 1) * TM_R;  // time step of chunk row 0 (may be < 0)
            const int r0 = TM_SLAB - (it.q + 1) * TM_R;                  // slab-buffer row of chunk row 0
            float (*ab)[TM_TC] = advb[slab & 1];
            if (it.q == 0) {
                if (it.hi == T) carry = 0.f;
                if (slab >= 2) {  // the TMA store of slab-2 must have read this buffer
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                }
            }
            const int slot = j % 2;
            mbar_wait(&graw_full[slot], (uint32_t)((j / 2) & 1));
            const float* rv = reinterpret_cast<const float*>(raw + slot * TM_RAW_BYTES);
            const float* rn = rv + TM_R * TM_TC;
            const float* rr = rn + TM_R * TM_TC;
            const float* rd = rr + TM_R * TM_TC;
            const float* rt = rd + TM_R * TM_TC;
            // ---- delta / f: 32 lanes = 8 time steps x 4 column quads per pass ------------------------------------------
            const int cq = (lane & 3) * 4;
#pragma unroll
            for (int p = 0; p < TM_R / 8; ++p) {
                const int jj = p * 8 + (lane >> 2);
                const long long t = t0 + jj;
                if (t >= 0 && c0 + cq < B) {
                    const int o = jj * TM_TC + cq;
                    const float4 v4 = *reinterpret_cast<const float4*>(rv + o);
                    const float4 n4 = *reinterpret_cast<const float4*>(rn + o);
                    const float4 r4 = *reinterpret_cast<const float4*>(rr + o);
                    const float4 d4 = has_done ? *reinterpret_cast<const float4*>(rd + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 t4 = has_traj ? *reinterpret_cast<const float4*>(rt + o) : d4;
                    float vv[4] = {v4.x, v4.y, v4.z, v4.w}, nn[4] = {n4.x, n4.y, n4.z, n4.w};
                    float rw[4] = {r4.x, r4.y, r4.z, r4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
                    float tt[4] = {t4.x, t4.y, t4.z, t4.w};
                    float de[4], fa[4];
                    bool changed = false;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (has_done) {
                            changed |= (dd[k] != 0.f);
                            nn[k] = fmul(nn[k], fsub(1.f, dd[k]));
                        }
                        de[k] = fsub(fadd(rw[k], fmul(f.gamma, nn[k])), vv[k]);
                        fa[k] = fmul(f.gl, fsub(1.f, tt[k]));
                    }
                    *reinterpret_cast<float4*>(&ab[r0 + jj][cq]) = make_float4(de[0], de[1], de[2], de[3]);
                    *reinterpret_cast<float4*>(&fbuf[jj][cq]) = make_float4(fa[0], fa[1], fa[2], fa[3]);
                    if (changed && f.mask_inplace)
                        *reinterpret_cast<float4*>(f.next_value + t * B + c0 + cq) =
                            make_float4(nn[0], nn[1], nn[2], nn[3]);
                }
            }
            __syncwarp();
            // ---- sequential scan, lane = column, newest time step first ---------------------------------------------------
            if (lane < TM_TC && c0 + lane < B) {
                if (t0 >= 0) {
                    float d[TM_R], g[TM_R];
#pragma unroll
                    for (int k = 0; k < TM_R; ++k) {
                        d[k] = ab[r0 + TM_R - 1 - k][lane];
                        g[k] = fbuf[TM_R - 1 - k][lane];
                    }
#pragma unroll
                    for (int k = 0; k < TM_R; ++k) {
                        carry = fadd(d[k], fmul(g[k], carry));
                        ab[r0 + TM_R - 1 - k][lane] = carry;
                    }
                } else {
                    for (int jj = TM_R - 1; jj >= 0 && t0 + jj >= 0; --jj) {
                        carry = fadd(ab[r0 + jj][lane], fmul(fbuf[jj][lane], carry));
                        ab[r0 + jj][lane] = carry;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&adv_ready[(slab & 1) * 4 + it.q]);
                if (j == 0) TM_TRACE(19);
            }
            if (last_in_slab(it)) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&m.adv, (int)c0, (int)(it.hi - TM_SLAB), &ab[0][0]);
                    tma_store_commit();
                }
                ++slab;
            }
            item_next(it);
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else {
        // =============================================== consumers ========================================================
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
        TmItem it = first;
        int slab = 0;
        for (int j = 0; item_valid(it); ++j) {
            const int sg = j % TM_STAGES;
            const long long c0 = it.tile * TM_TC;
            const long long t0 = it.hi - (long long)(it.q + 1) * TM_R;
            const int r0 = TM_SLAB - (it.q + 1) * TM_R;
            const int W = (int)((B - c0) < TM_TC ? (B - c0) : TM_TC);
            unsigned char* st = smem + sg * L.stage_bytes;
            mbar_wait(&full[sg], (uint32_t)((j / TM_STAGES) & 1));
            if (tid == 0 && j < 6) TM_TRACE(1 + 3 * j);
            mbar_wait(&adv_ready[(slab & 1) * 4 + it.q], (uint32_t)((slab >> 1) & 1));
            if (tid == 0 && j < 6) TM_TRACE(2 + 3 * j);
            const float (*ab)[TM_TC] = advb[slab & 1];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = h * TM_CT + tid;
                const int jj = idx / TM_TC, c = idx % TM_TC;
                if (t0 + jj >= 0 && c < W) {
                    float* grow = GRADS ? reinterpret_cast<float*>(st) + idx * N : nullptr;
                    float* gval = GRADS ? reinterpret_cast<float*>(st + L.off_vn) + idx : nullptr;
                    ppo_row_compute_to<NC, true, GRADS>(a, L, st, idx, N, ab[r0 + jj][c], grow, gval, up, acc);
                }
            }
            if (GRADS) fence_proxy_async_smem();  // gradient rows -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[sg]);
            if (tid == 0 && j < 6) TM_TRACE(3 + 3 * j);
            if (last_in_slab(it)) ++slab;
            item_next(it);
        }
    }
    grid_store_partials<6, TM_THREADS>(acc, ws);  // summed by finalize_sums_kernel
}

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}

// [rows][inner] row-major tensor of `esz`-byte elements, box = box_rows x box_inner
static bool make_map(CUtensorMap* out, CUtensorMapDataType dt, int esz, const void* base, long long rows,
                     long long inner, int box_rows, int box_inner) {
    if (!base) {
        memset(out, 0, sizeof(*out));
        return true;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)inner * esz};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return encode_fn()(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool coltma_ok(const FusedArgs& f) {
    static int off = -1;
    if (off < 0) {
        const char* e = getenv("B200RL_COL_TMA");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    if (off || !coltile_ok(f) || !encode_fn()) return false;
    const PpoArgs& a = f.p;
    // ragged T would need negative box coordinates, which the TMA unit rejects (cudaErrorIllegalInstruction on B200)
    if ((f.T % TM_SLAB) != 0) return false;
    if (a.N * TM_TC > 256 || f.B < TM_TC || f.T > 0x7fffff00LL || f.B * a.N > 0x7fffff00LL) return false;
    if ((f.B * 4) % 16 != 0) return false;
    const int stage = (TM_ITEMS * (2 * a.N * 4 + 8 + 12 + (a.weight ? 4 : 0) + (a.logit_pre ? a.N * 4 : 0)) + 127) & ~127;
    const size_t smem = (size_t)TM_STAGES * stage + 2 * TM_RAW_BYTES + 2 * TM_SLAB * TM_TC * 4 + TM_R * TM_TC * 4 + 256;
    return smem <= 227 * 1024;
}

template <int NC, bool GRADS>
static int launch_tma(const FusedArgs& f, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    const PpoArgs& a = f.p;
    const int N = a.N;
    const int stage = (TM_ITEMS * (2 * N * 4 + 8 + 12 + (a.weight ? 4 : 0) + (a.logit_pre ? N * 4 : 0)) + 127) & ~127;
    const size_t smem = (size_t)TM_STAGES * stage + 2 * TM_RAW_BYTES + 2 * TM_SLAB * TM_TC * 4 + TM_R * TM_TC * 4 + 256;
    auto kern = gae_ppo_tma_kernel<NC, GRADS>;
    static int sm_count = 0;
    static size_t smem_set = 0;
    cudaError_t e;
    if (sm_count == 0) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    }
    if (smem > smem_set) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return (int)e;
        smem_set = smem;
    }
    static size_t occ_smem = (size_t)-1;
    static int per_sm = 0;
    if (occ_smem != smem) {
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TM_THREADS, smem)) != cudaSuccess)
            return (int)e;
        occ_smem = smem;
    }
    if (per_sm < 1) return B200RL_ERR_ARG;
    const long long n_tiles = (f.B + TM_TC - 1) / TM_TC;
    long long grid = (long long)sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (ws_bytes < WS_MIN_BYTES || !ws_partials_fit((long long)(grid * 6), ws_bytes))
        return B200RL_ERR_WORKSPACE;
    ColMaps m;
    const long long T = f.T, B = f.B;
    const CUtensorMapDataType F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    bool ok = make_map(&m.ln, F32, 4, a.logit_new, T, B * N, TM_R, TM_TC * N) &&
              make_map(&m.lo, F32, 4, a.logit_old, T, B * N, TM_R, TM_TC * N) &&
              make_map(&m.lp, F32, 4, a.logit_pre, T, B * N, TM_R, TM_TC * N) &&
              make_map(&m.act, CU_TENSOR_MAP_DATA_TYPE_INT64, 8, a.action, T, B, TM_R, TM_TC) &&
              make_map(&m.vn, F32, 4, a.value_new, T, B, TM_R, TM_TC) &&
              make_map(&m.vo, F32, 4, a.value_old, T, B, TM_R, TM_TC) &&
              make_map(&m.ret, F32, 4, a.ret, T, B, TM_R, TM_TC) && make_map(&m.w, F32, 4, a.weight, T, B, TM_R, TM_TC) &&
              make_map(&m.gl, F32, 4, GRADS ? a.grad_logit : nullptr, T, B * N, TM_R, TM_TC * N) &&
              make_map(&m.gv, F32, 4, GRADS ? a.grad_value : nullptr, T, B, TM_R, TM_TC) &&
              make_map(&m.value, F32, 4, f.value, T, B, TM_R, TM_TC) &&
              make_map(&m.nv, F32, 4, f.next_value, T, B, TM_R, TM_TC) &&
              make_map(&m.reward, F32, 4, f.reward, T, B, TM_R, TM_TC) &&
              make_map(&m.done, F32, 4, f.done, T, B, TM_R, TM_TC) &&
              make_map(&m.traj, F32, 4, f.traj, T, B, TM_R, TM_TC) &&
              make_map(&m.adv, F32, 4, a.adv_out, T, B, TM_SLAB, TM_TC);
    if (!ok) return B200RL_ERR_ARG;
    (void)launch_k(kern, (int)grid, TM_THREADS, smem, st, m, f, ws);
    FinalizeArgs fa{};
    const double is = 1.0 / (double)a.S;
    fa.scale[0] = is; fa.scale[1] = 0.5 * is; fa.scale[2] = is; fa.scale[3] = a.logit_pre ? is : 0.0;
    fa.scale[4] = is; fa.scale[5] = is;
    fa.k = 6; fa.n_blocks = (int)grid;
    (void)launch_finalize(ws, out, fa, st);
    return (int)cudaGetLastError();
}

template <bool GRADS>
static int dispatch_tma(const FusedArgs& f, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    switch (f.p.N) {
#define B200RL_CASE(n) \
    case n:            \
        return launch_tma<n, GRADS>(f, out, ws, ws_bytes, st);
        B200RL_CASE(2) B200RL_CASE(3) B200RL_CASE(4) B200RL_CASE(5) B200RL_CASE(6) B200RL_CASE(7) B200RL_CASE(8)
        B200RL_CASE(9) B200RL_CASE(10) B200RL_CASE(12) B200RL_CASE(14) B200RL_CASE(16)
#undef B200RL_CASE
        default:
            return launch_tma<0, GRADS>(f, out, ws, ws_bytes, st);
    }
}

int launch_coltma(const FusedArgs& f, bool grads, float* out, float* ws, size_t ws_bytes, cudaStream_t st) {
    return grads ? dispatch_tma<true>(f, out, ws, ws_bytes, st) : dispatch_tma<false>(f, out, ws, ws_bytes, st);
}

}  // namespace b200rl
