#pragma once
// Arguments shared by the two implementations of the one-launch learner step gae -> ppo_error:
// fused.cu (row tiles + cross-CTA chunk counters) and coltile.cu (column tiles, no cross-CTA dependency).
#include "ppo_math.cuh"

namespace b200rl {

struct FusedArgs {
    PpoArgs p;  // p.adv = the (T*B) advantage buffer this kernel WRITES (phase G) and reads (phase P)
    const float* value;
    float* next_value;
    const float* reward;
    const float* done;
    const float* traj;
    long long T, B;
    float gamma, gl;
    int mask_inplace;
    int trace;
    int wait_ns;  // colws.cu consumers: suspend-time hint of the mbarrier waits in ns (0 = none; B200RL_COL_WAIT_NS)
    int loader;  // colws.cu loader warp: 0 = cheap copies for the first stage only (default), 1 = flat loop, 2 | 3 = cheap copies everywhere
    // optional data-parallel exchange of the six loss scalars, fused into the step's finalize launch (colws.cu; common.cuh)
    const unsigned long long* x_mailboxes;
    unsigned int* x_seq;
    float* x_out_mean;
    int x_rank, x_world;
};

// column-tile implementations.  coltile.cu: all threads copy (cp.async) and compute; it also hosts the dispatcher:
// variant 0 = best available (colws > coltile), 1 = coltile.cu only, 2 = coltma.cu if it supports the call, 3 = colws.cu
bool coltile_ok(const FusedArgs& f);
int launch_coltile(const FusedArgs& f, bool grads, int variant, float* out, float* ws, size_t ws_bytes,
                   cudaStream_t st);
// colws.cu (warp-specialised: loader / scanner / consumer warps, mbarrier pipeline, cp.async copies)
bool colws_ok(const FusedArgs& f);
int launch_colws(const FusedArgs& f, bool grads, float* out, float* ws, size_t ws_bytes, cudaStream_t st);
// coltma.cu (2-D tensor-map TMA copies)
bool coltma_ok(const FusedArgs& f);
int launch_coltma(const FusedArgs& f, bool grads, float* out, float* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace b200rl
