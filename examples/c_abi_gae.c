/* A plain-C caller of the C ABI (include/b200rl.h): no torch, no Python -- cudaMalloc'd buffers and a cudaStream_t.
 *
 *     gcc -std=c11 -O2 -ffp-contract=off -I include -I /usr/local/cuda/include examples/c_abi_gae.c \
 *         -L di-engine_b200/lib -lb200rl -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/di-engine_b200/lib -o c_abi_gae
 *
 * Runs b200rl_gae (ding/rl_utils/gae.py:25-70) on a (T, B) batch and checks the result bit for bit against the same
 * recurrence on the host (separate multiply and add, like the reference's torch ops; hence -ffp-contract=off), including the
 * in-place next_value mask.  This is what a reference-side binding does underneath.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200rl.h"

#define CHECK(x)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

static float frand(unsigned* s) {
    *s = *s * 1664525u + 1013904223u;
    return (float)((*s >> 8) & 0xffff) / 32768.0f - 1.0f;
}

int main(void) {
    const long long T = 128, B = 1024, n = T * B;
    const double gamma = 0.99, lambda_ = 0.95;
    const size_t bytes = (size_t)n * sizeof(float);
    float *v = malloc(bytes), *nv = malloc(bytes), *r = malloc(bytes), *d = malloc(bytes), *tf = malloc(bytes);
    float *adv = malloc(bytes), *nv_out = malloc(bytes), *ref = malloc(bytes);
    unsigned seed = 12345u;
    for (long long i = 0; i < n; ++i) {
        v[i] = frand(&seed);
        nv[i] = frand(&seed);
        r[i] = frand(&seed);
        d[i] = (frand(&seed) > 0.97f) ? 1.0f : 0.0f;
        tf[i] = (i / B == T - 1) ? 1.0f : d[i];
    }
    /* host recurrence in the reference's operation order: nv *= 1-done; delta = r + g*nv - v; f = (g*l)*(1-tf) */
    const float g = (float)gamma, gl = (float)(gamma * lambda_);
    for (long long c = 0; c < B; ++c) {
        float carry = 0.0f;
        for (long long t = T - 1; t >= 0; --t) {
            const long long i = t * B + c;
            const float m = nv[i] * (1.0f - d[i]);
            const float delta = (r[i] + g * m) - v[i];
            const float f = gl * (1.0f - tf[i]);
            carry = delta + f * carry;
            ref[i] = carry;
        }
    }
    float *dv, *dnv, *dr, *dd, *dtf, *dadv;
    CHECK(cudaMalloc((void**)&dv, bytes)); CHECK(cudaMalloc((void**)&dnv, bytes)); CHECK(cudaMalloc((void**)&dr, bytes));
    CHECK(cudaMalloc((void**)&dd, bytes)); CHECK(cudaMalloc((void**)&dtf, bytes)); CHECK(cudaMalloc((void**)&dadv, bytes));
    CHECK(cudaMemcpy(dv, v, bytes, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dnv, nv, bytes, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dr, r, bytes, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dd, d, bytes, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dtf, tf, bytes, cudaMemcpyHostToDevice));
    cudaStream_t st;
    CHECK(cudaStreamCreate(&st));
    const int rc = b200rl_gae(dv, dnv, dr, dd, dtf, dadv, T, B, 1, gamma, lambda_, 1, (void*)st);
    if (rc != 0) {
        fprintf(stderr, "b200rl_gae returned %d\n", rc);
        return 3;
    }
    CHECK(cudaStreamSynchronize(st));
    CHECK(cudaMemcpy(adv, dadv, bytes, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(nv_out, dnv, bytes, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        if (memcmp(&adv[i], &ref[i], sizeof(float)) != 0) ++bad;
        const float m = nv[i] * (1.0f - d[i]);
        if (memcmp(&nv_out[i], &m, sizeof(float)) != 0) ++bad;
    }
    printf("b200rl version %d (sm_%d): gae on %lld x %lld via the C ABI, %lld mismatching values\n", b200rl_version(),
           b200rl_built_for_sm(), T, B, bad);
    return bad == 0 ? 0 : 1;
}
