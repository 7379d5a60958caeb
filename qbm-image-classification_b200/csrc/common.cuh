// Shared device/host helpers of libqbm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/qbm_b200.h"

// ---- error plumbing -------------------------------------------------------------------------
void qbm_set_error(const char *fmt, ...);   // capi.cu

#define QBM_CHECK_ARG(cond, ...)                          \
    do {                                                  \
        if (!(cond)) {                                    \
            qbm_set_error(__VA_ARGS__);                   \
            return QBM_EINVAL;                            \
        }                                                 \
    } while (0)

#define QBM_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            qbm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return QBM_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

#define QBM_LAUNCH_OK(name)                                                                 \
    do {                                                                                    \
        cudaError_t err__ = cudaGetLastError();                                             \
        if (err__ != cudaSuccess) {                                                         \
            qbm_set_error("launch of %s failed: %s", name, cudaGetErrorString(err__));      \
            return QBM_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

// ---- Philox4x32-10 (counter-based; key = seed, counter = (chain, sweep, block)) --------------
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// ---- exp_spec: FMA-only exp on (-87, 0], identical operation sequence to the replay oracle ---
__device__ __forceinline__ float exp_spec(float x)
{
    const float t = __fmul_rn(x, 1.44269504f);
    const float k = rintf(t);
    float f = __fmaf_rn(k, -0.693145751953125f, x);
    f = __fmaf_rn(k, -1.42860677e-06f, f);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, f, 1.3981999507e-3f);
    p = __fmaf_rn(p, f, 8.3334519073e-3f);
    p = __fmaf_rn(p, f, 4.1665795894e-2f);
    p = __fmaf_rn(p, f, 1.6666665459e-1f);
    p = __fmaf_rn(p, f, 5.0000001201e-1f);
    const float f2 = __fmul_rn(f, f);
    float r = __fmaf_rn(p, f2, f);
    r = __fadd_rn(r, 1.0f);
    return __int_as_float(__float_as_int(r) + (__float2int_rn(k) << 23));
}

// Variable -> storage position inside a 128-variable window: the 4 variables a lane owns
// (v = base + k*32 + lane, k = 0..3) sit in 4 consecutive floats so one 128-bit load fetches them.
__host__ __device__ __forceinline__ int p128_pos(int v)
{
    return (v & ~127) | ((v & 31) << 2) | ((v >> 5) & 3);
}
