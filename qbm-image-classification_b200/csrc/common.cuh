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

// ---- neg_log_u32: FMA-only  -ln(u / 2^32)  for a 32-bit uniform, identical operation sequence to the
// replay oracle (oracle/replay_sa.c).  The Metropolis test  u/2^32 < exp(-beta dE)  is evaluated in
// the log domain,  dE < -ln(u/2^32) / beta,  so the transcendental is paid once per (chain, sweep,
// variable) instead of once per re-evaluation of a proposal.  u = 0 -> +inf (bound falls back to the
// threshold 44.36142/beta).
__device__ __forceinline__ float neg_log_u32(uint32_t u)
{
    if (u == 0u) return __int_as_float(0x7f800000);
    const float x = __uint2float_rn(u);                         // [1, 2^32], round to nearest even
    const int bits = __float_as_int(x);
    int e = (bits >> 23) - 127;
    float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);  // [1, 2)
    if (m > 1.41421354f) { m = __fmul_rn(m, 0.5f); e += 1; }
    const float f = __fadd_rn(m, -1.0f);                         // exact
    const float z = __fmul_rn(f, f);
    float y = 7.0376836292e-2f;
    y = __fmaf_rn(y, f, -1.1514610310e-1f);
    y = __fmaf_rn(y, f, 1.1676998740e-1f);
    y = __fmaf_rn(y, f, -1.2420140846e-1f);
    y = __fmaf_rn(y, f, 1.4249322787e-1f);
    y = __fmaf_rn(y, f, -1.6668057665e-1f);
    y = __fmaf_rn(y, f, 2.0000714765e-1f);
    y = __fmaf_rn(y, f, -2.4999993993e-1f);
    y = __fmaf_rn(y, f, 3.3333331174e-1f);
    y = __fmul_rn(y, f);
    y = __fmul_rn(y, z);
    y = __fmaf_rn(-0.5f, z, y);
    const float r = __fadd_rn(f, y);                             // ln(m)
    const float E = (float)(32 - e);
    float nl = __fmaf_rn(E, 0.693359375f, -r);
    nl = __fmaf_rn(E, -2.12194440e-4f, nl);
    return nl;
}

// Variable -> storage position inside a 128-variable window: the 4 variables a lane owns
// (v = base + k*32 + lane, k = 0..3) sit in 4 consecutive floats so one 128-bit load fetches them.
__host__ __device__ __forceinline__ int p128_pos(int v)
{
    return (v & ~127) | ((v & 31) << 2) | ((v >> 5) & 3);
}
