// K0: dense QUBO -> spin model on the device (batched).
//
// Replaces dimod.BQM(Q, "BINARY") + change_vartype(SPIN) + the two reductions neal's legacy
// _default_ising_beta_range needs (call sites
// src/qubo/sampler.py:7-8,31 and src/model/faster_dqbm.py:577,619).  One CTA per problem, one warp
// per row; all sums are float64 and reduced in a fixed order.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int WARPS = 8;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(WARPS * 32) qubo_to_ising_kernel(const double *__restrict__ Q, int n,
                                                                   float *__restrict__ J, float *__restrict__ h,
                                                                   double *__restrict__ offset, double *__restrict__ range)
{
    const size_t q = blockIdx.x;
    const double *Qq = Q + q * (size_t)n * (size_t)n;
    float *Jq = J + q * (size_t)n * (size_t)n;
    float *hq = h + q * (size_t)n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    double w_off = 0.0, w_max = 0.0, w_min = INFINITY;   // per-warp running values (uniform across lanes)
    for (int i = warp; i < n; i += WARPS) {
        double s = 0.0, sa = 0.0, mn = INFINITY;
        for (int j = lane; j < n; j += 32) {
            double b = 0.0;
            if (j != i) b = Qq[(size_t)i * n + j] + Qq[(size_t)j * n + i];
            const double jv = b * 0.25;
            Jq[(size_t)i * n + j] = (float)jv;
            s += jv;
            const double aj = fabs(jv);
            sa += aj;
            if (aj != 0.0) mn = fmin(mn, aj);
        }
        s = warp_sum(s); sa = warp_sum(sa); mn = warp_min(mn);
        const double a = Qq[(size_t)i * n + i];
        const double hi = a * 0.5 + s;
        if (lane == 0) hq[i] = (float)hi;
        const double ah = fabs(hi);
        if (ah != 0.0) mn = fmin(mn, ah);
        w_min = fmin(w_min, mn);
        w_max = fmax(w_max, ah + sa);
        w_off += a * 0.5 + 0.5 * s;                 // each coupler is seen from both of its rows
    }
    __shared__ double sh[3][WARPS];
    if (lane == 0) { sh[0][warp] = w_off; sh[1][warp] = w_max; sh[2][warp] = w_min; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double o = 0.0, mx = 0.0, mn = INFINITY;
        for (int w = 0; w < WARPS; ++w) { o += sh[0][w]; mx = fmax(mx, sh[1][w]); mn = fmin(mn, sh[2][w]); }
        if (offset != nullptr) offset[q] = o;
        if (range != nullptr) {
            range[2 * q + 0] = isinf(mn) ? 0.0 : mn;
            range[2 * q + 1] = mx;
        }
    }
}

// neal's legacy default beta range (sampler.py:281) from K0's two reductions, then
// np.geomspace(hot, cold, num_betas) = 10 ** linspace(log10 hot, log10 cold) with pinned end points, cast to fp32
__global__ void beta_schedule_kernel(const double *__restrict__ range, const long long batch, const int num_betas,
                                     float *__restrict__ betas)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * num_betas) return;
    const long long q = e / num_betas;
    const int i = (int)(e % num_betas);
    const double mn = range[2 * q], mx = range[2 * q + 1];
    const bool empty = mx == 0.0;                                  // no bias at all: neal falls back to [0.1, 1.0]
    const double hot = empty ? 0.1 : 0.6931471805599453 / mx;      // np.log(2)
    const double cold = empty ? 1.0 : 4.605170185988092 / mn;      // np.log(100)
    double v;
    if (i == 0) v = hot;
    else if (i == num_betas - 1) v = cold;
    else {
        const double ls = log10(hot), le = log10(cold);
        const double step = (le - ls) / (double)(num_betas - 1);
        v = pow(10.0, (double)i * step + ls);
    }
    betas[e] = (float)v;
}

}  // namespace

extern "C" QBM_API int qbm_beta_schedule(const double *range, long long batch, int num_betas, float *betas_out, void *stream)
{
    QBM_CHECK_ARG(range && betas_out, "qbm_beta_schedule: null pointer argument");
    QBM_CHECK_ARG(batch >= 1 && num_betas >= 0, "qbm_beta_schedule: bad sizes");
    if (num_betas == 0) return QBM_OK;
    const long long total = batch * num_betas;
    beta_schedule_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(range, batch, num_betas, betas_out);
    QBM_LAUNCH_OK("beta_schedule_kernel");
    return QBM_OK;
}

extern "C" QBM_API int qbm_qubo_to_ising(const double *Q, int n, long long batch, float *J_out, float *h_out,
                                 double *offset, double *range, void *stream)
{
    QBM_CHECK_ARG(Q && J_out && h_out, "qbm_qubo_to_ising: null pointer argument");
    QBM_CHECK_ARG(n >= 1 && batch >= 1, "qbm_qubo_to_ising: n and batch must be >= 1");
    QBM_CHECK_ARG(batch <= 0x7fffffffLL, "qbm_qubo_to_ising: batch too large");
    qubo_to_ising_kernel<<<(unsigned)batch, WARPS * 32, 0, (cudaStream_t)stream>>>(Q, n, J_out, h_out, offset, range);
    QBM_LAUNCH_OK("qubo_to_ising_kernel");
    return QBM_OK;
}
