// K0: dense QUBO -> spin model on the device (batched).
//
// Replaces dimod.BQM(Q, "BINARY") + change_vartype(SPIN) + the two reductions neal's legacy
// _default_ising_beta_range needs (call sites
// src/qubo/sampler.py:7-8,31 and src/model/faster_dqbm.py:577,619).  One CTA per problem, one warp
// per row; all sums are float64 in numpy's pairwise order, so h, J and the beta-range reductions equal
// the host path (qbm_b200/ising.py) bit for bit.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int WARPS = 16;
constexpr int MAX_LEAVES = 64;      // leaves of numpy's pairwise-summation tree over one row (n <= 4096: <= 64)

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// numpy's float64 pairwise summation of n contiguous values (numpy/core/src/umath/loops_utils.h.src):
//   n < 8: sequential from 0;  n <= 128: 8 running sums r[j] += a[i + j], combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//   then the n % 8 tail sequentially;  n > 128: split at n2 = n/2 - (n/2) % 8 and add the two halves.
// The row sums of the host path (qbm_b200/ising.py, i.e. dimod's BINARY->SPIN arithmetic as the oracle restates it) are
// reproduced bit for bit by giving every leaf (<= 128 values) of that tree to one lane and combining the leaf sums along
// the tree.  build_leaves() enumerates the leaves in order; combine() replays the additions of the inner nodes.
__device__ int build_leaves(int n, int *start, int *len)
{
    int cnt = 0;
    int st_s[16], st_l[16], sp = 0;
    st_s[0] = 0; st_l[0] = n; sp = 1;
    while (sp > 0) {
        --sp;
        const int s0 = st_s[sp], l0 = st_l[sp];
        if (l0 <= 128) { start[cnt] = s0; len[cnt] = l0; ++cnt; continue; }
        int n2 = l0 / 2; n2 -= n2 % 8;
        st_s[sp] = s0 + n2; st_l[sp] = l0 - n2; ++sp;       // right child is popped after the left one
        st_s[sp] = s0; st_l[sp] = n2; ++sp;
    }
    return cnt;
}
__device__ double combine(int n, const double *leaf, int &next)
{
    if (n <= 128) return leaf[next++];
    int n2 = n / 2; n2 -= n2 % 8;
    const double l = combine(n2, leaf, next);
    const double r = combine(n - n2, leaf, next);
    return __dadd_rn(l, r);
}

__global__ void __launch_bounds__(WARPS * 32) qubo_to_ising_kernel(const double *__restrict__ Q, int n,
                                                                   float *__restrict__ J, float *__restrict__ h,
                                                                   double *__restrict__ offset, double *__restrict__ range)
{
    __shared__ int lstart[MAX_LEAVES], llen[MAX_LEAVES], nleaf_s;
    __shared__ double leafsum[WARPS][2][MAX_LEAVES];
    __shared__ double sh[3][WARPS];
    const size_t q = blockIdx.x;
    const double *Qq = Q + q * (size_t)n * (size_t)n;
    float *Jq = J + q * (size_t)n * (size_t)n;
    float *hq = h + q * (size_t)n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) nleaf_s = build_leaves(n, lstart, llen);
    __syncthreads();
    const int nleaf = nleaf_s;

    double w_off = 0.0, w_max = 0.0, w_min = INFINITY;   // per-warp running values (kept by lane 0)
    for (int i = warp; i < n; i += WARPS) {
        double mn = INFINITY;
        // b_ij = Q_ij + Q_ji (0 on the diagonal) and |b_ij / 4|, summed leaf by leaf: 8 lanes per leaf, lane e of a group
        // is numpy's running sum r[e]; four leaves at a time
        const int grp = lane >> 3, e = lane & 7;
        auto elem = [&](int j, double &b, double &a) {
            b = (j != i) ? __dadd_rn(Qq[(size_t)i * n + j], Qq[(size_t)j * n + i]) : 0.0;
            const double jv = b * 0.25;
            Jq[(size_t)i * n + j] = (float)jv;
            a = fabs(jv);
            if (a != 0.0) mn = fmin(mn, a);
        };
        for (int L0 = 0; L0 < nleaf; L0 += 4) {
            const int L = L0 + grp;
            const bool valid = L < nleaf;
            const int s0 = valid ? lstart[L] : 0, l0 = valid ? llen[L] : 0;
            const int body = l0 - (l0 % 8);
            double rb = 0.0, ra = 0.0;
            if (valid && body > 0) elem(s0 + e, rb, ra);                     // r[e] = a[e]
#pragma unroll 1
            for (int m = 1; m < 16; ++m) {
                const int idx = 8 * m + e;
                if (valid && idx < body) { double b, a; elem(s0 + idx, b, a); rb = __dadd_rn(rb, b); ra = __dadd_rn(ra, a); }
            }
            // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) as an xor butterfly inside the group of 8 lanes
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                rb = __dadd_rn(rb, __shfl_xor_sync(0xffffffffu, rb, o));
                ra = __dadd_rn(ra, __shfl_xor_sync(0xffffffffu, ra, o));
            }
            if (valid && e == 0) {
                double sb = body > 0 ? rb : 0.0, sa = body > 0 ? ra : 0.0;
                for (int tl = body; tl < l0; ++tl) { double b, a; elem(s0 + tl, b, a); sb = __dadd_rn(sb, b); sa = __dadd_rn(sa, a); }
                leafsum[warp][0][L] = sb;
                leafsum[warp][1][L] = sa;
            }
        }
        mn = warp_min(mn);
        __syncwarp();
        if (lane == 0) {
            int nx = 0;
            const double sb = combine(n, leafsum[warp][0], nx);            // sum_j b_ij       (== Bm.sum(axis=2))
            nx = 0;
            const double sa = combine(n, leafsum[warp][1], nx);            // sum_j |J_ij|     (== np.abs(J).sum(axis=2))
            const double a = Qq[(size_t)i * n + i];
            const double hi = __dadd_rn(a / 2.0, sb / 4.0);                // a / 2 + Bm.sum(axis=2) / 4
            hq[i] = (float)hi;
            const double ah = fabs(hi);
            if (ah != 0.0) mn = fmin(mn, ah);
            w_min = fmin(w_min, mn);
            w_max = fmax(w_max, __dadd_rn(ah, sa));
            w_off += a * 0.5 + 0.125 * sb;            // each coupler is seen from both of its rows (information only)
        }
        __syncwarp();
    }
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
    if (n > 4096) {       // the pairwise-summation leaf table (MAX_LEAVES = 64 leaves of <= 128 elements) covers n <= 4096 only
        qbm_set_error("qbm_qubo_to_ising: n=%d exceeds the 4096 variables the row-sum kernel is built for", n);
        return QBM_EUNSUPPORTED;
    }
    qubo_to_ising_kernel<<<(unsigned)batch, WARPS * 32, 0, (cudaStream_t)stream>>>(Q, n, J_out, h_out, offset, range);
    QBM_LAUNCH_OK("qubo_to_ising_kernel");
    return QBM_OK;
}
