// K3: phase statistics -- first and second moments <s_i>, <s_i s_j> of a 0/1 sample set.
//
// Replaces the np.average / (block^T @ block) / n_reads reductions of
// src/train/train.py:135-253 (get_average_configuration_single) and
// src/model/discriminative_qbm.py:696-760 (get_average_configuration).  Samples are binary, so
// the contraction is done on bit-planes: reads are packed 32 per word per variable and
// count_ij = sum_words popc(a_i & a_j) -- exact integers, divided by R once (the reference divides
// float64 sums of 0/1 products by R, which is the same rational number).
#include "common.cuh"

namespace {

// bits[q][i][w] : bit r of word w = states[q][32*w + r][i]
__global__ void stats_pack_kernel(const int8_t *__restrict__ states, long long R, int n, int Rw,
                                  uint32_t *__restrict__ bits)
{
    const size_t q = blockIdx.z;
    const int w = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int8_t *S = states + q * (size_t)R * (size_t)n;
    const long long rbase = (long long)w * 32;
    uint32_t word = 0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
        const long long gr = rbase + r;
        if (gr < R && S[(size_t)gr * n + i]) word |= (1u << r);
    }
    bits[(q * (size_t)n + i) * (size_t)Rw + w] = word;
}

constexpr int TS = 32;    // variables per tile side
constexpr int WCH = 32;   // words staged per chunk

__global__ void __launch_bounds__(256) stats_pair_kernel(const uint32_t *__restrict__ bits, long long R, int n, int Rw,
                                                         double *__restrict__ mean_out, double *__restrict__ second_out)
{
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bj < bi) return;                          // symmetric: upper tiles only, mirrored on store
    if (second_out == nullptr && bi != bj) return; // means only: the diagonal tiles suffice
    const size_t q = blockIdx.z;
    __shared__ uint32_t A[TS][WCH + 1];
    __shared__ uint32_t B[TS][WCH + 1];
    const uint32_t *Bq = bits + q * (size_t)n * (size_t)Rw;
    const int tid = threadIdx.x;
    const int tj = tid & 15, ti = tid >> 4;       // thread owns pairs (2ti+{0,1}, 2tj+{0,1})
    uint32_t cnt[2][2] = {{0u, 0u}, {0u, 0u}};

    for (int w0 = 0; w0 < Rw; w0 += WCH) {
        for (int idx = tid; idx < TS * WCH; idx += 256) {
            const int v = idx / WCH, w = idx % WCH;
            const int gi = bi * TS + v, gj = bj * TS + v, gw = w0 + w;
            A[v][w] = (gi < n && gw < Rw) ? Bq[(size_t)gi * Rw + gw] : 0u;
            B[v][w] = (gj < n && gw < Rw) ? Bq[(size_t)gj * Rw + gw] : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int w = 0; w < WCH; ++w) {
            const uint32_t a0 = A[2 * ti][w], a1 = A[2 * ti + 1][w];
            const uint32_t b0 = B[2 * tj][w], b1 = B[2 * tj + 1][w];
            cnt[0][0] += __popc(a0 & b0);
            cnt[0][1] += __popc(a0 & b1);
            cnt[1][0] += __popc(a1 & b0);
            cnt[1][1] += __popc(a1 & b1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int gi = bi * TS + 2 * ti + a, gj = bj * TS + 2 * tj + b;
            if (gi < n && gj < n) {
                const double val = (double)cnt[a][b] / (double)R;   // the exact rational, rounded once
                if (second_out != nullptr) {
                    double *S2 = second_out + q * (size_t)n * (size_t)n;
                    S2[(size_t)gi * n + gj] = val;
                    S2[(size_t)gj * n + gi] = val;
                }
                if (gi == gj) mean_out[q * (size_t)n + gi] = val;
            }
        }
}

}  // namespace

extern "C" QBM_API size_t qbm_phase_stats_workspace_bytes(long long batch_q, long long R, int n)
{
    if (batch_q <= 0 || R <= 0 || n <= 0) return 0;
    const size_t Rw = (size_t)((R + 31) / 32);
    return (size_t)batch_q * (size_t)n * Rw * sizeof(uint32_t);
}

extern "C" QBM_API int qbm_phase_stats(const int8_t *states, long long batch_q, long long R, int n, double *mean_out,
                               double *second_out, void *workspace, size_t workspace_bytes, void *stream)
{
    QBM_CHECK_ARG(states && mean_out && workspace, "qbm_phase_stats: null pointer argument");
    QBM_CHECK_ARG(batch_q >= 1 && R >= 1 && n >= 1, "qbm_phase_stats: batch_q, R and n must be >= 1");
    QBM_CHECK_ARG(batch_q <= 65535, "qbm_phase_stats: batch_q > 65535 not supported in one call");
    if (workspace_bytes < qbm_phase_stats_workspace_bytes(batch_q, R, n)) {
        qbm_set_error("qbm_phase_stats: workspace of %zu bytes, need %zu", workspace_bytes,
                      qbm_phase_stats_workspace_bytes(batch_q, R, n));
        return QBM_EWORKSPACE;
    }
    const long long Rw = (R + 31) / 32;
    QBM_CHECK_ARG(Rw <= 65535, "qbm_phase_stats: more than 2097120 reads per problem not supported");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *bits = reinterpret_cast<uint32_t *>(workspace);
    stats_pack_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)Rw, (unsigned)batch_q), 128, 0, st>>>(
        states, R, n, (int)Rw, bits);
    QBM_LAUNCH_OK("stats_pack_kernel");
    const unsigned nt = (unsigned)((n + TS - 1) / TS);
    stats_pair_kernel<<<dim3(nt, nt, (unsigned)batch_q), 256, 0, st>>>(bits, R, n, (int)Rw, mean_out, second_out);
    QBM_LAUNCH_OK("stats_pair_kernel");
    return QBM_OK;
}
