// K2: batched QUBO energies E[q,r] = x^T Q_q x in float64.
//
// Replaces neal's get_state_energy + dimod's offset, i.e. the BINARY energy the reference receives
// in SampleSet.record.energy.  With x in {0,1} this is Y = X Q
// followed by a masked row sum, so it is laid out as a register-tiled FP64 GEMM: a CTA owns 64
// reads, walks the (j-tile, i-tile) grid of Q through shared memory, skips tiles of Q that are
// entirely zero (the reference's Q is upper-triangular) and reduces in a fixed order, so the result
// is deterministic.
#include "common.cuh"

namespace {

constexpr int RT = 64;   // reads per CTA
constexpr int TJ = 64;   // columns of Q per tile
constexpr int TI = 32;   // rows of Q per tile

__global__ void __launch_bounds__(256) qubo_energy_kernel(const double *__restrict__ Q, int n,
                                                          const int8_t *__restrict__ states, long long R,
                                                          double *__restrict__ energy)
{
    __shared__ double Qs[TI][TJ];        // 16 KB
    __shared__ double Xs[TI][RT + 2];    // x[r][i] transposed, as 0.0 / 1.0

    const size_t q = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * RT;
    const double *Qq = Q + q * (size_t)n * (size_t)n;
    const int8_t *Sq = states + q * (size_t)R * (size_t)n;
    const int tid = threadIdx.x;
    const int tx = tid & 15;             // column group: columns tx*4 .. tx*4+3 of the j tile
    const int ty = tid >> 4;             // read group:   reads   ty*4 .. ty*4+3

    double e[4] = {0.0, 0.0, 0.0, 0.0};

    for (int j0 = 0; j0 < n; j0 += TJ) {
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

        for (int i0 = 0; i0 < n; i0 += TI) {
            // stage Q[i0:i0+TI, j0:j0+TJ]
            int nonzero = 0;
            for (int idx = tid; idx < TI * TJ; idx += 256) {
                const int ii = idx / TJ, jj = idx % TJ;
                const int gi = i0 + ii, gj = j0 + jj;
                const double v = (gi < n && gj < n) ? Qq[(size_t)gi * n + gj] : 0.0;
                Qs[ii][jj] = v;
                nonzero |= (v != 0.0);
            }
            const int any = __syncthreads_or(nonzero);
            if (!any) continue;           // uniform: the whole tile is zero
            // stage x[r0:r0+RT, i0:i0+TI] transposed
            for (int idx = tid; idx < TI * RT; idx += 256) {
                const int rr = idx / TI, ii = idx % TI;
                const long long gr = r0 + rr;
                const int gi = i0 + ii;
                const int8_t b = (gr < R && gi < n) ? Sq[(size_t)gr * n + gi] : (int8_t)0;
                Xs[ii][rr] = b ? 1.0 : 0.0;
            }
            __syncthreads();
#pragma unroll 4
            for (int ii = 0; ii < TI; ++ii) {
                double qv[4], xv[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) qv[b] = Qs[ii][tx * 4 + b];
#pragma unroll
                for (int a = 0; a < 4; ++a) xv[a] = Xs[ii][ty * 4 + a];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(xv[a], qv[b], acc[a][b]);
            }
            __syncthreads();
        }
        // e[r] += sum_j x[r][j] * Y[r][j]
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const long long gr = r0 + ty * 4 + a;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int gj = j0 + tx * 4 + b;
                if (gr < R && gj < n && Sq[(size_t)gr * n + gj]) e[a] += acc[a][b];
            }
        }
    }
    // fixed-order reduction over the 16 column groups (tx = low 4 bits of the lane index)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double v = e[a];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const long long gr = r0 + ty * 4 + a;
        if (tx == 0 && gr < R) energy[q * (size_t)R + gr] = v;
    }
}

}  // namespace

extern "C" QBM_API int qbm_qubo_energy(const double *Q, int n, long long batch_q, const int8_t *states, long long R,
                               double *energy_out, void *stream)
{
    QBM_CHECK_ARG(Q && states && energy_out, "qbm_qubo_energy: null pointer argument");
    QBM_CHECK_ARG(n >= 1 && batch_q >= 1 && R >= 1, "qbm_qubo_energy: n, batch_q and R must be >= 1");
    QBM_CHECK_ARG(batch_q <= 65535, "qbm_qubo_energy: batch_q > 65535 not supported in one call");
    const long long tiles = (R + RT - 1) / RT;
    QBM_CHECK_ARG(tiles <= 0x7fffffffLL, "qbm_qubo_energy: too many reads");
    qubo_energy_kernel<<<dim3((unsigned)tiles, (unsigned)batch_q), 256, 0, (cudaStream_t)stream>>>(
        Q, n, states, R, energy_out);
    QBM_LAUNCH_OK("qubo_energy_kernel");
    return QBM_OK;
}
