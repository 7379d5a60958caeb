// K7 / K8 / K9: the parts of the Disc_QBM training step that surround the sampler, one launch each for a
// whole minibatch (float64 like the reference's numpy).
//
//   K7 disc_build_qubo   create_qubo_matrix_from        src/model/faster_dqbm.py:225-284 (== discriminative_qbm.py:240-300)
//   K8 disc_errors       get_average_configuration      src/model/discriminative_qbm.py:696-760   (stats_mode "loop")
//                        get_average_configuration_batch src/model/faster_dqbm.py:754-848          (stats_mode "faster",
//                        with its divergences) summed over the minibatch, clamped - unclamped,
//                        + the NLL of faster_dqbm.py:972-994
//   K9 sgd_apply         param -= lr * (err / batch)    faster_dqbm.py:1042-1059
//
// Parameter layout (one flat float64 buffer, also the layout of the error buffer):
//   [ b_h (h) | b_o (no) | W_vh ((no+di) x h) | W_vo (di x no) | W_oo (no x no) | W_hh (h x h, absent when restricted) ]
#include "common.cuh"

namespace {

struct DiscDims {
    int di, no, h, restricted;
    __host__ __device__ long long off_bh() const { return 0; }
    __host__ __device__ long long off_bo() const { return h; }
    __host__ __device__ long long off_vh() const { return (long long)h + no; }
    __host__ __device__ long long off_vo() const { return off_vh() + (long long)(no + di) * h; }
    __host__ __device__ long long off_oo() const { return off_vo() + (long long)di * no; }
    __host__ __device__ long long off_hh() const { return off_oo() + (long long)no * no; }
    __host__ __device__ long long total() const { return off_hh() + (restricted ? 0 : (long long)h * h); }
};

// one CTA per image
__global__ void __launch_bounds__(256) disc_build_qubo_kernel(const double *__restrict__ P, const DiscDims d,
                                                              const double *__restrict__ X, const double *__restrict__ Y,
                                                              const double beta_eff, double *__restrict__ Q)
{
    extern __shared__ double diag[];                      // [n]
    const size_t b = blockIdx.x;
    const int di = d.di, no = d.no, h = d.h;
    const double *bh = P + d.off_bh(), *bo = P + d.off_bo(), *Wvh = P + d.off_vh(), *Wvo = P + d.off_vo();
    const double *Woo = P + d.off_oo(), *Whh = d.restricted ? nullptr : P + d.off_hh();
    const double *x = X + b * (size_t)di;
    const bool clamped = Y != nullptr;
    const int n = clamped ? h : no + h;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double s = 0.0;
        if (clamped) {
            // [label, x] . W_vh : label rows FIRST (faster_dqbm.py:236)
            const double *y = Y + b * (size_t)no;
            for (int o = 0; o < no; ++o) s = fma(y[o], Wvh[(size_t)o * h + j], s);
            for (int i = 0; i < di; ++i) s = fma(x[i], Wvh[(size_t)(no + i) * h + j], s);
            diag[j] = bh[j] + s;
        } else if (j < no) {
            for (int i = 0; i < di; ++i) s = fma(x[i], Wvo[(size_t)i * no + j], s);
            diag[j] = bo[j] + s;
        } else {
            for (int i = 0; i < di; ++i) s = fma(x[i], Wvh[(size_t)(no + i) * h + (j - no)], s);
            diag[j] = bh[j - no] + s;
        }
    }
    __syncthreads();
    double *q = Q + b * (size_t)n * (size_t)n;
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = threadIdx.x >> 5; r < n; r += nwarps) {                         // one warp per row, lanes over columns
        double *qr = q + (size_t)r * n;
        for (int c = lane; c < n; c += 32) {
            double v = (r == c) ? diag[r] : 0.0;
            if (clamped) {
                if (Whh != nullptr) v += Whh[(size_t)r * h + c];
            } else {
                if (r < no && c >= no) v += Wvh[(size_t)r * h + (c - no)];      // output -> hidden couplings
                else if (r < no && c < no) v += Woo[(size_t)r * no + c];
                else if (r >= no && c >= no && Whh != nullptr) v += Whh[(size_t)(r - no) * h + (c - no)];
            }
            qr[c] = v / beta_eff;
        }
    }
}

// one thread per element of the error buffer; sums over the local images in image order
__global__ void disc_errors_kernel(const DiscDims d, const int faster, const long long B, const double *__restrict__ X,
                                   const double *__restrict__ Y, const double *__restrict__ Mc, const double *__restrict__ Sc,
                                   const double *__restrict__ Mu, const double *__restrict__ Su, double *__restrict__ E)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = d.total();
    if (e > total) return;
    const int di = d.di, no = d.no, h = d.h, nu = no + h;
    double s = 0.0;
    if (e == total) {
        // NLL of output node 0, float32 like torch.tensor(output_probs) (faster_dqbm.py:972-994); the loop variant
        // of the reference has its NLL commented out (total stays 0)
        if (faster) {
            for (long long b = 0; b < B; ++b) {
                const float p1 = (float)Mu[b * nu];
                const float v = (Y[b * no] != 0.0) ? p1 : (1.0f - p1);
                s -= (double)logf(v + 1e-12f);
            }
        }
    } else if (e < d.off_bo()) {
        const int j = (int)e;
        for (long long b = 0; b < B; ++b) s += Mc[b * h + j] - Mu[b * nu + no + j];
    } else if (e < d.off_vh()) {
        const int o = (int)(e - d.off_bo());
        for (long long b = 0; b < B; ++b) s += Y[b * no + o] - Mu[b * nu + o];
    } else if (e < d.off_vo()) {
        const long long r = e - d.off_vh();
        const int v = (int)(r / h), j = (int)(r % h);
        if (v < di) {       // statistics rows: x first, then the label (Appendix B Q1)
            for (long long b = 0; b < B; ++b) s = fma(X[b * di + v], Mc[b * h + j] - Mu[b * nu + no + j], s);
        } else {
            for (long long b = 0; b < B; ++b) s = fma(Y[b * no + (v - di)], Mc[b * h + j], s);
        }
    } else if (e < d.off_oo()) {
        const long long r = e - d.off_vo();
        const int v = (int)(r / no), o = (int)(r % no);
        for (long long b = 0; b < B; ++b) s = fma(X[b * di + v], Y[b * no + o] - Mu[b * nu + o], s);
    } else if (e < d.off_hh()) {
        const long long r = e - d.off_oo();
        const int o = (int)(r / no), o2 = (int)(r % no);
        if (o < o2) {
            for (long long b = 0; b < B; ++b) s += Y[b * no + o] * Y[b * no + o2] - Su[(b * nu + o) * nu + o2];
            if (faster && !d.restricted) s *= 2.0;       // faster_dqbm.py:831-845 adds the o-o term twice (Q2)
        }
    } else {
        const long long r = e - d.off_hh();
        const int i = (int)(r / h), j = (int)(r % h);
        if (i < j && !faster) {                          // faster_dqbm.py never accumulates <h h'> (Q2)
            for (long long b = 0; b < B; ++b) s += Sc[(b * h + i) * h + j] - Su[(b * nu + no + i) * nu + no + j];
        }
    }
    E[e] = s;
}

__global__ void sgd_apply_kernel(double *__restrict__ P, const double *__restrict__ E, const long long count, const double lr,
                                 const double batch)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) P[i] -= lr * (E[i] / batch);
}

int check_dims(const char *who, int di, int no, int h)
{
    if (di < 1 || no < 1 || h < 1) {
        qbm_set_error("%s: need dim_input, n_output, n_hidden >= 1 (got %d, %d, %d)", who, di, no, h);
        return QBM_EINVAL;
    }
    return QBM_OK;
}

}  // namespace

extern "C" QBM_API long long qbm_disc_param_count(int dim_input, int n_output, int n_hidden, int restricted)
{
    if (dim_input < 1 || n_output < 1 || n_hidden < 1) return 0;
    DiscDims d{dim_input, n_output, n_hidden, restricted ? 1 : 0};
    return d.total();
}

extern "C" QBM_API int qbm_disc_build_qubo(const double *params, int dim_input, int n_output, int n_hidden, int restricted,
                                           const double *X, const double *Y, long long B, double beta_eff, double *Q_out,
                                           void *stream)
{
    if (int rc = check_dims("qbm_disc_build_qubo", dim_input, n_output, n_hidden)) return rc;
    QBM_CHECK_ARG(params && X && Q_out, "qbm_disc_build_qubo: null pointer argument");
    QBM_CHECK_ARG(B >= 1 && B <= 0x7fffffffLL, "qbm_disc_build_qubo: bad batch size");
    QBM_CHECK_ARG(beta_eff != 0.0, "qbm_disc_build_qubo: beta_eff must not be 0");
    DiscDims d{dim_input, n_output, n_hidden, restricted ? 1 : 0};
    const int n = Y != nullptr ? n_hidden : n_output + n_hidden;
    disc_build_qubo_kernel<<<(unsigned)B, 256, (size_t)n * sizeof(double), (cudaStream_t)stream>>>(params, d, X, Y, beta_eff, Q_out);
    QBM_LAUNCH_OK("disc_build_qubo_kernel");
    return QBM_OK;
}

extern "C" QBM_API int qbm_disc_errors(int dim_input, int n_output, int n_hidden, int restricted, int faster_mode,
                                       const double *X, const double *Y, long long B, const double *mean_c,
                                       const double *second_c, const double *mean_u, const double *second_u, double *err_out,
                                       void *stream)
{
    if (int rc = check_dims("qbm_disc_errors", dim_input, n_output, n_hidden)) return rc;
    QBM_CHECK_ARG(X && Y && mean_c && mean_u && second_u && err_out, "qbm_disc_errors: null pointer argument");
    QBM_CHECK_ARG(B >= 1, "qbm_disc_errors: bad batch size");
    QBM_CHECK_ARG(restricted || faster_mode || second_c != nullptr, "qbm_disc_errors: second_c is required for <h h'>");
    DiscDims d{dim_input, n_output, n_hidden, restricted ? 1 : 0};
    const long long total = d.total() + 1;
    disc_errors_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d, faster_mode ? 1 : 0, B, X, Y, mean_c,
                                                                                         second_c, mean_u, second_u, err_out);
    QBM_LAUNCH_OK("disc_errors_kernel");
    return QBM_OK;
}

extern "C" QBM_API int qbm_sgd_apply(double *params, const double *err, long long count, double lr, double batch, void *stream)
{
    QBM_CHECK_ARG(params && err && count >= 0 && batch > 0.0, "qbm_sgd_apply: bad argument");
    if (count == 0) return QBM_OK;
    sgd_apply_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, err, count, lr, batch);
    QBM_LAUNCH_OK("sgd_apply_kernel");
    return QBM_OK;
}
