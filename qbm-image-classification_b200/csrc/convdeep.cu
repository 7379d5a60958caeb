// K6: Conv-Deep inference context on the device -- valid 2-D convolution with the shared kernel,
// deterministic p x p pooling (the unit with the SMALLEST feature-map value of each window is the
// active one) and the input patch of every active unit.
//
// Replaces, for a whole minibatch in one launch, the per-image Python loops of
//   src/model/geometry.py:37-53  (conv2d_valid_stride: out[i,j] = np.sum(img[ii:ii+k, jj:jj+k] * kernel))
//   src/model/layers.py:65-84    (pooled_indices_for_input: ids[np.argmin(fmap_flat[ids])] per window)
//   src/train/train.py:188-191   (x_input[np.ix_(rows, cols)] of input_groups[pooled_idx[i]])
// reached through src/model/inference.py:16-44 (prepare_context).
//
// The feature map is reproduced BIT FOR BIT: products in float64 and numpy's pairwise summation order
// for a contiguous k*k block (8 running sums, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the
// tail), so the argmin -- and with it the variable layout of the QUBO -- cannot differ from the
// reference's through rounding.
#include "common.cuh"

namespace {

constexpr int MAX_KK = 128;   // numpy switches to recursive halving above 128 elements

__device__ __forceinline__ double numpy_pairwise_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// one CTA per image
__global__ void __launch_bounds__(256) convdeep_context_kernel(const double *__restrict__ X, const double *__restrict__ kern,
                                                              int ih, int iw, int k, int stride, int pool, int oh, int ow,
                                                              int ph, int pw, double *__restrict__ fmap,
                                                              int *__restrict__ pooled, double *__restrict__ patches)
{
    const size_t b = blockIdx.x;
    const double *x = X + b * (size_t)ih * (size_t)iw;
    double *fm = fmap + b * (size_t)oh * (size_t)ow;
    const int kk = k * k;
    for (int u = threadIdx.x; u < oh * ow; u += blockDim.x) {
        const int i0 = (u / ow) * stride, j0 = (u % ow) * stride;
        double prod[MAX_KK];
        for (int e = 0; e < kk; ++e)
            prod[e] = __dmul_rn(x[(size_t)(i0 + e / k) * iw + (j0 + e % k)], __ldg(kern + e));
        fm[u] = numpy_pairwise_sum(prod, kk);
    }
    __syncthreads();
    const int P = (pool <= 1) ? oh * ow : ph * pw;
    int *pi = pooled + b * (size_t)P;
    double *pt = patches == nullptr ? nullptr : patches + b * (size_t)P * (size_t)kk;
    for (int w = threadIdx.x; w < P; w += blockDim.x) {
        int pick;
        if (pool <= 1) {
            pick = w;                                             // no windows configured: keep every unit
        } else {
            const int wi = (w / pw) * pool, wj = (w % pw) * pool;
            pick = wi * ow + wj;
            double best = fm[pick];
            for (int di = 0; di < pool; ++di)
                for (int dj = 0; dj < pool; ++dj) {
                    const int id = (wi + di) * ow + (wj + dj);
                    const double v = fm[id];
                    if (v < best) { best = v; pick = id; }        // strict '<': first minimum, like np.argmin
                }
        }
        pi[w] = pick;
        if (pt != nullptr) {
            const int i0 = (pick / ow) * stride, j0 = (pick % ow) * stride;
            for (int e = 0; e < kk; ++e) pt[(size_t)w * kk + e] = x[(size_t)(i0 + e / k) * iw + (j0 + e % k)];
        }
    }
}

}  // namespace

extern "C" QBM_API int qbm_convdeep_num_pooled(int ih, int iw, int k, int stride, int pool)
{
    if (ih < k || iw < k || k < 1 || stride < 1 || pool < 0) return -1;
    const int oh = (ih - k) / stride + 1, ow = (iw - k) / stride + 1;
    if (pool <= 1) return oh * ow;
    return (oh / pool) * (ow / pool);
}

extern "C" QBM_API int qbm_convdeep_context(const double *X, const double *kernel, long long B, int ih, int iw, int k,
                                            int stride, int pool, double *fmap_out, int *pooled_out, double *patches_out,
                                            void *stream)
{
    QBM_CHECK_ARG(X && kernel && fmap_out && pooled_out, "qbm_convdeep_context: null pointer argument");
    QBM_CHECK_ARG(B >= 1 && B <= 0x7fffffffLL, "qbm_convdeep_context: bad batch size");
    QBM_CHECK_ARG(k >= 1 && stride >= 1 && pool >= 0 && ih >= k && iw >= k,
                  "qbm_convdeep_context: bad geometry (image %dx%d kernel %d stride %d pool %d)", ih, iw, k, stride, pool);
    if (k * k > MAX_KK) {
        qbm_set_error("qbm_convdeep_context: kernel_size %d not supported (k*k must be <= %d)", k, MAX_KK);
        return QBM_EUNSUPPORTED;
    }
    const int oh = (ih - k) / stride + 1, ow = (iw - k) / stride + 1;
    const int ph = pool > 1 ? oh / pool : 0, pw = pool > 1 ? ow / pool : 0;
    QBM_CHECK_ARG(pool <= 1 || (ph >= 1 && pw >= 1), "qbm_convdeep_context: pooling window larger than the feature map");
    convdeep_context_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(X, kernel, ih, iw, k, stride, pool, oh, ow, ph, pw,
                                                                          fmap_out, pooled_out, patches_out);
    QBM_LAUNCH_OK("convdeep_context_kernel");
    return QBM_OK;
}

// ================================================================================================
// K10 / K11: the Conv-Deep training step around the sampler, one launch each per minibatch (float64).
//   K10 convdeep_build_qubo  build_unclamped_qubo / build_clamped_qubo        src/qubo/builder.py:21-110
//   K11 convdeep_errors      get_average_configuration_single, clamped - unclamped, summed over the minibatch,
//                            + the loss of train_one_iteration                 src/train/train.py:12-253
// Variable layout of a QUBO: [pooled conv units (P) | sequential layers | outputs (unclamped only)].
// Parameter layout (one flat float64 buffer, also the layout of the error buffer):
//   [ b_conv (1 when shared, else absent) | b_seq | b_out | kernel (k*k) | W_seq[0..L-1] | W_intra[0..L-1] (absent
//     when restricted) | W_hy (last x nl) | W_oo (nl x nl) ]
// ================================================================================================
namespace {

constexpr int CD_MAXL = 8;

struct ConvDeepDims {
    int P, L, nl, kk, restricted, shared_bias;
    int sizes[CD_MAXL];
    // variable index where layer li starts: li = 0 pooled, 1..L sequential, L+1 outputs
    __host__ __device__ int start(int li) const
    {
        int s = 0;
        if (li >= 1) s = P;
        for (int i = 1; i < li; ++i) s += sizes[i - 1];
        return s;
    }
    __host__ __device__ int width(int li) const { return li == 0 ? P : (li <= L ? sizes[li - 1] : nl); }
    __host__ __device__ int nh() const { return start(L + 1); }
    __host__ __device__ int nseq() const { return nh() - P; }
    __host__ __device__ int layer_of(int v) const
    {
        int li = 0;
        while (li <= L && v >= start(li + 1)) ++li;
        return li;
    }
    __host__ __device__ long long off_bconv() const { return 0; }
    __host__ __device__ long long off_bseq() const { return shared_bias ? 1 : 0; }
    __host__ __device__ long long off_bout() const { return off_bseq() + nseq(); }
    __host__ __device__ long long off_kernel() const { return off_bout() + nl; }
    __host__ __device__ long long off_wseq(int li) const
    {
        long long o = off_kernel() + kk;
        for (int i = 0; i < li; ++i) o += (long long)width(i) * width(i + 1);
        return o;
    }
    __host__ __device__ long long off_wintra(int li) const
    {
        long long o = off_wseq(L);
        for (int i = 0; i < li; ++i) o += (long long)sizes[i] * sizes[i];
        return o;
    }
    __host__ __device__ long long off_why() const { return restricted ? off_wseq(L) : off_wintra(L); }
    __host__ __device__ long long off_woo() const { return off_why() + (long long)width(L) * nl; }
    __host__ __device__ long long total() const { return off_woo() + (long long)nl * nl; }
};

// one CTA per image
__global__ void __launch_bounds__(256) convdeep_build_qubo_kernel(const double *__restrict__ Pm, const ConvDeepDims d,
                                                                  const double *__restrict__ fmap, const int num_conv,
                                                                  const int *__restrict__ pooled, const double *__restrict__ Y,
                                                                  const double beta_eff, double *__restrict__ Q)
{
    const size_t b = blockIdx.x;
    const bool clamped = Y != nullptr;
    const int nh = d.nh(), n = clamped ? nh : nh + d.nl;
    const int L = d.L, ls = d.start(L), lw = d.width(L);
    const double *fm = fmap + b * (size_t)num_conv;
    const int *pi = pooled + b * (size_t)d.P;
    const double *y = clamped ? Y + b * (size_t)d.nl : nullptr;
    double *q = Q + b * (size_t)n * (size_t)n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int r = e / n, c = e % n;
        const int lr = d.layer_of(r), lc = d.layer_of(c);
        double v = 0.0;
        if (r == c) {
            if (lr == 0) {
                v = fm[pi[r]];
                if (d.shared_bias) v += Pm[d.off_bconv()];
            } else if (lr <= L) {
                v = Pm[d.off_bseq() + (r - d.P)];
            } else {
                v = Pm[d.off_bout() + (r - nh)];
            }
            if (clamped && r >= ls && r < ls + lw) {
                double eff = 0.0;                                   // W_hy @ label (builder.py:106-108)
                for (int o = 0; o < d.nl; ++o) eff = fma(Pm[d.off_why() + (size_t)(r - ls) * d.nl + o], y[o], eff);
                v += eff;
            }
        } else if (lc == L + 1) {
            if (lr == L) v = Pm[d.off_why() + (size_t)(r - ls) * d.nl + (c - nh)];                   // last hidden -> outputs
            else if (lr == L + 1 && c > r) v = Pm[d.off_woo() + (size_t)(r - nh) * d.nl + (c - nh)]; // triu(W_oo, 1)
        } else if (lr <= L && lc == lr + 1) {
            v = Pm[d.off_wseq(lr) + (size_t)(r - d.start(lr)) * d.width(lc) + (c - d.start(lc))];     // layer -> next layer
        } else if (lr == lc && lr >= 1 && lr <= L && c > r && !d.restricted) {
            const int s0 = d.start(lr), w = d.width(lr);
            v = Pm[d.off_wintra(lr - 1) + (size_t)(r - s0) * w + (c - s0)];                           // triu(W_intra, 1)
        }
        q[e] = v / beta_eff;
    }
}

// one thread per element of the error buffer (+ 1 for the loss); sums over the local images in image order
__global__ void convdeep_errors_kernel(const ConvDeepDims d, const long long B, const int round32, const int one_hot,
                                       const double *__restrict__ patches, const double *__restrict__ Y,
                                       const int *__restrict__ ylab, const double *__restrict__ Mc, const double *__restrict__ Sc,
                                       const double *__restrict__ Mu, const double *__restrict__ Su, double *__restrict__ E)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = d.total();
    if (e > total) return;
    const int nh = d.nh(), nu = nh + d.nl, P = d.P, nl = d.nl, L = d.L;
    // the reference averages float32 samples (src/qubo/sampler.py:33): moments optionally rounded through float32
    auto R = [round32](double v) { return round32 ? (double)(float)v : v; };
    auto mc = [&](long long b, int i) { return R(Mc[b * nh + i]); };
    auto mu = [&](long long b, int i) { return R(Mu[b * nu + i]); };
    auto sc = [&](long long b, int i, int j) { return R(Sc[(b * nh + i) * nh + j]); };
    auto su = [&](long long b, int i, int j) { return R(Su[(b * nu + i) * nu + j]); };
    double s = 0.0;
    if (e == total) {
        // loss of the unclamped phase (src/train/pipeline.py:22-28, src/train/train.py:45-50): float32 probabilities
        for (long long b = 0; b < B; ++b) {
            float py;
            if (!one_hot) {
                double p1 = (double)(float)Mu[b * nu + nh];
                p1 = fmin(fmax(p1, 1e-12), 1.0 - 1e-12);
                py = ylab[b] ? (float)p1 : (float)(1.0 - p1);
            } else {
                float tot = 0.0f;
                for (int o = 0; o < nl; ++o) tot += (float)Mu[b * nu + nh + o];
                py = tot > 0.0f ? (float)Mu[b * nu + nh + ylab[b]] / tot : 1.0f / (float)nl;
            }
            s -= (double)logf(fmaxf(py, 1e-12f));
        }
    } else if (d.shared_bias && e < d.off_bseq()) {
        for (long long b = 0; b < B; ++b)
            for (int i = 0; i < P; ++i) s += mc(b, i) - mu(b, i);
    } else if (e < d.off_bout()) {
        const int j = P + (int)(e - d.off_bseq());
        for (long long b = 0; b < B; ++b) s += mc(b, j) - mu(b, j);
    } else if (e < d.off_kernel()) {
        const int o = (int)(e - d.off_bout());
        for (long long b = 0; b < B; ++b) s += Y[b * nl + o] - mu(b, nh + o);
    } else if (e < d.off_wseq(0)) {
        const int ij = (int)(e - d.off_kernel());
        for (long long b = 0; b < B; ++b)
            for (int i = 0; i < P; ++i) s = fma(patches[(b * P + i) * d.kk + ij], mc(b, i) - mu(b, i), s);
    } else if (e < d.off_wseq(L)) {
        int li = 0;
        while (e >= d.off_wseq(li + 1)) ++li;
        const long long r = e - d.off_wseq(li);
        const int wc = d.width(li + 1);
        const int i = d.start(li) + (int)(r / wc), j = d.start(li + 1) + (int)(r % wc);
        for (long long b = 0; b < B; ++b) s += sc(b, i, j) - su(b, i, j);
    } else if (e < d.off_why()) {
        int li = 0;
        while (e >= d.off_wintra(li + 1)) ++li;
        const long long r = e - d.off_wintra(li);
        const int w = d.sizes[li], s0 = d.start(li + 1);
        const int i = (int)(r / w), j = (int)(r % w);
        if (i < j)
            for (long long b = 0; b < B; ++b) s += sc(b, s0 + i, s0 + j) - su(b, s0 + i, s0 + j);
    } else if (e < d.off_woo()) {
        const long long r = e - d.off_why();
        const int i = d.start(L) + (int)(r / nl), o = (int)(r % nl);
        for (long long b = 0; b < B; ++b) s += mc(b, i) * Y[b * nl + o] - su(b, i, nh + o);
    } else {
        const long long r = e - d.off_woo();
        const int o = (int)(r / nl), o2 = (int)(r % nl);
        if (o < o2)
            for (long long b = 0; b < B; ++b) s += Y[b * nl + o] * Y[b * nl + o2] - su(b, nh + o, nh + o2);
    }
    E[e] = s;
}

int make_dims(const char *who, int P, int L, const int *sizes, int nl, int k, int restricted, int shared_bias, ConvDeepDims *d)
{
    if (P < 1 || L < 0 || L > CD_MAXL || nl < 1 || k < 1 || (L > 0 && sizes == nullptr)) {
        qbm_set_error("%s: bad model structure (P=%d layers=%d labels=%d kernel=%d; at most %d sequential layers)", who, P, L, nl, k,
                      CD_MAXL);
        return QBM_EINVAL;
    }
    d->P = P; d->L = L; d->nl = nl; d->kk = k * k; d->restricted = restricted ? 1 : 0; d->shared_bias = shared_bias ? 1 : 0;
    for (int i = 0; i < CD_MAXL; ++i) d->sizes[i] = i < L ? sizes[i] : 0;
    for (int i = 0; i < L; ++i)
        if (sizes[i] < 1) { qbm_set_error("%s: sequential layer %d has size %d", who, i, sizes[i]); return QBM_EINVAL; }
    return QBM_OK;
}

}  // namespace

extern "C" QBM_API long long qbm_convdeep_param_count(int P, int num_layers, const int *layer_sizes, int n_labels, int kernel_size,
                                                      int restricted, int shared_bias)
{
    ConvDeepDims d;
    if (make_dims("qbm_convdeep_param_count", P, num_layers, layer_sizes, n_labels, kernel_size, restricted, shared_bias, &d)) return 0;
    return d.total();
}

extern "C" QBM_API int qbm_convdeep_build_qubo(const double *params, int P, int num_layers, const int *layer_sizes, int n_labels,
                                               int kernel_size, int restricted, int shared_bias, const double *fmap, int num_conv,
                                               const int *pooled, const double *Y, long long B, double beta_eff, double *Q_out,
                                               void *stream)
{
    ConvDeepDims d;
    if (int rc = make_dims("qbm_convdeep_build_qubo", P, num_layers, layer_sizes, n_labels, kernel_size, restricted, shared_bias, &d))
        return rc;
    QBM_CHECK_ARG(params && fmap && pooled && Q_out, "qbm_convdeep_build_qubo: null pointer argument");
    QBM_CHECK_ARG(B >= 1 && B <= 0x7fffffffLL && num_conv >= P, "qbm_convdeep_build_qubo: bad sizes");
    QBM_CHECK_ARG(beta_eff != 0.0, "qbm_convdeep_build_qubo: beta_eff must not be 0");
    convdeep_build_qubo_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(params, d, fmap, num_conv, pooled, Y, beta_eff, Q_out);
    QBM_LAUNCH_OK("convdeep_build_qubo_kernel");
    return QBM_OK;
}

extern "C" QBM_API int qbm_convdeep_errors(int P, int num_layers, const int *layer_sizes, int n_labels, int kernel_size, int restricted,
                                           int shared_bias, int round_float32, int one_hot, const double *patches, const double *Y,
                                           const int *labels, long long B, const double *mean_c, const double *second_c,
                                           const double *mean_u, const double *second_u, double *err_out, void *stream)
{
    ConvDeepDims d;
    if (int rc = make_dims("qbm_convdeep_errors", P, num_layers, layer_sizes, n_labels, kernel_size, restricted, shared_bias, &d)) return rc;
    QBM_CHECK_ARG(patches && Y && labels && mean_c && second_c && mean_u && second_u && err_out, "qbm_convdeep_errors: null pointer argument");
    QBM_CHECK_ARG(B >= 1, "qbm_convdeep_errors: bad batch size");
    const long long total = d.total() + 1;
    convdeep_errors_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d, B, round_float32 ? 1 : 0, one_hot ? 1 : 0, patches, Y, labels, mean_c, second_c, mean_u, second_u, err_out);
    QBM_LAUNCH_OK("convdeep_errors_kernel");
    return QBM_OK;
}
