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
