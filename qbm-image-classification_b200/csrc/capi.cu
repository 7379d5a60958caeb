// C-ABI plumbing: version, thread-local error string, device info and the device test hooks.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void qbm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" QBM_API int qbm_version(void) { return 1; }
extern "C" QBM_API const char *qbm_last_error(void) { return g_err; }

extern "C" QBM_API int qbm_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    QBM_CUDA_OK(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { QBM_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { QBM_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { QBM_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return QBM_OK;
}

namespace {
__global__ void test_philox_kernel(const uint32_t *ctr, const uint32_t *key, uint32_t *out, long long count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Philox4 o = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
    out[4 * i] = o.x; out[4 * i + 1] = o.y; out[4 * i + 2] = o.z; out[4 * i + 3] = o.w;
}
__global__ void test_neg_log_kernel(const uint32_t *u, float *out, long long count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = neg_log_u32(u[i]);
}
}  // namespace

extern "C" QBM_API int qbm_test_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out, long long count, void *stream)
{
    QBM_CHECK_ARG(ctr && key && out && count >= 0, "qbm_test_philox: bad argument");
    if (count == 0) return QBM_OK;
    test_philox_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctr, key, out, count);
    QBM_LAUNCH_OK("test_philox_kernel");
    return QBM_OK;
}

extern "C" QBM_API int qbm_test_neg_log(const uint32_t *u, float *out, long long count, void *stream)
{
    QBM_CHECK_ARG(u && out && count >= 0, "qbm_test_neg_log: bad argument");
    if (count == 0) return QBM_OK;
    test_neg_log_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, out, count);
    QBM_LAUNCH_OK("test_neg_log_kernel");
    return QBM_OK;
}
