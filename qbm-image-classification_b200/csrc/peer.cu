// Peer-memory plumbing of the data-parallel ClassificationRBM step: the all-reduce of the flat
// gradient buffer and the parameter update as ONE pass over NVLink peer memory instead of an NCCL all-reduce followed by
// an apply kernel.  Every rank owns one IPC-shared allocation  [ gradient buffer, two halves | arrival flags | error word ]
// that all ranks of the node map (cudaIpc*); after its gradient kernels a rank stores the step's token into its slot of
// every rank's flag array, and the apply kernel of every rank waits for all slots of its own array, then reads the same
// element of every rank's gradient (rank order: the sum is deterministic and identical everywhere) and applies it.  The two
// gradient halves alternate by step parity, which is what makes one flag round per step sufficient: a rank can only start
// writing half (t + 2) & 1 after every peer has signalled step t + 1, i.e. finished reading half t & 1.
//
// The reference has no distributed layer; this is the sum-then-divide-by-the-global-batch order of
// src/ClassificationRBM.py:88-99 applied to gradient sums of minibatch shards.
#include "common.cuh"
#include <string.h>

namespace {

constexpr int MAX_PEERS = 16;
constexpr int FLAG_WORDS = 64;                 // arrival flags [world] padded; word FLAG_WORDS is the error flag
constexpr long long SPIN_LIMIT = 240000000000LL; // ~2 min of SM clocks: a slow rank is waited for (as NCCL would), a dead one
                                                 // must not hang the GPU for ever

__host__ __device__ inline long long pld4(long long c) { return (c + 3) & ~3LL; }
inline size_t peer_grad_floats(int V, int H, int C)
{
    const size_t lH = pld4(H);
    return (size_t)V * lH + (size_t)C * lH + pld4(V) + pld4(H) + pld4(C) + 4;
}

struct PeerSet {
    const float *grad[MAX_PEERS];      // this step's half of every rank's gradient buffer
    unsigned int *flags[MAX_PEERS];    // every rank's arrival flags
    int world;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// "the gradient of step `token` of rank `rank` is complete", written to every rank's flag array (own included)
__global__ void rbm_peer_signal_kernel(PeerSet ps, int rank, unsigned int token_host, const unsigned int *__restrict__ tick_dev)
{
    const unsigned int token = token_host + (tick_dev != nullptr ? *tick_dev : 0u);
    __threadfence_system();
    if ((int)threadIdx.x < ps.world)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ps.flags[threadIdx.x] + rank), "r"(token) : "memory");
}

// sum over the ranks in rank order.  The loads are issued together (a remote load is ~2 us of NVLink latency: eight of them
// one after the other, four elements per thread, were most of the pass) and bypass L1, which is never coherent with peers
__device__ __forceinline__ float peer_sum(const PeerSet &ps, size_t off)
{
    float v[MAX_PEERS];
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p) v[p] = (p < ps.world) ? __ldcg(ps.grad[p] + off) : 0.0f;
    float s = 0.0f;
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
        if (p < ps.world) s += v[p];
    return s;
}

// rbm_apply_kernel (rbm.cu) with the gradient summed over the ranks on the fly
__global__ void __launch_bounds__(256) rbm_apply_peer_kernel(float *__restrict__ W, float *__restrict__ Wt, float *__restrict__ U,
                                                            float *__restrict__ b_v, float *__restrict__ b_h, float *__restrict__ b_c,
                                                            PeerSet ps, int rank, unsigned int token_host,
                                                            const unsigned int *__restrict__ tick_dev, int V, int H, int C,
                                                            long long lH, long long lV, float scale, float sparse,
                                                            float *__restrict__ loss_out, float loss_scale, int tiles_x, int tiles)
{
    __shared__ float tile[32][33];
    const unsigned int token = token_host + (tick_dev != nullptr ? *tick_dev : 0u);
    __shared__ int give_up;
    if (threadIdx.x == 0) give_up = (ld_acquire_sys(ps.flags[rank] + FLAG_WORDS) != 0u);      // sticky: a broken group stays broken
    __syncthreads();
    if ((int)threadIdx.x < ps.world && !give_up) {
        const unsigned int *f = ps.flags[rank] + threadIdx.x;
        const long long t0 = clock64();
        bool ok = false;
        do {
            ok = (int)(ld_acquire_sys(f) - token) >= 0;
        } while (!ok && clock64() - t0 < SPIN_LIMIT);
        if (!ok) { atomicExch(ps.flags[rank] + FLAG_WORDS, 1u); give_up = 1; }
    }
    __syncthreads();
    if (give_up) return;                    // no update from a gradient that never arrived; the host raises (qbm_rbm_peer_error)
    const size_t oU = (size_t)V * lH, obv = oU + (size_t)C * lH, obh = obv + pld4(V), obc = obh + pld4(H), oloss = obc + pld4(C);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if ((int)blockIdx.x < tiles) {
        const int h0 = ((int)blockIdx.x % tiles_x) * 32, v0 = ((int)blockIdx.x / tiles_x) * 32;
        float g[4], w0[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {               // all 4 x world remote loads of this thread in flight at once
            const int v = v0 + ty + 8 * k, h = h0 + tx;
            const bool in = v < V && h < H;
            g[k] = in ? peer_sum(ps, (size_t)v * lH + h) : 0.0f;
            w0[k] = in ? W[(size_t)v * lH + h] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = v0 + ty + 8 * k, h = h0 + tx;
            const float w = w0[k] + scale * g[k];
            if (v < V && h < H) W[(size_t)v * lH + h] = w;
            tile[ty + 8 * k][tx] = (v < V && h < H) ? w : 0.0f;
        }
        __syncthreads();
        for (int i = ty; i < 32; i += 8) {
            const int h = h0 + i, v = v0 + tx;
            if (h < H && v < V) Wt[(size_t)h * lV + v] = tile[tx][i];
        }
        return;
    }
    const int e0 = ((int)blockIdx.x - tiles) * 256 + threadIdx.x;
    const int stride = ((int)gridDim.x - tiles) * 256;
    for (int e = e0; e < C * H; e += stride) {
        const int c = e / H, h = e % H;
        U[(size_t)c * lH + h] += scale * peer_sum(ps, oU + (size_t)c * lH + h);
    }
    for (int v = e0; v < V; v += stride) b_v[v] = b_v[v] + scale * peer_sum(ps, obv + v) - sparse;
    for (int h = e0; h < H; h += stride) b_h[h] = b_h[h] + scale * peer_sum(ps, obh + h) - sparse;
    for (int c = e0; c < C; c += stride) b_c[c] = b_c[c] + scale * peer_sum(ps, obc + c) - sparse;
    if (e0 == 0 && loss_out != nullptr) loss_out[0] = peer_sum(ps, oloss) * loss_scale;
}

}  // namespace

extern "C" QBM_API size_t qbm_rbm_peer_bytes(int V, int H, int C)
{
    if (V < 1 || H < 1 || C < 1) return 0;
    return 2 * peer_grad_floats(V, H, C) * sizeof(float) + (FLAG_WORDS + 4) * sizeof(unsigned int);
}

extern "C" QBM_API int qbm_peer_alloc(size_t bytes, void **ptr)
{
    QBM_CHECK_ARG(ptr && bytes > 0, "qbm_peer_alloc: bad arguments");
    QBM_CUDA_OK(cudaMalloc(ptr, bytes));                       // plain cudaMalloc: the only kind cudaIpcGetMemHandle accepts
    QBM_CUDA_OK(cudaMemset(*ptr, 0, bytes));
    return QBM_OK;
}

extern "C" QBM_API int qbm_peer_free(void *ptr)
{
    if (ptr != nullptr) QBM_CUDA_OK(cudaFree(ptr));
    return QBM_OK;
}

extern "C" QBM_API int qbm_peer_export(void *ptr, unsigned char *handle64)
{
    QBM_CHECK_ARG(ptr && handle64, "qbm_peer_export: null pointer argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
    cudaIpcMemHandle_t h;
    QBM_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return QBM_OK;
}

extern "C" QBM_API int qbm_peer_import(const unsigned char *handle64, void **ptr)
{
    QBM_CHECK_ARG(ptr && handle64, "qbm_peer_import: null pointer argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    QBM_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QBM_OK;
}

extern "C" QBM_API int qbm_peer_close(void *ptr)
{
    if (ptr != nullptr) QBM_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return QBM_OK;
}

// 1 when a wait for the peers' gradients ever timed out on this rank (the update of that step is then meaningless)
extern "C" QBM_API int qbm_rbm_peer_error(const void *own_base, int V, int H, int C, unsigned int *flag_out)
{
    QBM_CHECK_ARG(own_base && flag_out, "qbm_rbm_peer_error: null pointer argument");
    const unsigned int *f = reinterpret_cast<const unsigned int *>(reinterpret_cast<const float *>(own_base) + 2 * peer_grad_floats(V, H, C));
    QBM_CUDA_OK(cudaMemcpy(flag_out, f + FLAG_WORDS, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return QBM_OK;
}

// update_weights (ref :88-99) from the gradient sums of ALL ranks, read through peer memory: signal, wait, reduce, apply.
//   peer_bases  host array [world] of the ranks' qbm_rbm_peer_bytes allocations as mapped into this process
//   parity      which half of the gradient buffers holds this step (the caller alternates it every step)
//   token       step number (+ *tick_dev when given), strictly increasing by one per step on every rank
extern "C" QBM_API int qbm_rbm_apply_grad_peer(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c,
                                               void *const *peer_bases, int world, int rank, int parity, int V, int H, int C,
                                               float scale, float sparse_constant, float *loss_out, float loss_scale,
                                               unsigned int token, const unsigned int *tick_dev, void *stream)
{
    QBM_CHECK_ARG(W && Wt && U && b_v && b_h && b_c && peer_bases, "qbm_rbm_apply_grad_peer: null pointer argument");
    QBM_CHECK_ARG(V >= 1 && H >= 1 && C >= 1, "qbm_rbm_apply_grad_peer: bad dimensions");
    QBM_CHECK_ARG(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world && (parity == 0 || parity == 1),
                  "qbm_rbm_apply_grad_peer: need 1 <= world <= %d, 0 <= rank < world, parity 0 or 1", MAX_PEERS);
    const size_t cnt = peer_grad_floats(V, H, C);
    PeerSet ps = {};
    ps.world = world;
    for (int p = 0; p < world; ++p) {
        QBM_CHECK_ARG(peer_bases[p], "qbm_rbm_apply_grad_peer: null peer allocation");
        float *base = reinterpret_cast<float *>(peer_bases[p]);
        ps.grad[p] = base + (size_t)parity * cnt;
        ps.flags[p] = reinterpret_cast<unsigned int *>(base + 2 * cnt);
    }
    cudaStream_t st = (cudaStream_t)stream;
    rbm_peer_signal_kernel<<<1, 32, 0, st>>>(ps, rank, token, tick_dev);
    QBM_LAUNCH_OK("rbm_peer_signal_kernel");
    const int tx = (H + 31) / 32, ty = (V + 31) / 32;
    const int tiles = tx * ty;
    rbm_apply_peer_kernel<<<tiles + 8, 256, 0, st>>>(W, Wt, U, b_v, b_h, b_c, ps, rank, token, tick_dev, V, H, C, pld4(H), pld4(V),
                                                     scale, sparse_constant, loss_out, loss_scale, tx, tiles);
    QBM_LAUNCH_OK("rbm_apply_peer_kernel");
    return QBM_OK;
}
