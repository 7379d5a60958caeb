// K4: TF32 tensor-core GEMM for the RBM/QBM dense contractions, written directly on the Blackwell
// pipeline: TMA (cp.async.bulk.tensor) stages 128x32 fp32 tiles of both operands into 128B-swizzled
// shared memory, one elected thread issues tcgen05.mma.kind::tf32 with the accumulator in tensor
// memory (TMEM), and the four warps read it back with tcgen05.ld for a fused epilogue (bias, per-row
// class bias, sigmoid, Bernoulli sampling, SGD accumulate, transposed copy).
//
// Computes  C[M,N] = epi( alpha * A[M,K] . B[N,K]^T )  with both operands K-major (row-major [rows, K]).
// Replaces torch.matmul / expand-mul-sum in src/ClassificationRBM.py:44-56,106,118-128:
//   x.W        : A = x [B,V],   B = W^T [H,V]
//   h.W^T      : A = h [B,H],   B = W   [V,H]
//   x^T.D      : A = x^T [V,B], B = D^T [H,B]      (epilogue: W += lr/B * acc)
#include "gemm.cuh"
#include <cuda.h>
#include <cstdlib>

namespace {

constexpr int BM = 128;          // rows of A per CTA  (= UMMA M, TMEM lanes)
// BN = rows of B per CTA (= UMMA N, TMEM columns) is a template parameter: the widest of 256 / 128 / 64 / 32 that still gives
// every SM a tile (256: twice the MMA work per byte staged; 32: a training step at batch 256 is a handful of tiles, and
// the time of such a launch is the epilogue of ONE tile -- 128 threads x BN columns each -- plus the depth of the K loop,
// so narrow tiles on many SMs and a deeper TMA ring (8 stages) cut it from ~30 us to the launch floor)
constexpr int BK = 32;           // K elements per stage: 32 x 4 B = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // K per tcgen05.mma for tf32 (32 bytes)
constexpr int A_TILE_BYTES = BM * BK * 4;                     // 16 KB per stage
template <int BN> __host__ __device__ constexpr int n_stages() { return BN <= 64 ? 8 : 4; }
template <int BN> __host__ __device__ constexpr int stage_bytes() { return A_TILE_BYTES + BN * BK * 4; }
template <int BN> __host__ __device__ constexpr int smem_bytes() { return n_stages<BN>() * stage_bytes<BN>() + 1024; }   // + slack for 1024-byte alignment

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar)) : "memory");
}
// K-major operand tile in 128B-swizzled shared memory: 8-row groups are 1024 B apart (SBO), LBO unused (=1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address  [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (16 B units) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell) [46,48)
    d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B [61,64)
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
template <int BN> __host__ __device__ constexpr uint32_t idesc()
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t IDESC, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Epilogue features.  FM = EPI_DYN: every feature is tested at run time (the generic instantiation: ~4700 instructions, of
// which a launch executes a fraction -- at batch 256 the instruction fetches of that epilogue were a measurable part of the
// launch).  Otherwise FM is the exact set of features of the call and the rest of the code does not exist.
constexpr unsigned EPI_C = 1u, EPI_CT = 2u, EPI_S = 4u, EPI_ST = 8u, EPI_BIAS = 16u, EPI_TAB = 32u, EPI_CIN = 64u, EPI_ACT = 128u;
constexpr unsigned EPI_DYN = 0xffffffffu;
template <unsigned FM, unsigned BIT> __device__ __forceinline__ bool epi_has(bool dyn)
{
    if constexpr (FM == EPI_DYN) return dyn;
    else return (FM & BIT) != 0u;
}

template <int BN, unsigned FM>
__global__ void __launch_bounds__(128, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams ep,
                 int M, int N, int K)
{
    constexpr int STAGES = n_stages<BN>();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = (K + BK - 1) / BK;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 2) {   // one warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer ----
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);                    // slot free (passes on the first lap)
            mbar_expect_tx(&full_bar[s], stage_bytes<BN>());      // OOB parts of a box are zero-filled and counted
            tma_load_2d(smem + (size_t)s * stage_bytes<BN>(), &tmA, kb * BK, m0, &full_bar[s]);
            tma_load_2d(smem + (size_t)s * stage_bytes<BN>() + A_TILE_BYTES, &tmB, kb * BK, n0, &full_bar[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer ----
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(&full_bar[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes<BN>());
            const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t ad = umma_desc_sw128(a_addr + k * UMMA_K * 4);
                const uint64_t bd = umma_desc_sw128(b_addr + k * UMMA_K * 4);
                umma_tf32(tmem_base, ad, bd, idesc<BN>(), (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);                           // frees the smem slot when these MMAs retire
        }
        umma_commit(&done_bar);                                   // accumulator complete
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w reads TMEM lanes 32w..32w+31 (row m0 + 32w + lane), 32 columns at a time ----
    const bool hasC = epi_has<FM, EPI_C>(ep.C != nullptr), hasCt = epi_has<FM, EPI_CT>(ep.Ct != nullptr);
    const bool hasS = epi_has<FM, EPI_S>(ep.S != nullptr), hasSt = epi_has<FM, EPI_ST>(ep.St != nullptr);
    const bool hasBias = epi_has<FM, EPI_BIAS>(ep.bias_n != nullptr), hasTab = epi_has<FM, EPI_TAB>(ep.rowtab != nullptr);
    const bool hasCin = epi_has<FM, EPI_CIN>(ep.Cin != nullptr), hasAct = epi_has<FM, EPI_ACT>(ep.act == 1);
    const int m = m0 + warp * 32 + lane;
    const bool row_ok = m < M;
    const float *tabrow = (hasTab && row_ok) ? ep.rowtab + (size_t)ep.ridx[m] * ep.ldtab : nullptr;
    const uint32_t k0 = (uint32_t)ep.seed, k1 = (uint32_t)(ep.seed >> 32);
    const uint32_t strm = ep.stream + (ep.step_dev != nullptr ? 4u * __ldg(ep.step_dev) : 0u);
    const bool vecc = hasC && (ep.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.C) & 15) == 0;
    const bool vecs = hasS && (ep.lds & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.S) & 15) == 0;
    const bool vecin = hasCin && (ep.ldcin & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.Cin) & 15) == 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int nb = n0 + c * 32;
        if (nb >= N) continue;                                    // uniform
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            Philox4 u4 = {0u, 0u, 0u, 0u};
            if (hasS || hasSt)
                u4 = philox4x32_10((uint32_t)m, (uint32_t)((nb >> 2) + j4), strm, 0x52424Du, k0, k1);
            const uint32_t us[4] = {u4.x, u4.y, u4.z, u4.w};
            // a lane owns 4 consecutive columns of its row here: the row-major outputs (C, S) and Cin move as one
            // 16-byte access per lane when the addresses allow it (rows of a warp are ld apart, so a 4-byte store per
            // lane would touch one sector per element); the transposed outputs are coalesced across lanes as they are
            const int nq = nb + j4 * 4;
            const bool full4 = row_ok && (nq + 3 < N);
            float cin4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (hasCin && row_ok) {
                if (full4 && vecin) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(ep.Cin + (size_t)m * ep.ldcin + nq);
                    cin4[0] = t4.x; cin4[1] = t4.y; cin4[2] = t4.z; cin4[3] = t4.w;
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (nq + jj < N) cin4[jj] = ep.Cin[(size_t)m * ep.ldcin + nq + jj];
                }
            }
            float v4[4], s4v[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int n = nq + jj;
                float v = ep.alpha * __uint_as_float(r[j4 * 4 + jj]);
                if (n < N) {
                    if (hasBias) v += __ldg(ep.bias_n + n);
                    if (hasTab && tabrow != nullptr) v += __ldg(tabrow + n);
                }
                if (hasAct) v = sigmoidf_(v);
                if (hasCin) v += ep.beta * cin4[jj];
                v4[jj] = v;
                // u in [0,1) with 24 bits; sample = 1 with probability v
                s4v[jj] = ((float)(us[jj] >> 8) * 5.9604644775390625e-8f < v) ? 1.0f : 0.0f;
            }
            if (row_ok) {
                if (hasC) {
                    if (full4 && vecc) *reinterpret_cast<float4 *>(ep.C + (size_t)m * ep.ldc + nq) = make_float4(v4[0], v4[1], v4[2], v4[3]);
                    else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (nq + jj < N) ep.C[(size_t)m * ep.ldc + nq + jj] = v4[jj];
                    }
                }
                if (hasS) {
                    if (full4 && vecs) *reinterpret_cast<float4 *>(ep.S + (size_t)m * ep.lds + nq) = make_float4(s4v[0], s4v[1], s4v[2], s4v[3]);
                    else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (nq + jj < N) ep.S[(size_t)m * ep.lds + nq + jj] = s4v[jj];
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int n = nq + jj;
                    if (n >= N) continue;
                    if (hasCt) ep.Ct[(size_t)n * ep.ldct + m] = v4[jj];
                    if (hasSt) ep.St[(size_t)n * ep.ldst + m] = s4v[jj];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 matrix [rows, cols] with leading dimension ld; box = 32 columns x box_rows rows, 128B swizzle
int make_map(CUtensorMap *map, const float *ptr, long long rows, long long cols, long long ld, int box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        qbm_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return QBM_ECUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qbm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld);
        return QBM_ECUDA;
    }
    return QBM_OK;
}

}  // namespace

// internal C++ entry used by rbm.cu as well
int qbm_gemm_tf32_launch(const float *A, long long lda, const float *B, long long ldb, int M, int N, int K,
                         const EpiParams &ep, cudaStream_t st)
{
    QBM_CHECK_ARG(A && B, "qbm_gemm_tf32: null operand");
    QBM_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "qbm_gemm_tf32: M, N, K must be >= 1");
    QBM_CHECK_ARG(lda >= K && ldb >= K && lda % 4 == 0 && ldb % 4 == 0,
                  "qbm_gemm_tf32: leading dimensions must be >= K and multiples of 4 floats (TMA needs 16-byte row strides)");
    QBM_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "qbm_gemm_tf32: operands must be 16-byte aligned");
    // the widest tile that still gives every SM a CTA; problems smaller than that take the narrowest
    const long long mt = (M + BM - 1) / BM;
    int bn = 32;
    for (int cand : {256, 128, 64})
        if (mt * ((N + cand - 1) / cand) >= 148) { bn = cand; break; }
    if (const char *e = getenv("QBM_GEMM_BN")) {            // measurement aid (tools/probe_gemm.py): force a tile width
        const int v = atoi(e);
        if (v == 32 || v == 64 || v == 128 || v == 256) bn = v;
    }
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_map(&tmB, B, N, K, ldb, bn);
    if (rc) return rc;
    const dim3 grid((unsigned)((N + bn - 1) / bn), (unsigned)mt);
    // the attribute belongs to the current device (a process may drive several), so it is set per launch
#define QBM_GEMM_LAUNCH(BN_, FM_)                                                                                              \
    do {                                                                                                                       \
        QBM_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN_, FM_>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                         smem_bytes<BN_>()));                                                                  \
        gemm_tf32_kernel<BN_, FM_><<<grid, 128, smem_bytes<BN_>(), st>>>(tmA, tmB, ep, M, N, K);                               \
    } while (0)
    if (bn >= 128) {
        if (bn == 256) QBM_GEMM_LAUNCH(256, EPI_DYN);
        else QBM_GEMM_LAUNCH(128, EPI_DYN);
    } else {
        // small problems (the training steps at batch 256): the feature sets the RBM steps use have their own instantiation
        const unsigned fm = (ep.C ? EPI_C : 0u) | (ep.Ct ? EPI_CT : 0u) | (ep.S ? EPI_S : 0u) | (ep.St ? EPI_ST : 0u) |
                            (ep.bias_n ? EPI_BIAS : 0u) | (ep.rowtab ? EPI_TAB : 0u) | (ep.Cin ? EPI_CIN : 0u) |
                            (ep.act == 1 ? EPI_ACT : 0u);
#define QBM_GEMM_CASE(FM_)                                                                                                     \
    case (FM_):                                                                                                                \
        if (bn == 64) QBM_GEMM_LAUNCH(64, (FM_));                                                                              \
        else QBM_GEMM_LAUNCH(32, (FM_));                                                                                       \
        break;
        switch (fm) {
            QBM_GEMM_CASE(EPI_C | EPI_BIAS)                                              // x.W + b_h
            QBM_GEMM_CASE(EPI_C | EPI_CIN | EPI_CT)                                      // W += s x^T.D, W^T
            QBM_GEMM_CASE(EPI_C | EPI_CIN)                                               // accumulate
            QBM_GEMM_CASE(EPI_C)                                                         // gradient
            QBM_GEMM_CASE(EPI_C | EPI_CT | EPI_S | EPI_BIAS | EPI_TAB | EPI_ACT)         // CD-1: ph0, ph0^T, h0
            QBM_GEMM_CASE(EPI_S | EPI_ST | EPI_BIAS | EPI_ACT)                           // CD-1: v1, v1^T
            QBM_GEMM_CASE(EPI_CT | EPI_BIAS | EPI_TAB | EPI_ACT)                         // CD-1: ph1^T
            QBM_GEMM_CASE(EPI_C | EPI_BIAS | EPI_TAB | EPI_ACT)                          // sample_hidden
            QBM_GEMM_CASE(EPI_C | EPI_BIAS | EPI_ACT)                                    // sample_visible
            default:
                if (bn == 64) QBM_GEMM_LAUNCH(64, EPI_DYN);
                else QBM_GEMM_LAUNCH(32, EPI_DYN);
                break;
        }
#undef QBM_GEMM_CASE
    }
#undef QBM_GEMM_LAUNCH
    QBM_LAUNCH_OK("gemm_tf32_kernel");
    return QBM_OK;
}

extern "C" QBM_API int qbm_gemm_tf32(const float *A, long long lda, const float *B, long long ldb, int M, int N, int K,
                                     float alpha, float beta, const float *Cin, long long ldcin, const float *bias_n,
                                     int act, float *C, long long ldc, float *Ct, long long ldct, void *stream)
{
    QBM_CHECK_ARG(C != nullptr || Ct != nullptr, "qbm_gemm_tf32: no output given");
    QBM_CHECK_ARG(act == 0 || act == 1, "qbm_gemm_tf32: act must be 0 (identity) or 1 (sigmoid)");
    EpiParams ep = {};
    ep.C = C; ep.ldc = ldc; ep.Ct = Ct; ep.ldct = ldct; ep.bias_n = bias_n; ep.Cin = Cin; ep.ldcin = ldcin;
    ep.alpha = alpha; ep.beta = beta; ep.act = act;
    return qbm_gemm_tf32_launch(A, lda, B, ldb, M, N, K, ep, (cudaStream_t)stream);
}
