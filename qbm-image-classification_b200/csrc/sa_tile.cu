// K1b: simulated-annealing QUBO sampler for sm_100a -- a register tile of 16 chains per CTA.
//
// Same trajectory as sa_kernel.cu (oracle/replay_sa.c), different execution.  In
// the hot part of neal's legacy schedule nearly every proposal is accepted (acceptance > 0.9 for the
// first ~10 % of the sweeps, where ~97 % of all flips happen), so every chain needs nearly every
// coupling row in every sweep.  One warp per chain streams a row per flip per chain through L1 (the
// 128 B/clk/SM pipe is the wall).  Here the 16 chains of a CTA advance in lock-step over 32-variable
// sub-windows and each coupling row is fetched ONCE for all of them:
//
//   * fields: thread (warp w, lane l) holds NS columns x 16 chains in registers (column = variable
//     (jw*W + w)*128 + k*32 + l, i.e. lane-aligned sub-windows of the 128-variable windows warp w owns)
//   * per sub-window s (32 consecutive variables, owned by one warp):
//       pre-check  owner, lane = variable: can any chain accept anything here (dE < 44.36142/beta)?
//       bounds     all warps: Philox + -ln(u)/beta for the 128-variable window, once per window and sweep
//       scan       owner, lane = chain (16 chains x 2 halves): the 32 variables are visited in sweep order;
//                  a flip updates the 32 fields of the sub-window from the 32x32 diagonal block of J in
//                  shared memory.  Output: per chain a flip mask, the old spins and a coefficient matrix
//       update     all warps: for every flipped variable a (in sweep order) the row J[a] is read once
//                  from the ring and applied to all chains that flipped a: F[.][t] = fma(c_t, J[a][.], F[.][t])
//   * a producer warp feeds the ring: one cp.async.bulk (TMA, 1-D) per coupling row, one mbarrier
//     full/empty hand-shake per ring slot of GR rows; rows come from L2 (the matrix is read once per sweep
//     and SM, not once per flip and chain)
//
// Every field element receives exactly the FMA sequence of the sequential rule, in the same order, so
// the final states are bit-identical to the replay oracle (tests/test_gpu_sa.py).
//
// Use: as a whole-schedule sampler behind qbm_sa_sample flag bit 4, and -- by default for n > 1792 -- for the
// hot sweeps of the two-phase schedule: with hot_fraction > 0 the kernel stops after the first sweep that accepts
// less than that fraction of its proposals and exports fields (in the warp kernel's register layout), spins and
// the number of completed sweeps, from which sa_warp.cuh's resuming instantiation continues (sa_kernel.cu).
#include "sa_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int T = 16;             // chains per CTA
constexpr int XLD = 36;           // row stride of the transpose buffer
constexpr int GR = 4;             // coupling rows per ring slot: one full / empty hand-shake per GR rows
constexpr int NGS = 4;            // ring slots
constexpr int RB = GR * NGS;      // coupling rows in flight per CTA

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_s)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar_s), "r"(parity) : "memory");
}
// two fp32 FMAs per instruction (Blackwell FFMA2): acc.{x,y} = fma(a.{x,y}, b.{x,y}, acc.{x,y}), each IEEE round-to-nearest
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src, bool valid)
{
    const int sz = valid ? 4 : 0;                           // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t saddr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

struct TileSmem {
    float *ring;        // [RB][ld]
    float *Dbuf;        // [32][32]   diagonal block of the current sub-window (natural order)
    float *Xbuf;        // [T][XLD]   transpose buffer
    float *bounds;      // [4][32][T] acceptance bounds of the current 128-variable window
    float *Cbuf;        // [2][32][T] update coefficients of the record
    uint32_t *spinw;    // [4*WIN][T] spins, one word per (sub-window, chain)
    uint32_t *rec_flip; // [2][T]
    uint32_t *rec_old;  // [2][T]
    uint32_t *rec_meta; // [2][4]     union of the flip masks, number of flips, candidate flag
    uint32_t *work;     // [2][4]     producer work item: union, first row, exit flag
    uint64_t *full;     // [NGS]
    uint64_t *empty;    // [NGS]
    uint64_t *work_full;// [2]
};

__host__ __device__ inline size_t tile_smem_bytes(int ld)
{
    const int WIN = ld / 128;
    size_t b = (size_t)RB * ld * 4 + 32 * 32 * 4 + T * XLD * 4 + 4 * 32 * T * 4 + 2 * 32 * T * 4 + (size_t)4 * WIN * T * 4 +
               2 * T * 4 * 2 + 2 * 4 * 4 * 2 + (size_t)(2 * NGS + 2) * 8;
    return b + 128;
}

__device__ __forceinline__ TileSmem carve(uint8_t *base, int ld)
{
    const int WIN = ld / 128;
    TileSmem s;
    // `base` is the 128-byte aligned dynamic shared memory: every address below is base + a constant, which lets the
    // compiler fold the shared-space addresses of the row loop instead of re-deriving an aligned base in it
    uint8_t *p = base;
    s.ring = reinterpret_cast<float *>(p); p += (size_t)RB * ld * 4;
    s.Dbuf = reinterpret_cast<float *>(p); p += 32 * 32 * 4;
    s.Xbuf = reinterpret_cast<float *>(p); p += T * XLD * 4;
    s.bounds = reinterpret_cast<float *>(p); p += 4 * 32 * T * 4;
    s.Cbuf = reinterpret_cast<float *>(p); p += 2 * 32 * T * 4;
    s.spinw = reinterpret_cast<uint32_t *>(p); p += (size_t)4 * WIN * T * 4;
    s.rec_flip = reinterpret_cast<uint32_t *>(p); p += 2 * T * 4;
    s.rec_old = reinterpret_cast<uint32_t *>(p); p += 2 * T * 4;
    s.rec_meta = reinterpret_cast<uint32_t *>(p); p += 2 * 4 * 4;
    s.work = reinterpret_cast<uint32_t *>(p); p += 2 * 4 * 4;
    s.full = reinterpret_cast<uint64_t *>(p); p += (size_t)NGS * 8;
    s.empty = reinterpret_cast<uint64_t *>(p); p += (size_t)NGS * 8;
    s.work_full = reinterpret_cast<uint64_t *>(p);
    return s;
}

// ---- update: apply the flips of record `par` to every field this thread holds ----------------------
// rows arrive through the ring in ascending order of the flipped variable; `ri` counts rows since launch.
// Fields are kept as float2 pairs (two adjacent sub-windows of a lane) so that one FFMA2 updates two of them.
struct TileAddr {          // shared-space byte addresses, computed once per thread
    uint32_t ring;         // this thread's float4 of window 0 in ring slot 0
    uint32_t slot_stride;  // ld * 4
    uint32_t win_stride;   // W * 512
    uint32_t cbuf;         // Cbuf[0][0][0]
    uint32_t full, empty;  // barrier arrays
    uint32_t flip, old;    // rec_flip[0], rec_old[0]
};

template <int NS>
struct RowRegs {
    float4 r[NS / 4];
    float4 c[T / 4];
};

// rows of a record come in groups of GR per ring slot; `gi` counts groups since launch, `k` rows of this record.
// `rel` = the slot to hand back to the producer after this row has been applied (the last row of its group), else NONE
constexpr uint32_t NONE = 0xffffffffu;
template <int NS>
__device__ __forceinline__ void row_fetch(RowRegs<NS> &R, const TileAddr &A, uint32_t &u, uint32_t &gi, uint32_t &k, uint32_t cb_par,
                                          uint32_t &rel, bool dense)
{
    const int a = __ffs(u) - 1;
    u &= u - 1;
    const uint32_t j = k & (GR - 1);
    const uint32_t slot = gi & (NGS - 1);
    if (j == 0) mbar_wait_s(A.full + slot * 8u, (gi / NGS) & 1u);
#pragma unroll
    for (int jw = 0; jw < NS / 4; ++jw) R.r[jw] = lds128(A.ring + (slot * GR + j) * A.slot_stride + jw * A.win_stride);
    if (dense) {
#pragma unroll
        for (int q = 0; q < T / 4; ++q) R.c[q] = lds128(cb_par + (uint32_t)a * (T * 4u) + q * 16u);
    } else {
        R.c[0].x = __int_as_float(a);
    }
    ++k;
    const bool last = (j == GR - 1) || (u == 0u);
    rel = last ? slot : NONE;
    if (last) { ++gi; k = 0; }
}

template <int NS>
__device__ __forceinline__ void row_apply_dense(float2 (&F2)[NS / 2][T], const RowRegs<NS> &R)
{
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const float4 c4 = R.c[t >> 2];
        const float c = (t & 3) == 0 ? c4.x : ((t & 3) == 1 ? c4.y : ((t & 3) == 2 ? c4.z : c4.w));
        const float2 cc = make_float2(c, c);
#pragma unroll
        for (int jw = 0; jw < NS / 4; ++jw) {
            ffma2(F2[jw * 2 + 0][t], cc, make_float2(R.r[jw].x, R.r[jw].y));
            ffma2(F2[jw * 2 + 1][t], cc, make_float2(R.r[jw].z, R.r[jw].w));
        }
    }
}

template <int NS>
__device__ __forceinline__ void row_apply_sparse(float2 (&F2)[NS / 2][T], const RowRegs<NS> &R, const uint32_t (&fl)[T],
                                                 const uint32_t (&ol)[T])
{
    const int a = __float_as_int(R.c[0].x);
#pragma unroll
    for (int t = 0; t < T; ++t) {
        if ((fl[t] >> a) & 1u) {
            const float c = ((ol[t] >> a) & 1u) ? -2.0f : 2.0f;
            const float2 cc = make_float2(c, c);
#pragma unroll
            for (int jw = 0; jw < NS / 4; ++jw) {
                ffma2(F2[jw * 2 + 0][t], cc, make_float2(R.r[jw].x, R.r[jw].y));
                ffma2(F2[jw * 2 + 1][t], cc, make_float2(R.r[jw].z, R.r[jw].w));
            }
        }
    }
}

template <int NS>
__device__ __forceinline__ void apply_record(float2 (&F2)[NS / 2][T], const TileAddr &A, uint32_t u, uint32_t count, int par,
                                             int lane, uint32_t &gi, uint32_t dense_min)
{
    const uint32_t cb_par = A.cbuf + (uint32_t)par * (32u * T * 4u);
    RowRegs<NS> Ra, Rb;
    uint32_t sa, sb, k = 0;
    if (count >= dense_min) {
        // dense: nearly every chain flipped nearly every variable -- unconditional FMAs with c in {0, +-2};
        // two register buffers: the loads of the next row are in flight while the FMAs of this row issue
        row_fetch<NS>(Ra, A, u, gi, k, cb_par, sa, true);
        while (true) {
            const bool more_b = u != 0u;
            if (more_b) row_fetch<NS>(Rb, A, u, gi, k, cb_par, sb, true);
            row_apply_dense<NS>(F2, Ra);
            if (sa != NONE) {
                __syncwarp();
                if (lane == 0) mbar_arrive_s(A.empty + sa * 8u);
            }
            if (!more_b) break;
            const bool more_a = u != 0u;
            if (more_a) row_fetch<NS>(Ra, A, u, gi, k, cb_par, sa, true);
            row_apply_dense<NS>(F2, Rb);
            if (sb != NONE) {
                __syncwarp();
                if (lane == 0) mbar_arrive_s(A.empty + sb * 8u);
            }
            if (!more_a) break;
        }
    } else {
        // sparse: per chain a warp-uniform test of its flip mask
        uint32_t fl[T], ol[T];
#pragma unroll
        for (int q = 0; q < T / 4; ++q) {
            const uint4 f4 = lds128u(A.flip + (uint32_t)par * (T * 4u) + q * 16u);
            const uint4 o4 = lds128u(A.old + (uint32_t)par * (T * 4u) + q * 16u);
            fl[4 * q] = f4.x; fl[4 * q + 1] = f4.y; fl[4 * q + 2] = f4.z; fl[4 * q + 3] = f4.w;
            ol[4 * q] = o4.x; ol[4 * q + 1] = o4.y; ol[4 * q + 2] = o4.z; ol[4 * q + 3] = o4.w;
        }
        row_fetch<NS>(Ra, A, u, gi, k, cb_par, sa, false);
        while (true) {
            const bool more_b = u != 0u;
            if (more_b) row_fetch<NS>(Rb, A, u, gi, k, cb_par, sb, false);
            row_apply_sparse<NS>(F2, Ra, fl, ol);
            if (sa != NONE) {
                __syncwarp();
                if (lane == 0) mbar_arrive_s(A.empty + sa * 8u);
            }
            if (!more_b) break;
            const bool more_a = u != 0u;
            if (more_a) row_fetch<NS>(Ra, A, u, gi, k, cb_par, sa, false);
            row_apply_sparse<NS>(F2, Rb, fl, ol);
            if (sb != NONE) {
                __syncwarp();
                if (lane == 0) mbar_arrive_s(A.empty + sb * 8u);
            }
            if (!more_a) break;
        }
    }
}

// ---- producer: one bulk copy (TMA, 1-D) per requested coupling row -----------------------------------
__device__ __forceinline__ void tile_producer(const TileSmem &sm, const float *__restrict__ Jp, int ld)
{
    uint32_t kq = 0, ri = 0;
    while (true) {
        mbar_wait(&sm.work_full[kq & 1u], (kq >> 1) & 1u);
        uint32_t u = sm.work[(kq & 1u) * 4 + 0];
        const uint32_t row0 = sm.work[(kq & 1u) * 4 + 1];
        if (sm.work[(kq & 1u) * 4 + 2]) break;
        while (u) {
            // one ring slot = up to GR rows of this record, one expect_tx for all of them
            const uint32_t slot = ri & (NGS - 1);
            const uint32_t cnt = min((uint32_t)__popc(u), (uint32_t)GR);
            mbar_wait(&sm.empty[slot], ((ri / NGS) & 1u) ^ 1u);
            mbar_expect_tx(&sm.full[slot], cnt * (uint32_t)ld * 4u);
            for (uint32_t j = 0; j < cnt; ++j) {
                const int a = __ffs(u) - 1;
                u &= u - 1;
                bulk_g2s(sm.ring + (size_t)(slot * GR + j) * ld, Jp + (size_t)(row0 + a) * (size_t)ld, (uint32_t)ld * 4u, &sm.full[slot]);
            }
            ++ri;
        }
        ++kq;
    }
}

// NS == 8 (n > 1024): 384 threads = two consumer warpgroups (warps 0..W-1 work) + one producer warpgroup
// (warp 8 lane 0 works); registers are moved from the producer to the consumer warpgroups with setmaxnreg
// so that the 128 field registers + working set of a consumer fit without spilling.
// NS == 4 (n <= 1024): (W + 1) warps, the last one is the producer.
template <int NS, int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, 1) sa_tile_kernel(const SaParams p, const int W, const int ctas_per_problem)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n = p.n, ld = p.ld;
    const int S = (n + 31) >> 5;                       // populated 32-variable sub-windows
    const TileSmem sm = carve(smem_raw, ld);
    const int ncons = W * 32;

    const long long q = blockIdx.x / ctas_per_problem;
    const long long r0 = (long long)(blockIdx.x % ctas_per_problem) * T;
    const int nlive = (int)min((long long)T, p.num_reads - r0);
    const float *__restrict__ Jp = p.Jp + (size_t)q * (size_t)n * (size_t)ld;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NGS; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], (uint32_t)W); }
        mbar_init(&sm.work_full[0], 1);
        mbar_init(&sm.work_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (NS == 8) {
        if (warp >= 8) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            if (warp == 8 && lane == 0) tile_producer(sm, Jp, ld);
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        if (warp >= W) return;                          // idle consumer-warpgroup warps (W < 8)
    } else if (warp == W) {
        if (lane == 0) tile_producer(sm, Jp, ld);
        return;
    }

    // ================================= consumer warps =================================
    constexpr int NWIN = NS / 4;
    const float *__restrict__ hq = p.hp + (size_t)q * (size_t)ld;
    const float *__restrict__ Jn = p.Jnat + (size_t)q * (size_t)n * (size_t)p.ldj;
    const float *__restrict__ betas = p.beta + q * p.beta_stride;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const long long cl0 = q * p.num_reads + r0;        // row of chain 0 of this CTA in init / out
    const unsigned long long chain0 = p.chain_offset + (unsigned long long)((p.flags & 2u) ? r0 : cl0);
    const uint32_t dense_pct = ((p.flags >> 8) & 0xffu) ? ((p.flags >> 8) & 0xffu) : 40u;
    const uint32_t dense_min = max(1u, (uint32_t)nlive * 32u * dense_pct / 100u);
    const int tid = threadIdx.x;

    // ---- initial spins ----
    if (p.init != nullptr) {
        for (int st = warp; st < S * T; st += W) {
            const int s = st / T, t = st % T;
            const int v = s * 32 + lane;
            const bool b = (t < nlive && v < n) ? (p.init[(size_t)(cl0 + t) * (size_t)n + v] != 0) : false;
            const unsigned wd = __ballot_sync(FULL, b);
            if (lane == 0) sm.spinw[s * T + t] = wd;
        }
    } else {
        for (int st = tid; st < S * T; st += ncons) {
            const int s = st / T, t = st % T;
            const unsigned long long chain = chain0 + (unsigned long long)t;
            const Philox4 o = philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), 0xFFFFFFFFu, (uint32_t)(s >> 2), k0, k1);
            const int k = s & 3;
            uint32_t wd = k == 0 ? o.x : (k == 1 ? o.y : (k == 2 ? o.z : o.w));
            const int rem = n - s * 32;
            if (rem < 32) wd &= (1u << rem) - 1u;
            sm.spinw[s * T + t] = (t < nlive) ? wd : 0u;
        }
    }
    // ---- fields: F_i = h_i, then F_i = fma(J[j][i], s_j, F_i) for j = 0..n-1 (dense updates with c = s_j) ----
    float2 F2[NS / 2][T];      // F2[jw*2 + h][t] = fields of sub-windows (2h, 2h+1) of window jw, chain t
#pragma unroll
    for (int jw = 0; jw < NWIN; ++jw) {
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(hq + (size_t)(jw * W + warp) * 128 + lane * 4));
#pragma unroll
        for (int t = 0; t < T; ++t) { F2[jw * 2 + 0][t] = make_float2(hv.x, hv.y); F2[jw * 2 + 1][t] = make_float2(hv.z, hv.w); }
    }
    TileAddr A;
    A.ring = smem_u32(sm.ring) + (uint32_t)(warp * 32 + lane) * 16u;
    A.slot_stride = (uint32_t)ld * 4u;
    A.win_stride = (uint32_t)W * 512u;
    A.cbuf = smem_u32(sm.Cbuf);
    A.full = smem_u32(sm.full);
    A.empty = smem_u32(sm.empty);
    A.flip = smem_u32(sm.rec_flip);
    A.old = smem_u32(sm.rec_old);
    uint32_t kq = 0, ri = 0;
    consumer_sync(ncons);
    for (int s = 0; s < S; ++s) {
        const int par = s & 1;
        if (warp == 0) {
            const int rem = n - s * 32;
            const uint32_t live = rem >= 32 ? FULL : ((1u << rem) - 1u);
            // lane = row a of the panel: coefficients s_a(t) = +-1 for all chains
            float *crow = sm.Cbuf + (size_t)(par * 32 + lane) * T;
#pragma unroll
            for (int t = 0; t < T; ++t) crow[t] = ((sm.spinw[s * T + t] >> lane) & 1u) ? 1.0f : -1.0f;
            if (lane == 0) {
                sm.rec_meta[par * 4 + 0] = live;
                sm.rec_meta[par * 4 + 1] = 0xffffffffu;          // dense
                sm.work[(kq & 1u) * 4 + 0] = live;
                sm.work[(kq & 1u) * 4 + 1] = (uint32_t)(s * 32);
                sm.work[(kq & 1u) * 4 + 2] = 0u;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.work_full[kq & 1u]);
        }
        ++kq;
        consumer_sync(ncons);
        apply_record<NS>(F2, A, sm.rec_meta[par * 4 + 0], 0xffffffffu, par, lane, ri, 0u);
    }
    consumer_sync(ncons);

    // ---- annealing ----
    unsigned long long nacc = 0ull;
    uint32_t t_sweep = 0;
    bool handed_over = false;
    const unsigned long long hot_min =
        p.hot_fraction > 0.0f ? (unsigned long long)((double)p.hot_fraction * (double)n * (double)nlive) : 0ull;
    uint32_t pi = 0;                                   // running sub-window counter: parity of the record buffers
    for (int b = 0; b < p.num_betas && !handed_over; ++b) {
        const float beta = __ldg(betas + b);
        const float thr = __fdiv_rn(44.36142f, beta);
        for (int sw = 0; sw < p.sweeps_per_beta && !handed_over; ++sw, ++t_sweep) {
            const unsigned long long nacc_before = nacc;
            int bounds_window = -1;
            for (int s = 0; s < S; ++s, ++pi) {
                const int g = s >> 2, k = s & 3;
                const int owner = g % W, jw_own = g / W;
                const int par = (int)(pi & 1u);
                const bool is_owner = (warp == owner);
                float Fs[T];
                if (is_owner) {
                    // ---- pre-check (lane = variable 32 s + lane) ----
                    const int slot = jw_own * 4 + k;
#pragma unroll
                    for (int j = 0; j < NS; ++j)
                        if (j == slot) {
#pragma unroll
                            for (int t = 0; t < T; ++t) Fs[t] = (j & 1) ? F2[j >> 1][t].y : F2[j >> 1][t].x;
                        }
                    const bool vlive = (s * 32 + lane) < n;
                    bool cand = false;
#pragma unroll
                    for (int q4 = 0; q4 < T / 4; ++q4) {
                        const uint4 w4 = reinterpret_cast<const uint4 *>(sm.spinw + s * T)[q4];
                        const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int t = q4 * 4 + e;
                            const float dE = __fmul_rn(Fs[t], ((ws[e] >> lane) & 1u) ? -2.0f : 2.0f);
                            cand |= (t < nlive) && (dE < thr);
                        }
                    }
                    cand = __any_sync(FULL, cand && vlive);
                    if (lane == 0) sm.rec_meta[par * 4 + 2] = cand ? 1u : 0u;
                    if (cand) {
                        // diagonal block J[32s + a][32s + lane], natural order, for the scan
                        const int col = s * 32 + lane;
#pragma unroll 8
                        for (int a = 0; a < 32; ++a) {
                            const int row = s * 32 + a;
                            const bool ok = (row < n) && (col < n);
                            cp_async4(sm.Dbuf + a * 32 + lane, ok ? (Jn + (size_t)row * (size_t)p.ldj + col) : Jn, ok);
                        }
                    }
                }
                consumer_sync(ncons);                                                          // (A)
                if (sm.rec_meta[par * 4 + 2] == 0u) continue;
                if (bounds_window != g) {
                    // ---- acceptance bounds of window g for all chains: min(thr, -ln(u/2^32)/beta) ----
                    for (int task = tid; task < 32 * T; task += ncons) {
                        const int tc = task % T, ln = task / T;
                        const unsigned long long chain = chain0 + (unsigned long long)tc;
                        const Philox4 o = philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), t_sweep, (uint32_t)(g * 32 + ln), k0, k1);
                        const uint32_t us[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            sm.bounds[(kk * 32 + ln) * T + tc] = fminf(thr, __fdiv_rn(neg_log_u32(us[kk]), beta));
                    }
                    bounds_window = g;
                    consumer_sync(ncons);                                                      // (B)
                }
                if (is_owner) {
                    // ---- scan (lane = chain tc + 16 * half; 16 variables of the sub-window per lane) ----
                    const int tc = lane & 15, hf = lane >> 4;
#pragma unroll
                    for (int t = 0; t < T; ++t) sm.Xbuf[t * XLD + lane] = Fs[t];
                    __syncwarp();
                    float G[16], bnd[16];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float4 x4 = reinterpret_cast<const float4 *>(sm.Xbuf + tc * XLD + hf * 16)[q4];
                        G[4 * q4] = x4.x; G[4 * q4 + 1] = x4.y; G[4 * q4 + 2] = x4.z; G[4 * q4 + 3] = x4.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) bnd[i] = sm.bounds[(k * 32 + hf * 16 + i) * T + tc];
                    const uint32_t old = sm.spinw[s * T + tc];
                    const uint32_t dbuf_s = smem_u32(sm.Dbuf) + (uint32_t)hf * 64u;
                    const uint32_t cbuf_s = smem_u32(sm.Cbuf) + (uint32_t)(par * 32 * T + tc) * 4u;
                    const bool alive = tc < nlive;
                    const int rem = n - s * 32;
                    uint32_t flipm = 0u;
                    cp_async_wait_all();
                    __syncwarp();
                    // branch-free: every step applies c * D[a][.] with c = 0 for chains that keep variable a; the
                    // shared-memory reads do not depend on the decisions (row a+1 is fetched before the vote of row a),
                    // only field -> dE -> compare -> ballot -> c -> fma is a dependent chain
                    const bool h0 = hf == 0;
                    float4 dcur[4], dnxt[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) dcur[q4] = lds128(dbuf_s + (uint32_t)(q4 * 4) * 4u);
#pragma unroll
                    for (int a = 0; a < 32; ++a) {
                        const int ha = a >> 4, i = a & 15;
                        if (a < 31) {
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) dnxt[q4] = lds128(dbuf_s + (uint32_t)((a + 1) * 32 + q4 * 4) * 4u);
                        }
                        const float sg = ((old >> a) & 1u) ? -2.0f : 2.0f;        // variable a has not been visited yet
                        const float dE = __fmul_rn(G[i], sg);
                        const bool acc = ((ha == 0) == h0) & alive & (a < rem) & ((dE <= 0.0f) | (dE < bnd[i]));
                        const uint32_t bal = __ballot_sync(FULL, acc);
                        const bool mine = (bal >> (16 * ha + tc)) & 1u;
                        const float c = mine ? sg : 0.0f;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            G[4 * q4 + 0] = __fmaf_rn(c, dcur[q4].x, G[4 * q4 + 0]);
                            G[4 * q4 + 1] = __fmaf_rn(c, dcur[q4].y, G[4 * q4 + 1]);
                            G[4 * q4 + 2] = __fmaf_rn(c, dcur[q4].z, G[4 * q4 + 2]);
                            G[4 * q4 + 3] = __fmaf_rn(c, dcur[q4].w, G[4 * q4 + 3]);
                        }
                        if (h0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(cbuf_s + (uint32_t)a * (T * 4u)), "f"(c) : "memory");
                        flipm |= (mine ? 1u : 0u) << a;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) dcur[q4] = dnxt[q4];
                    }
                    const uint32_t spin = old ^ flipm;
                    const uint32_t unionm = __reduce_or_sync(FULL, flipm);
                    const uint32_t cnt = __reduce_add_sync(FULL, hf == 0 ? (uint32_t)__popc(flipm) : 0u);
                    if (hf == 0) {
                        sm.spinw[s * T + tc] = spin;
                        sm.rec_flip[par * T + tc] = flipm;
                        sm.rec_old[par * T + tc] = old;
                    }
                    if (lane == 0) {
                        sm.rec_meta[par * 4 + 0] = unionm;
                        sm.rec_meta[par * 4 + 1] = cnt;
                        if (unionm) {
                            sm.work[(kq & 1u) * 4 + 0] = unionm;
                            sm.work[(kq & 1u) * 4 + 1] = (uint32_t)(s * 32);
                            sm.work[(kq & 1u) * 4 + 2] = 0u;
                        }
                    }
                    __syncwarp();
                    if (lane == 0 && unionm) mbar_arrive(&sm.work_full[kq & 1u]);
                }
                consumer_sync(ncons);                                                          // (C)
                if (sm.rec_meta[par * 4 + 0] != 0u) {
                    nacc += sm.rec_meta[par * 4 + 1];
                    ++kq;
                    apply_record<NS>(F2, A, sm.rec_meta[par * 4 + 0], sm.rec_meta[par * 4 + 1], par, lane, ri, dense_min);
                }
            }
            // two-phase schedule: once a sweep accepts less than hot_fraction of its proposals the chains are cheaper to
            // advance one warp each (rows no longer shared by most chains); every consumer thread sees the same counts
            if (hot_min > 0ull && nacc - nacc_before < hot_min) handed_over = true;
        }
    }
    consumer_sync(ncons);
    if (p.fields != nullptr) {
        // fields in the warp kernel's layout: chain-major, float4 (sub-windows 0..3) per lane and 128-variable window
#pragma unroll
        for (int jw = 0; jw < NWIN; ++jw) {
            const size_t off = (size_t)(jw * W + warp) * 128 + (size_t)lane * 4;
#pragma unroll
            for (int t = 0; t < T; ++t)
                if (t < nlive)
                    *reinterpret_cast<float4 *>(p.fields + (size_t)(cl0 + t) * (size_t)ld + off) =
                        make_float4(F2[jw * 2][t].x, F2[jw * 2][t].y, F2[jw * 2 + 1][t].x, F2[jw * 2 + 1][t].y);
        }
        if (tid == 0) p.sweeps_done[blockIdx.x] = t_sweep;            // completed sweeps of this tile
    }

    // ---- write-back: states in natural variable order, 0/1; stop the producer ----
    if (tid == 0) {
        sm.work[(kq & 1u) * 4 + 0] = 0u;
        sm.work[(kq & 1u) * 4 + 2] = 1u;
        mbar_arrive(&sm.work_full[kq & 1u]);
    }
    for (int t = warp; t < nlive; t += W) {
        int8_t *o = p.out + (size_t)(cl0 + t) * (size_t)n;
        for (int s = 0; s < S; ++s) {
            const int v = s * 32 + lane;
            if (v < n) o[v] = (int8_t)((sm.spinw[s * T + t] >> lane) & 1u);
        }
    }
    if (p.counters != nullptr && tid == 0) {
        atomicAdd(p.counters + 0, nacc);
        atomicAdd(p.counters + 1, (unsigned long long)n * (unsigned long long)t_sweep * (unsigned long long)nlive);
    }
}

template <int NS, int NTHREADS>
int launch_tile(const SaParams &p, int W, cudaStream_t st)
{
    auto kern = sa_tile_kernel<NS, NTHREADS>;
    const long long cpp = (p.num_reads + T - 1) / T;
    const long long blocks = cpp * p.batch_q;
    if (blocks > 0x7fffffffLL) {
        qbm_set_error("qbm_sa_sample: too many chains for one launch (%lld)", p.total_chains);
        return QBM_EUNSUPPORTED;
    }
    const size_t smem = tile_smem_bytes(p.ld);
    static size_t attr_set[2] = {0, 0};
    size_t &cur = attr_set[NS == 8 ? 1 : 0];
    if (smem > cur) {
        QBM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    kern<<<(unsigned)blocks, NS == 8 ? NTHREADS : (W + 1) * 32, smem, st>>>(p, W, (int)cpp);
    QBM_LAUNCH_OK("sa_tile_kernel");
    return QBM_OK;
}

}  // namespace

bool sa_tile_supported(int n) { return n >= 1 && n <= QBM_SA_MAX_N; }

// row length of the permuted coupling matrix the tile kernel reads: whole 128-variable windows, one
// (n <= 1024) or two (n > 1024) per consumer warp
int sa_tile_ld(int n)
{
    const int win = (n + 127) / 128;
    if (win <= 8) return win * 128;
    return ((win + 1) / 2) * 2 * 128;
}

int sa_tile_launch(const SaParams &p, cudaStream_t st)
{
    const int win = p.ld / 128;
    if (win <= 8) return launch_tile<4, 9 * 32>(p, win, st);               // n <= 1024: W = win consumer warps x 4 columns
    return launch_tile<8, 384>(p, win / 2, st);                            // n >  1024: W = win / 2 consumer warps x 8 columns
}
