// K1b: simulated-annealing QUBO sampler for sm_100a -- a register tile of 16 chains per CTA, software-pipelined.
//
// Same trajectory as sa_kernel.cu (oracle/replay_sa.c), different execution.  In the hot part of
// neal's legacy schedule nearly every proposal is accepted (acceptance > 0.9 for the first ~10 % of the sweeps, where
// ~97 % of all flips happen), so every chain needs nearly every coupling row in every sweep.  One warp per chain streams
// a row per flip per chain through L1 (the 128 B/clk/SM pipe is the wall).  Here the 16 chains of a CTA advance in
// lock-step over 32-variable sub-windows, each coupling row is fetched ONCE for all of them, and the inherently serial
// part (deciding the 32 proposals of a sub-window in sweep order) runs one sub-window AHEAD of the field updates on a
// warp of its own, so the FMA pipe never waits for it.  Warp roles of a CTA (384 threads, one CTA per SM):
//
//   appliers  warps 0..W-1 (two warpgroups, registers raised with setmaxnreg): thread (warp w, lane l) holds NS columns
//             x 16 chains of local fields in registers (column = variable (jw*W + w)*128 + k*32 + l).  They consume
//             RECORDS in order: for every flipped variable a of a record (sweep order) the row J[a] is read once from
//             the ring and applied to all chains that flipped a: F[.][t] = fma(c_t, J[a][.], F[.][t]) (FFMA2).  Before
//             applying record r the owner of sub-window r+1 exports that sub-window's 32 x 16 fields to shared memory.
//   scanners  warps 9 and 11, TS = 8 chains each, lane = chain x part (a part = 8 variables of the sub-window): take the
//             exported fields of sub-window r+1 (they contain every flip up to record r-1), apply record r to them
//             themselves from the 32x32 block J[rows of r][columns of r+1] -- the same FMAs in the same order as the
//             appliers will -- then visit the 32 variables in sweep order (a flip updates the 32 fields from the
//             diagonal block) and publish their half of record r+1: per chain a flip mask, the old spins and a
//             coefficient matrix c[a][t] in {0, +-2}.  Spins live with the scanners only.  The scan is a chain of
//             dependent instructions (one warp retires one every ~5 clocks); the chains of a tile are independent, so
//             two warps on two schedulers halve the time per sub-window, which is what bounds the kernel.
//   bounds    warp 10: Philox + -ln(u)/beta acceptance bounds of the next 128-variable window, double-buffered
//   producer  warp 8, one lane: one cp.async.bulk (TMA, 1-D) per flipped coupling row into a ring of 16 rows, one
//             mbarrier full/empty hand-shake per ring slot of 4 rows; rows come from L2 (the matrix is read once per
//             sweep and SM, not once per flip and chain)
//
// All hand-offs (records, field exports, bounds, ring slots) are mbarrier full/empty pairs; there is no CTA-wide
// barrier inside the sweep loop.  Every field element receives exactly the FMA sequence of the sequential rule, in the
// same order, so the final states are bit-identical to the replay oracle (tests/test_gpu_sa.py).
//
// Use: as a whole-schedule sampler behind qbm_sa_sample flag bit 4, and -- by default for large n -- for the hot sweeps
// of the two-phase schedule: with hot_fraction > 0 the kernel stops after the first sweep that accepts less than that
// fraction of its proposals and exports fields (in the warp kernel's register layout), spins and the number of
// completed sweeps, from which sa_warp.cuh's resuming instantiation continues (sa_kernel.cu).
#include "sa_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// -DQBM_TILE_PROF: clocks the warp roles spend waiting on each hand-off (tools/probe_tile_prof.py); never in the shipped build
#ifdef QBM_TILE_PROF
__device__ unsigned long long g_tile_prof[32];
#define PROF_VAR(x) uint32_t x = 0u
#define PROF_WAIT(x, stmt) do { const uint32_t _t0 = (uint32_t)clock(); stmt; x += (uint32_t)clock() - _t0; } while (0)
#define PROF_PUT(i, x) atomicAdd(&g_tile_prof[i], (unsigned long long)(x))
#else
#define PROF_VAR(x)
#define PROF_WAIT(x, stmt) do { stmt; } while (0)
#define PROF_PUT(i, x)
#endif
constexpr int T = 16;             // chains per CTA
#ifndef QBM_TILE_GR
#define QBM_TILE_GR 4
#define QBM_TILE_NGS 4
#endif
constexpr int GR = QBM_TILE_GR;   // coupling rows per ring slot: one full / empty hand-shake per GR rows
constexpr int NGS = QBM_TILE_NGS; // ring slots (a power of two)
constexpr int RB = GR * NGS;      // coupling rows in flight per CTA
constexpr int NREC = 4;           // record buffers (the scanner runs at most one sub-window ahead of the slowest applier)
constexpr int FXLD = 32 * T + 32; // floats per field-export buffer: [column][chain], the upper 16 columns shifted by 16
// Warp slots of a CTA: AW applier slots (AW = 8: one CTA per SM; AW = 4, problems of at most four 128-variable windows: two
// CTAs per SM -- there the scanners, whose work per sub-window does not depend on n, bound a CTA, and a second one doubles
// the SM's rate), then producer, scanner 0, bounds, scanner 1.  setmaxnreg moves registers between the warpgroups:
//   AW = 8: launched with 168, appliers 208, helpers 88  (2 * 208 + 88 = 504 = 3 * 168)
//   AW = 4: launched with 128, appliers 160, helpers 96  (160 + 96 = 2 * 128; two CTAs = the whole register file)
template <int AW> struct Slots {
    static constexpr int nthreads = (AW + 4) * 32, min_ctas = AW == 8 ? 1 : 2;
    static constexpr int producer = AW, scanner0 = AW + 1, bounds = AW + 2, scanner1 = AW + 3;
    static constexpr int regs_applier = AW == 8 ? 208 : 160, regs_helper = AW == 8 ? 88 : 96;
};
constexpr int NSCAN = 2;          // scanner warps (warps 9 and 11: one per scheduler that holds no other helper warp)
constexpr int TS = T / NSCAN;     // chains per scanner warp; lane = chain (TS) x part (32 / TS), a part = TS variables
constexpr int NPART = 32 / TS;
static_assert(NSCAN == 1 || NSCAN == 2, "rowmask words are written as 32 / NSCAN-bit halves");
template <int AW> __device__ __forceinline__ int scanner_of_warp(int warp)
{
    return warp == Slots<AW>::scanner0 ? 0 : (NSCAN == 2 && warp == Slots<AW>::scanner1 ? 1 : -1);
}
// a rowmask word holds, per scanner, TS flip bits then TS old-spin bits: chain t flipped <=> bit flip_bit(t), its old spin
// is bit flip_bit(t) + TS
__host__ __device__ constexpr int flip_bit(int t) { return (t / TS) * 2 * TS + (t % TS); }
// record meta words
constexpr int META = 8, M_UNION = 0, M_COUNT = 2, M_ROW0 = 4, M_EXIT = 5;
constexpr uint32_t DENSE_MARK = 0x0fffffffu;

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
// wait that gives up when `*flag` becomes non-zero (the bounds warp runs ahead of a scanner that may stop early);
// try_wait suspends for a bounded time, so the flag is polled between attempts.  Returns false when it gave up.
__device__ __forceinline__ bool mbar_wait_or_flag(uint64_t *bar, uint32_t parity, volatile uint32_t *flag)
{
    const uint32_t a = smem_u32(bar);
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
        if (*flag != 0u) return false;
    }
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
// The accumulators are kept as packed 64-bit values for their whole life (fields of two adjacent sub-windows of a lane):
// with float2 accumulators the pack / unpack around every FFMA2 of the row loop survives as register moves.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ void ffma2(f32x2 &acc, const f32x2 a, const f32x2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
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
__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// appliers + scanner (the warps that touch the spin words): start-up and write-back rendezvous only
__device__ __forceinline__ void tile_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

struct TileSmem {
    float *ring;         // [RB][ld]
    float *Dbuf;         // [NSCAN][2][64][32]  rows 0..31: J[rows of the previous sub-window][columns of this one], 32..63: diagonal block
    float *Fx;           // [2][FXLD]    exported fields of the sub-window the scanner visits next
    float *bounds;       // [2][4][32][T] acceptance bounds of a 128-variable window
    float *Cbuf;         // [NREC][32][T]   update coefficients of a record (the FFMA2 takes them as a broadcast scalar operand)
    uint32_t *spinw;     // [MAXSUB][T]  spins, one word per (sub-window, chain)
    uint32_t *rowmask;   // [NREC][32]   per flipped variable: chains that flipped it and their old spins (flip_bit)
    uint32_t *rec_meta;  // [NREC][META] per scanner: union of the flip masks, number of flips; first row, exit flag
    uint32_t *ctl;       // [12]         [0] exit flag for the bounds warp, [1..] flips of a sweep per scanner (two parities)
    uint64_t *full;      // [NGS]  ring slot filled (producer -> appliers)
    uint64_t *empty;     // [NGS]  ring slot released (appliers -> producer)
    uint64_t *rec_full;  // [NREC] record published (scanner -> appliers, producer)
    uint64_t *rec_empty; // [NREC] record consumed (appliers, producer -> scanner)
    uint64_t *fx_full;   // [2]    fields exported (owning applier -> scanner)
    uint64_t *fx_empty;  // [2]
    uint64_t *bnd_full;  // [2]    bounds drawn (bounds warp -> scanner)
    uint64_t *bnd_empty; // [2]
};

// layout: every small buffer at a compile-time offset from the (128-byte aligned) dynamic shared memory base, the ring --
// the only part whose size depends on n -- last; all shared addresses of the control path fold to base + immediate
constexpr int MAXSUB = QBM_SA_MAX_N / 32;             // sub-windows of the largest problem
constexpr size_t OFF_DBUF = 0;
constexpr size_t OFF_FX = OFF_DBUF + (size_t)NSCAN * 2 * 64 * 32 * 4;
constexpr size_t OFF_BOUNDS = OFF_FX + 2 * FXLD * 4;
constexpr size_t OFF_CBUF = OFF_BOUNDS + 2 * 4 * 32 * T * 4;
constexpr size_t OFF_SPINW = OFF_CBUF + NREC * 32 * T * 4;
constexpr size_t OFF_ROWMASK = OFF_SPINW + (size_t)MAXSUB * T * 4;
constexpr size_t OFF_META = OFF_ROWMASK + NREC * 32 * 4;
constexpr size_t OFF_CTL = OFF_META + NREC * META * 4;
constexpr size_t OFF_BARS = OFF_CTL + 12 * 4;
static_assert(31 * T + (NPART - 1) * TS + T <= FXLD, "field-export buffer: last column + its part shift + 16 chains");
static_assert(OFF_BARS % 8 == 0 && OFF_FX % 16 == 0 && OFF_CBUF % 16 == 0 && OFF_ROWMASK % 4 == 0, "alignment of the carve-up");
constexpr size_t OFF_RING = (OFF_BARS + (size_t)(2 * NGS + 2 * NREC + 8) * 8 + 127) / 128 * 128;

__host__ __device__ inline size_t tile_smem_bytes(int ld) { return OFF_RING + (size_t)RB * ld * 4; }

__device__ __forceinline__ TileSmem carve(uint8_t *base)
{
    TileSmem s;
    s.Dbuf = reinterpret_cast<float *>(base + OFF_DBUF);
    s.Fx = reinterpret_cast<float *>(base + OFF_FX);
    s.bounds = reinterpret_cast<float *>(base + OFF_BOUNDS);
    s.Cbuf = reinterpret_cast<float *>(base + OFF_CBUF);
    s.spinw = reinterpret_cast<uint32_t *>(base + OFF_SPINW);
    s.rowmask = reinterpret_cast<uint32_t *>(base + OFF_ROWMASK);
    s.rec_meta = reinterpret_cast<uint32_t *>(base + OFF_META);
    s.ctl = reinterpret_cast<uint32_t *>(base + OFF_CTL);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + OFF_BARS);
    s.full = bars; s.empty = bars + NGS;
    s.rec_full = bars + 2 * NGS; s.rec_empty = bars + 2 * NGS + NREC;
    s.fx_full = bars + 2 * NGS + 2 * NREC; s.fx_empty = s.fx_full + 2;
    s.bnd_full = s.fx_full + 4; s.bnd_empty = s.fx_full + 6;
    s.ring = reinterpret_cast<float *>(base + OFF_RING);
    return s;
}

// ---- appliers: apply the flips of a record to every field this thread holds --------------------------
// rows arrive through the ring in ascending order of the flipped variable.  Fields are kept as float2 pairs (two
// adjacent sub-windows of a lane) so that one FFMA2 updates two of them.
struct TileAddr {          // shared-space byte addresses, computed once per thread
    uint32_t ring;         // this thread's float4 of window 0 in ring slot 0
    uint32_t slot_stride;  // ld * 4
    uint32_t win_stride;   // W * 512
    uint32_t cbuf;         // Cbuf[0][0][0]
    uint32_t full, empty;  // ring barrier arrays
    uint32_t rowmask;      // rowmask[0][0]
};

template <int NS>
struct RowRegs {
    float4 r[NS / 4];
    uint32_t c;            // dense: shared address of the row's coefficient pairs; sparse: the row's mask word
};

// one coupling row of this thread's columns, from ring position `raddr`
template <int NS>
__device__ __forceinline__ void row_load(RowRegs<NS> &R, const TileAddr &A, uint32_t raddr)
{
#pragma unroll
    for (int jw = 0; jw < NS / 4; ++jw) R.r[jw] = lds128(raddr + jw * A.win_stride);
}

template <int NS>
__device__ __forceinline__ void row_apply_dense(f32x2 (&F2)[NS / 2][T], const RowRegs<NS> &R)
{
    // four chains per 128-bit load (warp-uniform address: a broadcast); (c, c) is a scalar operand of the FFMA2
#pragma unroll
    for (int t4 = 0; t4 < T / 4; ++t4) {
        const float4 c4 = lds128(R.c + (uint32_t)t4 * 16u);
        const f32x2 c0 = pack2(c4.x, c4.x), c1 = pack2(c4.y, c4.y), c2 = pack2(c4.z, c4.z), c3 = pack2(c4.w, c4.w);
#pragma unroll
        for (int jw = 0; jw < NS / 4; ++jw) {
            const f32x2 r01 = pack2(R.r[jw].x, R.r[jw].y), r23 = pack2(R.r[jw].z, R.r[jw].w);
            ffma2(F2[jw * 2 + 0][4 * t4 + 0], c0, r01);
            ffma2(F2[jw * 2 + 1][4 * t4 + 0], c0, r23);
            ffma2(F2[jw * 2 + 0][4 * t4 + 1], c1, r01);
            ffma2(F2[jw * 2 + 1][4 * t4 + 1], c1, r23);
            ffma2(F2[jw * 2 + 0][4 * t4 + 2], c2, r01);
            ffma2(F2[jw * 2 + 1][4 * t4 + 2], c2, r23);
            ffma2(F2[jw * 2 + 0][4 * t4 + 3], c3, r01);
            ffma2(F2[jw * 2 + 1][4 * t4 + 3], c3, r23);
        }
    }
}

template <int NS>
__device__ __forceinline__ void row_apply_sparse(f32x2 (&F2)[NS / 2][T], const RowRegs<NS> &R)
{
    const uint32_t m = R.c;              // the row's mask word: chains that flipped this variable and their old spins
#pragma unroll
    for (int t = 0; t < T; ++t) {
        if ((m >> flip_bit(t)) & 1u) {   // warp-uniform
            const float c = ((m >> (flip_bit(t) + TS)) & 1u) ? -2.0f : 2.0f;
            const f32x2 cc = pack2(c, c);
#pragma unroll
            for (int jw = 0; jw < NS / 4; ++jw) {
                ffma2(F2[jw * 2 + 0][t], cc, pack2(R.r[jw].x, R.r[jw].y));
                ffma2(F2[jw * 2 + 1][t], cc, pack2(R.r[jw].z, R.r[jw].w));
            }
        }
    }
}

// The rows of a record arrive through the ring in ascending order of the flipped variable, in groups of up to GR rows per
// ring slot (`gi` counts groups since launch, exactly as the producer does): one full / empty hand-shake per group, the
// rows of a group at static offsets -- the per-row bookkeeping is the next set bit of `u` and two address adds.
template <int NS>
__device__ __forceinline__ void apply_record(f32x2 (&F2)[NS / 2][T], const TileAddr &A, uint32_t u, uint32_t count, uint32_t rb,
                                             int lane, uint32_t &gi, uint32_t dense_min, uint32_t &w_ring)
{
    const uint32_t cb_rec = A.cbuf + rb * (32u * T * 4u);
    const uint32_t rm_rec = A.rowmask + rb * (32u * 4u);
    // The dense and the sparse loop are kept apart: one loop with both bodies makes ptxas reconcile the 128 field
    // registers with moves
    if (count >= dense_min) {
        // dense: nearly every chain flipped nearly every variable -- unconditional FMAs with c in {0, +-2}
        while (u != 0u) {
            const uint32_t slot = gi & (NGS - 1);
            const uint32_t cnt = min((uint32_t)__popc(u), (uint32_t)GR);
            const uint32_t base = A.ring + slot * (GR * A.slot_stride);
            PROF_WAIT(w_ring, mbar_wait_s(A.full + slot * 8u, (gi / NGS) & 1u));
            if (cnt == (uint32_t)GR) {
                // a full group (nearly all of them while the sweeps are hot): four rows, straight-line
#pragma unroll
                for (int j = 0; j < GR; ++j) {
                    const int a = __ffs(u) - 1;
                    u &= u - 1;
                    RowRegs<NS> R;
                    row_load<NS>(R, A, base + j * A.slot_stride);
                    R.c = cb_rec + (uint32_t)a * (T * 4u);
                    row_apply_dense<NS>(F2, R);
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < cnt; ++j) {
                    const int a = __ffs(u) - 1;
                    u &= u - 1;
                    RowRegs<NS> R;
                    row_load<NS>(R, A, base + j * A.slot_stride);
                    R.c = cb_rec + (uint32_t)a * (T * 4u);
                    row_apply_dense<NS>(F2, R);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_s(A.empty + slot * 8u);
            ++gi;
        }
    } else {
        // sparse: per chain a warp-uniform test of the row's mask word
        while (u != 0u) {
            const uint32_t slot = gi & (NGS - 1);
            const uint32_t cnt = min((uint32_t)__popc(u), (uint32_t)GR);
            const uint32_t base = A.ring + slot * (GR * A.slot_stride);
            PROF_WAIT(w_ring, mbar_wait_s(A.full + slot * 8u, (gi / NGS) & 1u));
#pragma unroll 1
            for (uint32_t j = 0; j < cnt; ++j) {
                const int a = __ffs(u) - 1;
                u &= u - 1;
                RowRegs<NS> R;
                row_load<NS>(R, A, base + j * A.slot_stride);
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(R.c) : "r"(rm_rec + (uint32_t)a * 4u));
                row_apply_sparse<NS>(F2, R);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_s(A.empty + slot * 8u);
            ++gi;
        }
    }
}

// ---- producer: one bulk copy (TMA, 1-D) per flipped coupling row of every record ---------------------
__device__ __forceinline__ void tile_producer(const TileSmem &sm, const float *__restrict__ Jp, int ld)
{
    uint32_t r = 0, ri = 0;
    PROF_VAR(w_rec); PROF_VAR(w_empty);
#ifdef QBM_TILE_PROF
    const uint32_t t_start = (uint32_t)clock();
#endif
    while (true) {
        const uint32_t rb = r & (NREC - 1);
        PROF_WAIT(w_rec, mbar_wait(&sm.rec_full[rb], (r / NREC) & 1u));
        uint32_t u = sm.rec_meta[rb * META + M_UNION];
        if (NSCAN == 2) u |= sm.rec_meta[rb * META + M_UNION + 1];
        const uint32_t row0 = sm.rec_meta[rb * META + M_ROW0];
        const uint32_t ex = sm.rec_meta[rb * META + M_EXIT];
        mbar_arrive(&sm.rec_empty[rb]);
        if (ex) break;
        while (u) {
            // one ring slot = up to GR rows of this record, one expect_tx for all of them
            const uint32_t slot = ri & (NGS - 1);
            const uint32_t cnt = min((uint32_t)__popc(u), (uint32_t)GR);
            PROF_WAIT(w_empty, mbar_wait(&sm.empty[slot], ((ri / NGS) & 1u) ^ 1u));
            mbar_expect_tx(&sm.full[slot], cnt * (uint32_t)ld * 4u);
            for (uint32_t j = 0; j < cnt; ++j) {
                const int a = __ffs(u) - 1;
                u &= u - 1;
                bulk_g2s(sm.ring + (size_t)(slot * GR + j) * ld, Jp + (size_t)(row0 + a) * (size_t)ld, (uint32_t)ld * 4u, &sm.full[slot]);
            }
            ++ri;
        }
        ++r;
    }
#ifdef QBM_TILE_PROF
    PROF_PUT(16, (uint32_t)clock() - t_start); PROF_PUT(17, w_rec); PROF_PUT(18, w_empty);
#endif
}

// ---- bounds warp: min(thr, -ln(u/2^32)/beta) for every (variable, chain) of the next window -------------
__device__ __forceinline__ void tile_bounds(const TileSmem &sm, const SaParams &p, const float *__restrict__ betas,
                                            unsigned long long chain0, int lane)
{
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const int nwin = (p.n + 127) >> 7;
    volatile uint32_t *ctl = sm.ctl;
    uint32_t wi = 0, t_sweep = 0;
    for (int b = 0; b < p.num_betas; ++b) {
        const float beta = __ldg(betas + b);
        const float thr = __fdiv_rn(44.36142f, beta);
        for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++t_sweep) {
            for (int g = 0; g < nwin; ++g, ++wi) {
                const uint32_t buf = wi & 1u;
                if (!mbar_wait_or_flag(&sm.bnd_empty[buf], ((wi >> 1) & 1u) ^ 1u, ctl)) return;    // the scanner has stopped
                float *bo = sm.bounds + buf * (4 * 32 * T);
#pragma unroll 2
                for (int task = lane; task < 32 * T; task += 32) {
                    const int tc = task % T, ln = task / T;
                    const unsigned long long chain = chain0 + (unsigned long long)tc;
                    const Philox4 o = philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), t_sweep, (uint32_t)(g * 32 + ln), k0, k1);
                    const uint32_t us[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) bo[(kk * 32 + ln) * T + tc] = fminf(thr, __fdiv_rn(neg_log_u32(us[kk]), beta));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.bnd_full[buf]);
            }
        }
    }
}

// ---- scanner: the coupling blocks of a sub-window, fetched one sub-window ahead -------------------------
// rows 0..31 of the buffer: J[32 sp + j][32 sc + .] (the previous sub-window's rows), rows 32..63: the diagonal block
__device__ __forceinline__ void scan_prefetch(float *D, const float *__restrict__ Jn, int ldj, int n, int sp, int sc, int lane, bool vec16)
{
    if (vec16) {
        const int chunk = lane & 7, rsub = lane >> 3;
        const int col = sc * 32 + chunk * 4;
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
            const int j = it * 4 + rsub;
            const int row = (j < 32 ? sp * 32 + j : sc * 32 + j - 32);
            const int bytes = (row < n) ? max(0, min(16, (n - col) * 4)) : 0;
            cp_async16(D + j * 32 + chunk * 4, bytes > 0 ? (Jn + (size_t)row * (size_t)ldj + col) : Jn, bytes);
        }
    } else {
        const int col = sc * 32 + lane;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
            const int row = (j < 32 ? sp * 32 + j : sc * 32 + j - 32);
            const bool ok = (row < n) && (col < n);
            cp_async4(D + j * 32 + lane, ok ? (Jn + (size_t)row * (size_t)ldj + col) : Jn, ok);
        }
    }
    cp_async_commit();
}

__device__ __forceinline__ void scanners_sync() { asm volatile("bar.sync 2, %0;" ::"n"(NSCAN * 32) : "memory"); }

// scanner `wsc` decides the proposals of chains wsc * TS .. wsc * TS + TS - 1; lane = chain tc + TS * part, a part = TS
// consecutive variables of the 32-variable sub-window
__device__ __forceinline__ void tile_scanner(const TileSmem &sm, const SaParams &p, const float *__restrict__ Jn, const float *__restrict__ betas,
                             int W, int nlive, int lane, int wsc)
{
    constexpr int Q4 = TS / 4;                        // 128-bit loads per row of this lane's part
    constexpr uint32_t TSMASK = TS == 32 ? 0xffffffffu : ((1u << TS) - 1u);
    const int n = p.n;
    const int S = (n + 31) >> 5;
    const int tc = lane % TS, part = lane / TS;
    const int t = wsc * TS + tc;                      // this lane's chain of the tile
    const bool alive = t < nlive, p0 = part == 0;
    const int nlive_w = max(0, min(TS, nlive - wsc * TS));
    const bool vec16 = ((p.ldj & 3) == 0) && ((reinterpret_cast<uintptr_t>(Jn) & 15u) == 0);
    float *const Dw = sm.Dbuf + wsc * (2 * 64 * 32);  // this scanner's copy of the coupling blocks
    uint32_t r = 0;                                   // record counter
    // what the previous record did to this lane's chain: rows, signs and magnitude of its coefficients
    uint32_t prev_mask = 0u, prev_neg = 0u, prev_union = 0u;
    float prev_mag = 1.0f;

    // ---- records 0..S-1: the initial fields, F_i = h_i + sum_j J[j][i] s_j as dense updates with c = s_j ----
    for (int s = 0; s < S; ++s, ++r) {
        const uint32_t rb = r & (NREC - 1);
        mbar_wait(&sm.rec_empty[rb], ((r / NREC) & 1u) ^ 1u);
        const int rem = n - s * 32;
        const uint32_t live = rem >= 32 ? FULL : ((1u << rem) - 1u);
        float *crow = sm.Cbuf + (size_t)(rb * 32 + lane) * T + wsc * TS;     // lane = row a of the panel
#pragma unroll
        for (int j = 0; j < TS; ++j) crow[j] = ((sm.spinw[s * T + wsc * TS + j] >> lane) & 1u) ? 1.0f : -1.0f;
        if (lane == 0) {
            sm.rec_meta[rb * META + M_UNION + wsc] = live;
            sm.rec_meta[rb * META + M_COUNT + wsc] = DENSE_MARK;                   // dense
            sm.rec_meta[rb * META + M_ROW0] = (uint32_t)(s * 32);                  // (every scanner writes the same value)
            sm.rec_meta[rb * META + M_EXIT] = 0u;
        }
        prev_union = live; prev_mask = live; prev_neg = ~sm.spinw[s * T + t]; prev_mag = 1.0f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.rec_full[rb]);
    }

    // ---- annealing ----
    PROF_VAR(w_fx); PROF_VAR(w_bnd); PROF_VAR(w_rece); PROF_VAR(w_cp); PROF_VAR(c_catch); PROF_VAR(c_scan);
#ifdef QBM_TILE_PROF
    const uint32_t t_start = (uint32_t)clock();
#endif
    unsigned long long nacc = 0ull;
    uint32_t t_sweep = 0, e = 0, wi = 0;
    const unsigned long long hot_min =
        p.hot_fraction > 0.0f ? (unsigned long long)((double)p.hot_fraction * (double)n * (double)nlive) : 0ull;
    bool handed_over = false;
    if (p.num_betas > 0) scan_prefetch(Dw, Jn, p.ldj, n, S - 1, 0, lane, vec16);
    for (int b = 0; b < p.num_betas && !handed_over; ++b) {
        const float beta = __ldg(betas + b);
        const float thr = __fdiv_rn(44.36142f, beta);
        for (int sw = 0; sw < p.sweeps_per_beta && !handed_over; ++sw, ++t_sweep) {
            const unsigned long long nacc_before = nacc;
            for (int s = 0; s < S; ++s, ++r, ++e) {
                const int k = s & 3;
                const uint32_t rb = r & (NREC - 1), eb = e & 1u;
                const float *D = Dw + eb * (64 * 32);
                // blocks of this sub-window have landed; fetch the next one's (the last prefetch of a launch is unused)
                PROF_WAIT(w_cp, cp_async_wait_all());
                __syncwarp();
                scan_prefetch(Dw + (eb ^ 1u) * (64 * 32), Jn, p.ldj, n, s, (s + 1 == S) ? 0 : s + 1, lane, vec16);
                // exported fields: every flip up to record r-2 applied
                float G[TS];
                PROF_WAIT(w_fx, mbar_wait(&sm.fx_full[eb], (e >> 1) & 1u));
                {
                    const float *fx = sm.Fx + eb * FXLD + part * (TS * T + TS) + t;
#pragma unroll
                    for (int i = 0; i < TS; ++i) G[i] = fx[i * T];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.fx_empty[eb]);
                // record r-1 applied here, ahead of the appliers: same rows, same order, same coefficients
#ifdef QBM_TILE_PROF
                const uint32_t t_catch0 = (uint32_t)clock();
#endif
                {
                    const uint32_t off_s = smem_u32(D) + (uint32_t)part * (TS * 4u);
#pragma unroll 4
                    for (int a = 0; a < 32; ++a) {
                        if (!((prev_union >> a) & 1u)) continue;                              // warp-uniform
                        const float c = ((prev_mask >> a) & 1u) ? (((prev_neg >> a) & 1u) ? -prev_mag : prev_mag) : 0.0f;
#pragma unroll
                        for (int q4 = 0; q4 < Q4; ++q4) {
                            const float4 d = lds128(off_s + (uint32_t)(a * 32 + q4 * 4) * 4u);
                            G[4 * q4 + 0] = __fmaf_rn(c, d.x, G[4 * q4 + 0]);
                            G[4 * q4 + 1] = __fmaf_rn(c, d.y, G[4 * q4 + 1]);
                            G[4 * q4 + 2] = __fmaf_rn(c, d.z, G[4 * q4 + 2]);
                            G[4 * q4 + 3] = __fmaf_rn(c, d.w, G[4 * q4 + 3]);
                        }
                    }
                }
#ifdef QBM_TILE_PROF
                c_catch += (uint32_t)clock() - t_catch0;
#endif
                const uint32_t old = sm.spinw[s * T + t];
                const int rem = n - s * 32;
                // ---- pre-check: can any chain accept anything here (dE < 44.36142/beta)? ----
                bool cand = false;
#pragma unroll
                for (int i = 0; i < TS; ++i) {
                    const int a = part * TS + i;
                    const float dE = __fmul_rn(G[i], ((old >> a) & 1u) ? -2.0f : 2.0f);
                    cand |= (a < rem) && (dE < thr);
                }
                cand = __any_sync(FULL, cand && alive);
                if (k == 0) PROF_WAIT(w_bnd, mbar_wait(&sm.bnd_full[wi & 1u], (wi >> 1) & 1u));
                PROF_WAIT(w_rece, mbar_wait(&sm.rec_empty[rb], ((r / NREC) & 1u) ^ 1u));
#ifdef QBM_TILE_PROF
                const uint32_t t_scan0 = (uint32_t)clock();
#endif
                uint32_t flipm = 0u;
                if (cand) {
                    // ---- scan (lane = chain tc + TS * part; TS variables of the sub-window per lane) ----
                    // this lane's bound of variable a (its own part only): a load per step, independent of the decisions
                    const float *bo = sm.bounds + (wi & 1u) * (4 * 32 * T) + (k * 32 + part * TS) * T + t;
                    const uint32_t dbuf_s = smem_u32(D) + (uint32_t)(32 * 32) * 4u + (uint32_t)part * (TS * 4u);
                    const uint32_t cbuf_s = smem_u32(sm.Cbuf) + (uint32_t)(rb * 32 * T + t) * 4u;
                    const uint32_t rmask_s = smem_u32(sm.rowmask) + (uint32_t)(rb * 32) * 4u + (uint32_t)wsc * (4u / NSCAN);
                    // branch-free: every step applies c * D[a][.] with c = 0 for chains that keep variable a; the
                    // shared-memory reads do not depend on the decisions (row a+1 is fetched before the vote of row a),
                    // only field -> dE -> compare -> ballot -> c -> fma is a dependent chain
                    float4 dcur[Q4], dnxt[Q4];
#pragma unroll
                    for (int q4 = 0; q4 < Q4; ++q4) dcur[q4] = lds128(dbuf_s + (uint32_t)(q4 * 4) * 4u);
#pragma unroll
                    for (int a = 0; a < 32; ++a) {
                        const int pa = a / TS, i = a % TS;
                        if (a < 31) {
#pragma unroll
                            for (int q4 = 0; q4 < Q4; ++q4) dnxt[q4] = lds128(dbuf_s + (uint32_t)((a + 1) * 32 + q4 * 4) * 4u);
                        }
                        const float sg = ((old >> a) & 1u) ? -2.0f : 2.0f;        // variable a has not been visited yet
                        const float dE = __fmul_rn(G[i], sg);
                        const bool acc = (part == pa) & alive & (a < rem) & ((dE <= 0.0f) | (dE < bo[i * T]));
                        const uint32_t bal = __ballot_sync(FULL, acc);
                        const uint32_t oldbal = __ballot_sync(FULL, (old >> a) & 1u);          // off the dependent chain
                        if (lane == 0) {
                            const uint32_t w = ((bal >> (TS * pa)) & TSMASK) | ((oldbal & TSMASK) << TS);
                            if (NSCAN == 1) asm volatile("st.shared.u32 [%0], %1;" ::"r"(rmask_s + (uint32_t)a * 4u), "r"(w) : "memory");
                            else asm volatile("st.shared.u16 [%0], %1;" ::"r"(rmask_s + (uint32_t)a * 4u), "h"((unsigned short)w) : "memory");
                        }
                        const bool mine = (bal >> (TS * pa + tc)) & 1u;
                        const float c = mine ? sg : 0.0f;
#pragma unroll
                        for (int q4 = 0; q4 < Q4; ++q4) {
                            G[4 * q4 + 0] = __fmaf_rn(c, dcur[q4].x, G[4 * q4 + 0]);
                            G[4 * q4 + 1] = __fmaf_rn(c, dcur[q4].y, G[4 * q4 + 1]);
                            G[4 * q4 + 2] = __fmaf_rn(c, dcur[q4].z, G[4 * q4 + 2]);
                            G[4 * q4 + 3] = __fmaf_rn(c, dcur[q4].w, G[4 * q4 + 3]);
                        }
                        if (p0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(cbuf_s + (uint32_t)a * (T * 4u)), "f"(c) : "memory");
                        flipm |= (mine ? 1u : 0u) << a;
#pragma unroll
                        for (int q4 = 0; q4 < Q4; ++q4) dcur[q4] = dnxt[q4];
                    }
                } else if (NSCAN > 1) {
                    // nothing to decide for these chains, but another scanner's chains may flip rows of this record: the
                    // appliers then read this scanner's coefficients and mask halves too (lane = row a)
                    const uint32_t cz = smem_u32(sm.Cbuf) + (uint32_t)((rb * 32 + lane) * T + wsc * TS) * 4u;
#pragma unroll
                    for (int j = 0; j < TS / 4; ++j)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(cz + (uint32_t)j * 16u), "f"(0.0f) : "memory");
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(sm.rowmask) + (uint32_t)(rb * 32 + lane) * 4u + (uint32_t)wsc * 2u),
                                 "h"((unsigned short)0) : "memory");
                }
#ifdef QBM_TILE_PROF
                c_scan += (uint32_t)clock() - t_scan0;
#endif
                const uint32_t unionm = __reduce_or_sync(FULL, flipm);
                const uint32_t cnt = __reduce_add_sync(FULL, p0 ? (uint32_t)__popc(flipm) : 0u);
                if (p0) sm.spinw[s * T + t] = old ^ flipm;
                if (lane == 0) {
                    sm.rec_meta[rb * META + M_UNION + wsc] = unionm;
                    sm.rec_meta[rb * META + M_COUNT + wsc] = cnt;
                    sm.rec_meta[rb * META + M_ROW0] = (uint32_t)(s * 32);
                    sm.rec_meta[rb * META + M_EXIT] = 0u;
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sm.rec_full[rb]);
                    if (k == 3 || s + 1 == S) mbar_arrive(&sm.bnd_empty[wi & 1u]);      // bounds of this window consumed
                }
                if (k == 3 || s + 1 == S) ++wi;
                nacc += cnt;
                prev_union = unionm; prev_mask = flipm; prev_neg = old; prev_mag = 2.0f;
            }
            // two-phase schedule: once a sweep accepts less than hot_fraction of its proposals the chains are cheaper to
            // advance one warp each (rows no longer shared by most chains); the scanners add up their counts (one
            // rendezvous per sweep) and take the same decision
            if (hot_min > 0ull) {
                unsigned long long flips = nacc - nacc_before;
                if (NSCAN > 1) {
                    volatile uint32_t *cw = sm.ctl + 1 + (t_sweep & 1u) * NSCAN;
                    if (lane == 0) cw[wsc] = (uint32_t)flips;
                    scanners_sync();
                    flips = 0ull;
#pragma unroll
                    for (int j = 0; j < NSCAN; ++j) flips += cw[j];
                }
                if (flips < hot_min) handed_over = true;
            }
        }
    }
    cp_async_wait_all();
#ifdef QBM_TILE_PROF
    if (lane == 0 && wsc == 0) {
        PROF_PUT(8, (uint32_t)clock() - t_start); PROF_PUT(9, w_fx); PROF_PUT(10, w_bnd); PROF_PUT(11, w_rece); PROF_PUT(12, w_cp);
        PROF_PUT(13, c_catch); PROF_PUT(14, c_scan);
    }
#endif
    // ---- exit record; the flag stops the bounds warp, which polls it while it waits for a free buffer ----
    {
        const uint32_t rb = r & (NREC - 1);
        mbar_wait(&sm.rec_empty[rb], ((r / NREC) & 1u) ^ 1u);
        if (lane == 0) {
            sm.rec_meta[rb * META + M_UNION + wsc] = 0u;
            sm.rec_meta[rb * META + M_COUNT + wsc] = 0u;
            sm.rec_meta[rb * META + M_ROW0] = 0u;
            sm.rec_meta[rb * META + M_EXIT] = 1u;
            *reinterpret_cast<volatile uint32_t *>(sm.ctl) = 1u;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.rec_full[rb]);
    }
    if (lane == 0) {
        if (p.sweeps_done != nullptr && wsc == 0) p.sweeps_done[blockIdx.x] = t_sweep;      // completed sweeps of this tile
        if (p.counters != nullptr) {
            atomicAdd(p.counters + 0, nacc);
            atomicAdd(p.counters + 1, (unsigned long long)n * (unsigned long long)t_sweep * (unsigned long long)nlive_w);
        }
    }
    tile_sync((W + NSCAN) * 32);      // spins final: the appliers write the states
}

// NS == 8 (n > 1024): W = windows / 2 applier warps x 8 columns; NS == 4 (n <= 1024): W = windows applier warps x 4 columns.
template <int NS, int AW>
__global__ void __launch_bounds__(Slots<AW>::nthreads, Slots<AW>::min_ctas) sa_tile_kernel(const SaParams p, const int W, const int ctas_per_problem)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n = p.n, ld = p.ld;
    const int S = (n + 31) >> 5;                       // populated 32-variable sub-windows
    const TileSmem sm = carve(smem_raw);

    const long long q = blockIdx.x / ctas_per_problem;
    const long long r0 = (long long)(blockIdx.x % ctas_per_problem) * T;
    const int nlive = (int)min((long long)T, p.num_reads - r0);
    const float *__restrict__ Jp = p.Jp + (size_t)q * (size_t)n * (size_t)ld;
    const float *__restrict__ betas = p.beta + q * p.beta_stride;
    const long long cl0 = q * p.num_reads + r0;        // row of chain 0 of this CTA in init / out
    const unsigned long long chain0 = p.chain_offset + (unsigned long long)((p.flags & 2u) ? r0 : cl0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NGS; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], (uint32_t)W); }
        for (int i = 0; i < NREC; ++i) { mbar_init(&sm.rec_full[i], NSCAN); mbar_init(&sm.rec_empty[i], (uint32_t)W + 1u); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.fx_full[i], 1); mbar_init(&sm.fx_empty[i], NSCAN);
            mbar_init(&sm.bnd_full[i], 1); mbar_init(&sm.bnd_empty[i], NSCAN);
        }
        sm.ctl[0] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= AW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Slots<AW>::regs_helper));
        if (warp == Slots<AW>::producer) {
            if (lane == 0) tile_producer(sm, Jp, ld);
        } else if (warp == Slots<AW>::bounds) {
            tile_bounds(sm, p, betas, chain0, lane);
        } else if (scanner_of_warp<AW>(warp) >= 0) {
            tile_sync((W + NSCAN) * 32);                // initial spins written
            tile_scanner(sm, p, p.Jnat + (size_t)q * (size_t)n * (size_t)p.ldj, betas, W, nlive, lane, scanner_of_warp<AW>(warp));
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Slots<AW>::regs_applier));
    if (warp >= W) return;                              // idle warps of the applier warpgroups (W < 8)

    // ================================= applier warps =================================
    constexpr int NWIN = NS / 4;
    const float *__restrict__ hq = p.hp + (size_t)q * (size_t)ld;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const uint32_t dense_pct = ((p.flags >> 8) & 0xffu) ? ((p.flags >> 8) & 0xffu) : 40u;
    const uint32_t dense_min = max(1u, (uint32_t)nlive * 32u * dense_pct / 100u);
    const int tid = threadIdx.x;
    const int ncons = W * 32;

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
    tile_sync((W + NSCAN) * 32);                        // the scanners take the spins from here on

    f32x2 F2[NS / 2][T];       // F2[jw*2 + h][t] = packed fields of sub-windows (2h, 2h+1) of window jw, chain t
#pragma unroll
    for (int jw = 0; jw < NWIN; ++jw) {
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(hq + (size_t)(jw * W + warp) * 128 + lane * 4));
#pragma unroll
        for (int t = 0; t < T; ++t) { F2[jw * 2 + 0][t] = pack2(hv.x, hv.y); F2[jw * 2 + 1][t] = pack2(hv.z, hv.w); }
    }
    TileAddr A;
    A.ring = smem_u32(sm.ring) + (uint32_t)(warp * 32 + lane) * 16u;
    A.slot_stride = (uint32_t)ld * 4u;
    A.win_stride = (uint32_t)W * 512u;
    A.cbuf = smem_u32(sm.Cbuf);
    A.full = smem_u32(sm.full);
    A.empty = smem_u32(sm.empty);
    A.rowmask = smem_u32(sm.rowmask);

    // ---- records: S initial-field records, then one per sub-window and sweep, until the exit record ----
    PROF_VAR(w_ring); PROF_VAR(w_recf); PROF_VAR(w_fxe);
#ifdef QBM_TILE_PROF
    const uint32_t t_start = (uint32_t)clock();
#else
    uint32_t w_ring = 0u;      // (unused outside the profiling build)
#endif
    uint32_t gi = 0;
    int s_next = 0;                                     // sub-window of record r + 1 once annealing has started
    for (uint32_t r = 0;; ++r) {
        if (r + 1 >= (uint32_t)S) {
            // record r + 1 belongs to sub-window s_next: its owner hands the scanner that sub-window's fields (all
            // records before r applied), from which the scanner decides record r + 1 while record r is applied here
            const uint32_t e1 = r + 1 - (uint32_t)S;
            const int g1 = s_next >> 2, k1s = s_next & 3;
            if (g1 % W == warp) {
                const int slot = (g1 / W) * 4 + k1s;
                PROF_WAIT(w_fxe, mbar_wait(&sm.fx_empty[e1 & 1u], ((e1 >> 1) & 1u) ^ 1u));
                // [column][chain]; the columns of part q shifted by q * TS floats: the scanners' reads are conflict-free
                float4 *dst = reinterpret_cast<float4 *>(sm.Fx + (e1 & 1u) * FXLD + lane * T + (lane / TS) * TS);
                // register arrays need static indices: one copy of the four stores per column slot, the slot is warp-uniform
#pragma unroll
                for (int j = 0; j < NS; ++j)
                    if (j == slot) {
#pragma unroll
                        for (int q4 = 0; q4 < T / 4; ++q4)
                            dst[q4] = (j & 1) ? make_float4(hi2(F2[j >> 1][4 * q4]), hi2(F2[j >> 1][4 * q4 + 1]), hi2(F2[j >> 1][4 * q4 + 2]), hi2(F2[j >> 1][4 * q4 + 3]))
                                              : make_float4(lo2(F2[j >> 1][4 * q4]), lo2(F2[j >> 1][4 * q4 + 1]), lo2(F2[j >> 1][4 * q4 + 2]), lo2(F2[j >> 1][4 * q4 + 3]));
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.fx_full[e1 & 1u]);
            }
            s_next = (s_next + 1 == S) ? 0 : s_next + 1;
        }
        const uint32_t rb = r & (NREC - 1);
        PROF_WAIT(w_recf, mbar_wait(&sm.rec_full[rb], (r / NREC) & 1u));
        uint32_t u = sm.rec_meta[rb * META + M_UNION], cnt = sm.rec_meta[rb * META + M_COUNT];
        if (NSCAN == 2) { u |= sm.rec_meta[rb * META + M_UNION + 1]; cnt += sm.rec_meta[rb * META + M_COUNT + 1]; }
        const uint32_t ex = sm.rec_meta[rb * META + M_EXIT];
        if (ex) break;
        if (u != 0u) apply_record<NS>(F2, A, u, cnt, rb, lane, gi, dense_min, w_ring);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.rec_empty[rb]);
    }

#ifdef QBM_TILE_PROF
    if (lane == 0) {      // per applier warp: total, waits for ring slots / records / the export buffer
        PROF_PUT(0, (uint32_t)clock() - t_start); PROF_PUT(1, w_ring); PROF_PUT(2, w_recf); PROF_PUT(3, w_fxe);
        if (warp == 0) { PROF_PUT(4, (uint32_t)clock() - t_start); PROF_PUT(5, w_ring); PROF_PUT(6, w_recf); PROF_PUT(7, w_fxe); }
    }
#endif
    if (p.fields != nullptr) {
        // fields in the warp kernel's layout: chain-major, float4 (sub-windows 0..3) per lane and 128-variable window
#pragma unroll
        for (int jw = 0; jw < NWIN; ++jw) {
            const size_t off = (size_t)(jw * W + warp) * 128 + (size_t)lane * 4;
#pragma unroll
            for (int t = 0; t < T; ++t)
                if (t < nlive)
                    *reinterpret_cast<float4 *>(p.fields + (size_t)(cl0 + t) * (size_t)ld + off) =
                        make_float4(lo2(F2[jw * 2][t]), hi2(F2[jw * 2][t]), lo2(F2[jw * 2 + 1][t]), hi2(F2[jw * 2 + 1][t]));
        }
    }

    // ---- write-back: states in natural variable order, 0/1 ----
    tile_sync((W + NSCAN) * 32);
    for (int t = warp; t < nlive; t += W) {
        int8_t *o = p.out + (size_t)(cl0 + t) * (size_t)n;
        for (int s = 0; s < S; ++s) {
            const int v = s * 32 + lane;
            if (v < n) o[v] = (int8_t)((sm.spinw[s * T + t] >> lane) & 1u);
        }
    }
}

template <int NS, int AW>
int launch_tile(const SaParams &p, int W, cudaStream_t st)
{
    auto kern = sa_tile_kernel<NS, AW>;
    const long long cpp = (p.num_reads + T - 1) / T;
    const long long blocks = cpp * p.batch_q;
    if (blocks > 0x7fffffffLL) {
        qbm_set_error("qbm_sa_sample: too many chains for one launch (%lld)", p.total_chains);
        return QBM_EUNSUPPORTED;
    }
    const size_t smem = tile_smem_bytes(p.ld);
    // per launch, not cached: the attribute belongs to the current device, and a process may drive several
    QBM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, Slots<AW>::nthreads, smem, st>>>(p, W, (int)cpp);
    QBM_LAUNCH_OK("sa_tile_kernel");
    return QBM_OK;
}

}  // namespace

bool sa_tile_supported(int n) { return n >= 1 && n <= QBM_SA_MAX_N; }

// row length of the permuted coupling matrix the tile kernel reads: whole 128-variable windows, one
// (n <= 1024) or two (n > 1024) per applier warp
int sa_tile_ld(int n)
{
    const int win = (n + 127) / 128;
    if (win <= 8) return win * 128;
    return ((win + 1) / 2) * 2 * 128;
}

int sa_tile_launch(const SaParams &p, cudaStream_t st)
{
    const int win = p.ld / 128;
    if (win <= 4) return launch_tile<4, 4>(p, win, st);                    // n <= 512: the same with four applier slots, two CTAs per SM
    if (win <= 8) return launch_tile<4, 8>(p, win, st);                    // n <= 1024: W = win applier warps x 4 columns
    return launch_tile<8, 8>(p, win / 2, st);                              // n >  1024: W = win / 2 applier warps x 8 columns
}

#ifdef QBM_TILE_PROF
// profiling build only: read (and clear) the wait-clock counters of the chain-tile kernel
extern "C" QBM_API int qbm_debug_tile_prof(unsigned long long *out32, int reset)
{
    QBM_CUDA_OK(cudaDeviceSynchronize());
    if (out32) QBM_CUDA_OK(cudaMemcpyFromSymbol(out32, g_tile_prof, sizeof(unsigned long long) * 32));
    if (reset) {
        unsigned long long z[32] = {};
        QBM_CUDA_OK(cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z)));
    }
    return QBM_OK;
}
#endif
