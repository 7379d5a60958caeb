// K1: simulated-annealing QUBO sampler for sm_100a -- one warp per read (chain).
//
// Replaces the Metropolis loop of dwave-neal 0.5.9 (cpu_sa.cpp) that the reference reaches through
// src/qubo/sampler.py:31-33 and src/model/faster_dqbm.py:299-313.  The rule
// is the reference's (fixed sweep order, threshold skip, dE<=0 auto-accept, u < exp(-beta dE) -- evaluated
// in the log domain, dE < -ln(u)/beta, so a proposal that is re-evaluated after a neighbour flipped costs one
// compare); the layout is B200-first:
//   * the chain's n local fields live in registers, 4*NW per lane (variable v = w*128 + k*32 + lane
//     is field k of window w of `lane`); spins are 4*NW bits per lane
//   * the coupling matrix is stored column-permuted (p128_pos) so the four fields of a lane are one
//     128-bit load per window; rows stream through L1 (read-only path): the sweep order is fixed, so
//     the chains of an SM need row v at about the same time and mostly find it L1-resident (87 % hit
//     rate at n = 2048)
//   * proposals are evaluated 32 at a time (one sub-window of 32 consecutive variables, one per lane):
//     because the uniform for (chain, sweep, v) is a pure function of its index (Philox4x32-10), the
//     first accepted proposal of the sub-window is found with one ballot, every earlier proposal is
//     a rejection that changes nothing, and evaluation restarts right after the flipped variable.
//     The trajectory is therefore exactly the sequential one (oracle/replay_sa.c).
//
// Template parameters: NW = 128-variable windows held per lane, KS = sub-windows evaluated per window (KS < 4 only for
// NW == 1, i.e. n <= 32 / 64), WPC = warps (chains) per CTA, MINB = CTAs per SM the register budget is sized for, and three
// code-shape switches that do not change any result:
//   UW  the loop over windows is unrolled, so the window being swept is addressed statically and needs no working copy
//       (one 128-bit load and four FMAs less per flip; code size grows NW-fold, so only for small NW)
//   P2  field updates as packed fma.rn.f32x2 (two IEEE fp32 FMAs per instruction)
//   PIN the draw of the acceptance bounds is pinned inside its (rarely taken) branch; without it the compiler speculates the
//       pure Philox + log above the branch, which costs ~140 instructions per window in every cold sweep but overlaps load
//       latency in the hot ones (which wins depends on the instantiation; measured)
//   SH  the update coefficient -2 s_a comes from the flipping lane by shuffle instead of a warp-uniform spin mask
#pragma once
#include "sa_common.cuh"

namespace sa_warp {

constexpr unsigned FULL = 0xffffffffu;

// two fp32 FMAs per instruction (Blackwell FFMA2): acc.{x,y} = fma(a.{x,y}, b.{x,y}, acc.{x,y}), each round-to-nearest
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

// the four fields a lane holds of one window: p[0] = sub-windows (0, 1), p[1] = sub-windows (2, 3)
struct Win {
    float2 p[2];
};
__device__ __forceinline__ float wget(const Win &f, const int k) { return (k & 1) ? f.p[k >> 1].y : f.p[k >> 1].x; }

template <bool P2>
__device__ __forceinline__ void win_update(Win &f, const float4 r, const float c)
{
    if (P2) {
        const float2 cc = make_float2(c, c);
        ffma2(f.p[0], cc, make_float2(r.x, r.y));
        ffma2(f.p[1], cc, make_float2(r.z, r.w));
    } else {
        f.p[0].x = __fmaf_rn(c, r.x, f.p[0].x);
        f.p[0].y = __fmaf_rn(c, r.y, f.p[0].y);
        f.p[1].x = __fmaf_rn(c, r.z, f.p[1].x);
        f.p[1].y = __fmaf_rn(c, r.w, f.p[1].y);
    }
}

template <int NW, bool P2>
__device__ __forceinline__ void row_update(Win (&F)[NW], const float *__restrict__ row_lane, const float c)
{
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) win_update<P2>(F[w2], __ldg(reinterpret_cast<const float4 *>(row_lane + w2 * 128)), c);
}

// RT: the windows rotate through the register sets, set 0 is always the window being swept (and row r is stored with its
// own window first, see sa_permute_kernel), so the swept window is addressed statically without unrolling the window loop
template <int NW>
__device__ __forceinline__ void rotate_windows(Win (&F)[NW])
{
    const Win first = F[0];
#pragma unroll
    for (int j = 0; j + 1 < NW; ++j) F[j] = F[j + 1];
    F[NW - 1] = first;
}

template <int NW, int KS, int WPC, int MINB, bool UW, bool P2, bool PIN, bool SH, bool RT, bool RS>
__global__ void __launch_bounds__(WPC * 32, MINB) sa_kernel(const SaParams p)
{
    constexpr bool DIRECT = (NW == 1) || UW || RT;      // the swept window's fields are addressed statically
    static_assert(!(UW && RT), "unrolled or rotating windows, not both");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // RS (resuming the chains of the chain-tile kernel): a CTA is one of its tiles of 16 chains of one problem, so that the
    // first sweep (sweeps_done[tile]) is a CTA-uniform value and the sweep loops stay uniform for the compiler
    static_assert(!RS || WPC == 16, "a resuming CTA is one tile of the chain-tile kernel");
    const long long tiles_per_problem = (p.num_reads + WPC - 1) / WPC;
    long long cl = RS ? (blockIdx.x / tiles_per_problem) * p.num_reads + (blockIdx.x % tiles_per_problem) * WPC + warp
                      : (long long)blockIdx.x * WPC + warp;
    const bool live = RS ? ((blockIdx.x % tiles_per_problem) * WPC + warp < p.num_reads) : (cl < p.total_chains);
    if (!live) cl = p.total_chains - 1;           // idle warps shadow the last chain (they keep the barriers matched)
    const long long q = cl / p.num_reads;
    const int n = p.n;
    constexpr int ld = NW * 128;                  // rows are padded to whole windows of this instantiation (sa_ld)
    // this lane's float4 column of row 0; a row is reached with a 32-bit element offset (n * ld <= 2^22)
    const float *__restrict__ J = p.Jp + (size_t)q * (size_t)n * (size_t)ld + lane * 4;
    const float *__restrict__ hq = p.hp + (size_t)q * (size_t)ld;
    const float *__restrict__ betas = p.beta + q * p.beta_stride;
    // flag bit 1: key the stream by the read index only, so every problem of the batch sees the same
    // random stream -- what the reference does by passing the same seed to every call (Appendix B Q6)
    const unsigned long long chain = p.chain_offset + (unsigned long long)((p.flags & 2u) ? (cl - q * p.num_reads) : cl);
    const uint32_t c_lo = (uint32_t)chain, c_hi = (uint32_t)(chain >> 32);
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    // optional per-window rendezvous of the CTA's warps (flag bit 0).  Off by default: chains that anneal the same problem
    // stay on nearby coupling rows by themselves (a leading warp takes the L1 misses and is caught up by the others), and
    // the barrier measured 0..17 % slower (n = 384..2048)
    const bool rendezvous = (p.flags & 1u) != 0u && NW >= 3;
    const int nw_rt = (n + 127) >> 7;             // windows actually populated (<= NW)

    Win F[NW];
    unsigned long long spins = 0ull;              // bit (w*4+k) = spin of variable w*128 + k*32 + lane (1 = up)

    // ---- initial spins and local fields: F_i = h_i ; for j = 0..n-1: F_i = fma(J[j][i], s_j, F_i) ----
    // resuming chains handed over by the chain-tile kernel: fields as it left them (same register layout), spins from
    // p.init (= the states it wrote), first sweep = sweeps_done[chain]
    constexpr bool resume = RS;
    const float *__restrict__ f0 = resume ? p.fields + (size_t)cl * (size_t)ld : hq;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(f0 + w * 128 + lane * 4));
        F[w].p[0] = make_float2(hv.x, hv.y);
        F[w].p[1] = make_float2(hv.z, hv.w);
    }
#pragma unroll 1
    for (int w = 0; w < (RT ? NW : nw_rt); ++w) {
        if (RT && w >= nw_rt) {                   // empty window: only keep the rotation in step
            rotate_windows<NW>(F);
            continue;
        }
        uint32_t wd[4];
        if (p.init != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = w * 128 + k * 32 + lane;
                const int8_t b = (v < n) ? p.init[(size_t)cl * (size_t)n + v] : (int8_t)0;
                wd[k] = __ballot_sync(FULL, b != 0);
            }
        } else {
            const Philox4 o = philox4x32_10(c_lo, c_hi, 0xFFFFFFFFu, (uint32_t)w, k0, k1);
            wd[0] = o.x; wd[1] = o.y; wd[2] = o.z; wd[3] = o.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            spins |= (unsigned long long)((wd[k] >> lane) & 1u) << (w * 4 + k);
            const int jbase = w * 128 + k * 32;
            const int jend = resume ? 0 : min(32, n - jbase);
            for (int jj = 0; jj < jend; ++jj) {
                const float sj = ((wd[k] >> jj) & 1u) ? 1.0f : -1.0f;
                row_update<NW, false>(F, J + (uint32_t)(jbase + jj) * (uint32_t)ld, sj);
            }
        }
        if (RT) rotate_windows<NW>(F);
    }

    // ---- annealing ----
    uint32_t nacc = 0;
    uint32_t t = resume ? p.sweeps_done[blockIdx.x] : 0u;
    for (int b = resume ? (int)(t / (uint32_t)p.sweeps_per_beta) : 0; b < p.num_betas; ++b) {
        const float beta = __ldg(betas + b);
        const float thr = __fdiv_rn(44.36142f, beta);
        for (int s = resume ? (int)(t - (uint32_t)b * (uint32_t)p.sweeps_per_beta) : 0; s < p.sweeps_per_beta; ++s, ++t) {
            // one window: 4 sub-windows of 32 proposals; Fc = the window's four fields (the fields themselves when the
            // window index is static, else a working copy: register indices must be compile-time)
            auto sweep_window = [&](const int w, Win &Fc) {
                if (rendezvous) __syncthreads();
                uint32_t s4 = (uint32_t)(spins >> (w * 4)) & 15u;
                // acceptance bounds of this lane's four proposals: flip <=> dE <= 0 or dE < bnd, with
                // bnd = min(thr, -ln(u/2^32)/beta) drawn lazily (a pure function of (chain, sweep, variable))
                bool have_rng = false;
                float bnd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const int vbase = w * 128 + k * 32;
                    if (vbase >= n) break;
                    const int rem = n - vbase;
                    unsigned todo = rem >= 32 ? FULL : ((1u << rem) - 1u);   // proposals not yet passed, in sweep order
                    const bool up0 = (s4 >> k) & 1u;
                    unsigned upm = SH ? 0u : __ballot_sync(FULL, up0);       // spins of the sub-window (warp-uniform)
                    float sgn = up0 ? -2.0f : 2.0f;                          // dE = sgn * F
                    while (true) {
                        const float dE = __fmul_rn(wget(Fc, k), sgn);
                        if (!have_rng) {
                            const bool pend = (dE > 0.0f) && (dE < thr);
                            if (__ballot_sync(FULL, pend) & todo) {
                                uint32_t tt = t;
                                if (PIN) asm volatile("" : "+r"(tt));
                                const Philox4 o = philox4x32_10(c_lo, c_hi, tt, (uint32_t)(w * 32 + lane), k0, k1);
                                const uint32_t u[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                for (int q4 = 0; q4 < KS; ++q4) bnd[q4] = fminf(thr, __fdiv_rn(neg_log_u32(u[q4]), beta));
                                have_rng = true;
                            }
                        }
                        const bool acc = (dE <= 0.0f) || (dE < bnd[k]);
                        const unsigned m = __ballot_sync(FULL, acc) & todo;
                        if (m == 0u) break;
                        const int a = __ffs(m) - 1;                          // first accepted proposal in sweep order
                        // -2 * s_a(old): from the warp-uniform spin mask, or (SH) the flipping lane's own sign
                        const float c = SH ? __shfl_sync(FULL, sgn, a) : (((upm >> a) & 1u) ? -2.0f : 2.0f);
                        if (!SH) upm ^= 1u << a;
                        if (lane == a) sgn = -sgn;
                        const float *row = J + (uint32_t)(vbase + a) * (uint32_t)ld;
                        if (!DIRECT) win_update<P2>(Fc, __ldg(reinterpret_cast<const float4 *>(row + w * 128)), c);
                        row_update<NW, P2>(F, row, c);
                        todo &= ~((2u << a) - 1u);
                        ++nacc;
                    }
                    s4 = (s4 & ~(1u << k)) | ((SH ? (sgn < 0.0f ? 1u : 0u) : ((upm >> lane) & 1u)) << k);
                }
                spins = (spins & ~(15ull << (w * 4))) | ((unsigned long long)s4 << (w * 4));
            };
            if (NW == 1) {
                sweep_window(0, F[0]);
            } else if (UW) {
#pragma unroll
                for (int w = 0; w < NW; ++w)
                    if (w < nw_rt) sweep_window(w, F[w]);
            } else if (RT) {
#pragma unroll 1
                for (int w = 0; w < NW; ++w) {
                    if (w < nw_rt) sweep_window(w, F[0]);
                    rotate_windows<NW>(F);
                }
            } else {
                for (int w = 0; w < nw_rt; ++w) {
                    Win Fc;
                    Fc.p[0] = make_float2(0.0f, 0.0f);
                    Fc.p[1] = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int w2 = 0; w2 < NW; ++w2)
                        if (w2 == w) Fc = F[w2];
                    sweep_window(w, Fc);
                }
            }
        }
    }

    // ---- write-back: states in natural variable order, 0/1 ----
    if (live) {
        int8_t *o = p.out + (size_t)cl * (size_t)n;
        for (int w = 0; w < nw_rt; ++w) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = w * 128 + k * 32 + lane;
                if (v < n) o[v] = (int8_t)((spins >> (w * 4 + k)) & 1ull);
            }
        }
        if (p.counters != nullptr && lane == 0) {
            atomicAdd(p.counters + 0, (unsigned long long)nacc);
            atomicAdd(p.counters + 1, (unsigned long long)n * (unsigned long long)(t - (resume ? p.sweeps_done[blockIdx.x] : 0u)));
        }
    }
}

template <int NW, int KS, int WPC, int MINB, bool UW = false, bool P2 = false, bool PIN = (NW <= 2 || NW >= 6), bool SH = false,
          bool RT = false, bool RS = false>
int launch_sa(const SaParams &p, cudaStream_t st)
{
    auto kern = sa_kernel<NW, KS, WPC, MINB, UW, P2, PIN, SH, RT, RS>;
    // all on-chip memory as L1: coupling rows are shared between the chains of an SM through L1
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    const long long blocks = RS ? ((p.num_reads + WPC - 1) / WPC) * p.batch_q : (p.total_chains + WPC - 1) / WPC;
    if (blocks > 0x7fffffffLL) {
        qbm_set_error("qbm_sa_sample: too many chains for one launch (%lld)", p.total_chains);
        return QBM_EUNSUPPORTED;
    }
    if (p.ld != NW * 128) {
        qbm_set_error("qbm_sa_sample: internal error: row stride %d does not match the kernel variant (%d)", p.ld, NW * 128);
        return QBM_EINVAL;
    }
    kern<<<(unsigned)blocks, WPC * 32, 0, st>>>(p);
    QBM_LAUNCH_OK("sa_kernel");
    return QBM_OK;
}

}  // namespace sa_warp
