// K1c: simulated-annealing QUBO sampler for sm_100a -- one warp anneals T chains of the same problem and
// shares coupling-row loads between them.
//
// Same trajectory as sa_kernel.cu (oracle/replay_sa.c).  The wall of one-warp-one-chain
// is the L1 data pipe: every accepted flip moves a 4n-byte row from L1 into the registers of one warp.  In
// the hot part of neal's legacy schedule the acceptance rate is > 0.9, so chains that sweep together flip
// the SAME variable most of the time.  Here a warp holds the fields of T chains (4*NW*T registers per lane);
// in every round each chain that changed re-evaluates its 32 proposals (one ballot), the warp takes the
// earliest accepted variable a over its chains, loads row a ONCE and applies it to every chain whose first
// accepted proposal is a (coefficient -2 s_a) -- and with coefficient 0 to the others, which leaves their
// fields untouched.  Chains that did not flip keep their cached ballot.  Each chain still sees exactly its own
// sequential trajectory (flips in sweep order, the same FMA per field element in the same order).
#include "sa_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// two fp32 FMAs per instruction (Blackwell FFMA2): acc.{x,y} = fma(a.{x,y}, b.{x,y}, acc.{x,y}), each round-to-nearest
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

// NW = 128-variable windows per lane, T = chains per warp, WPC = warps per CTA
template <int NW, int T, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB) sa_multi_kernel(const SaParams p, const long long groups_per_problem,
                                                                 const long long total_groups)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    long long grp = (long long)blockIdx.x * WPC + warp;
    const bool warp_live = grp < total_groups;
    if (!warp_live) grp = total_groups - 1;          // idle warps shadow the last group (they keep the barriers matched)
    const long long q = grp / groups_per_problem;
    const long long r0 = (grp % groups_per_problem) * T;
    const int n = p.n;
    const int ld = p.ld;
    const float *__restrict__ J = p.Jp + (size_t)q * (size_t)n * (size_t)ld;
    const float *__restrict__ hq = p.hp + (size_t)q * (size_t)ld;
    const float *__restrict__ betas = p.beta + q * p.beta_stride;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const bool rendezvous = (p.flags & 1u) != 0u;
    const int nw_rt = (n + 127) >> 7;

    // chain t of this warp: read r0 + t of problem q (reads past num_reads shadow the last read and are not written)
    long long cl[T];
    uint32_t c_lo[T], c_hi[T];
    bool clive[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const long long r = min(r0 + t, p.num_reads - 1);
        clive[t] = warp_live && (r0 + t) < p.num_reads;
        cl[t] = q * p.num_reads + r;
        const unsigned long long chain = p.chain_offset + (unsigned long long)((p.flags & 2u) ? r : cl[t]);
        c_lo[t] = (uint32_t)chain; c_hi[t] = (uint32_t)(chain >> 32);
    }

    // fields as float2 pairs: F2[t][w][0] = sub-windows (0,1), F2[t][w][1] = sub-windows (2,3) of window w
    float2 F2[T][NW][2];
    unsigned long long spins[T];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(hq + w * 128 + lane * 4));
#pragma unroll
        for (int t = 0; t < T; ++t) { F2[t][w][0] = make_float2(hv.x, hv.y); F2[t][w][1] = make_float2(hv.z, hv.w); }
    }
    // ---- initial spins and fields: F_i = h_i ; for j = 0..n-1: F_i = fma(J[j][i], s_j, F_i) (row j loaded once for all chains) ----
#pragma unroll
    for (int t = 0; t < T; ++t) spins[t] = 0ull;
    for (int w = 0; w < nw_rt; ++w) {
        uint32_t wd[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (p.init != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int v = w * 128 + k * 32 + lane;
                    const int8_t b = (v < n) ? p.init[(size_t)cl[t] * (size_t)n + v] : (int8_t)0;
                    wd[t][k] = __ballot_sync(FULL, b != 0);
                }
            } else {
                const Philox4 o = philox4x32_10(c_lo[t], c_hi[t], 0xFFFFFFFFu, (uint32_t)w, k0, k1);
                wd[t][0] = o.x; wd[t][1] = o.y; wd[t][2] = o.z; wd[t][3] = o.w;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int t = 0; t < T; ++t) spins[t] |= (unsigned long long)((wd[t][k] >> lane) & 1u) << (w * 4 + k);
            const int jbase = w * 128 + k * 32;
            const int jend = min(32, n - jbase);
            for (int jj = 0; jj < jend; ++jj) {
                float2 cc[T];
#pragma unroll
                for (int t = 0; t < T; ++t) { const float s = ((wd[t][k] >> jj) & 1u) ? 1.0f : -1.0f; cc[t] = make_float2(s, s); }
                const float *row = J + (size_t)(jbase + jj) * (size_t)ld + lane * 4;
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) {
                    const float4 r = __ldg(reinterpret_cast<const float4 *>(row + w2 * 128));
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        ffma2(F2[t][w2][0], cc[t], make_float2(r.x, r.y));
                        ffma2(F2[t][w2][1], cc[t], make_float2(r.z, r.w));
                    }
                }
            }
        }
    }

    // ---- annealing ----
    uint32_t nacc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) nacc[t] = 0u;
    uint32_t ts = 0;
    for (int b = 0; b < p.num_betas; ++b) {
        const float beta = __ldg(betas + b);
        const float thr = __fdiv_rn(44.36142f, beta);
        for (int s = 0; s < p.sweeps_per_beta; ++s, ++ts) {
            for (int w = 0; w < nw_rt; ++w) {
                if (rendezvous) __syncthreads();
                // working copies of this window's fields (register index must be static)
                float2 Fc[T][2];
#pragma unroll
                for (int t = 0; t < T; ++t) { Fc[t][0] = make_float2(0.0f, 0.0f); Fc[t][1] = make_float2(0.0f, 0.0f); }
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2)
                    if (w2 == w) {
#pragma unroll
                        for (int t = 0; t < T; ++t) { Fc[t][0] = F2[t][w2][0]; Fc[t][1] = F2[t][w2][1]; }
                    }
                uint32_t s4[T];
                bool have_rng[T];
                float bnd[T][4];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    s4[t] = (uint32_t)(spins[t] >> (w * 4)) & 15u;
                    have_rng[t] = false;
                    bnd[t][0] = bnd[t][1] = bnd[t][2] = bnd[t][3] = 0.0f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int vbase = w * 128 + k * 32;
                    if (vbase >= n) break;
                    const int rem = n - vbase;
                    const unsigned livem = rem >= 32 ? FULL : ((1u << rem) - 1u);
                    unsigned upm[T], todo[T], m[T];
                    float sgn[T];
                    bool dirty[T];
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const bool up0 = (s4[t] >> k) & 1u;
                        upm[t] = __ballot_sync(FULL, up0);
                        sgn[t] = up0 ? -2.0f : 2.0f;
                        todo[t] = livem;
                        m[t] = 0u;
                        dirty[t] = true;
                    }
                    while (true) {
                        // (re)evaluate the chains that changed; the others keep their cached ballot
                        int a = 32;
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            if (dirty[t]) {
                                const float fk = (k == 0) ? Fc[t][0].x : ((k == 1) ? Fc[t][0].y : ((k == 2) ? Fc[t][1].x : Fc[t][1].y));
                                const float dE = __fmul_rn(fk, sgn[t]);
                                if (!have_rng[t]) {
                                    const bool pend = (dE > 0.0f) && (dE < thr);
                                    if (__ballot_sync(FULL, pend) & todo[t]) {
                                        uint32_t tt = ts;
                                        asm volatile("" : "+r"(tt));        // no speculative hoisting of the draw (see sa_kernel.cu)
                                        const Philox4 o = philox4x32_10(c_lo[t], c_hi[t], tt, (uint32_t)(w * 32 + lane), k0, k1);
                                        bnd[t][0] = fminf(thr, __fdiv_rn(neg_log_u32(o.x), beta));
                                        bnd[t][1] = fminf(thr, __fdiv_rn(neg_log_u32(o.y), beta));
                                        bnd[t][2] = fminf(thr, __fdiv_rn(neg_log_u32(o.z), beta));
                                        bnd[t][3] = fminf(thr, __fdiv_rn(neg_log_u32(o.w), beta));
                                        have_rng[t] = true;
                                    }
                                }
                                const bool acc = (dE <= 0.0f) || (dE < bnd[t][k]);
                                m[t] = __ballot_sync(FULL, acc) & todo[t];
                                dirty[t] = false;
                            }
                            const int at = m[t] ? (__ffs(m[t]) - 1) : 32;
                            a = min(a, at);
                        }
                        if (a == 32) break;
                        float2 cc[T];
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            const bool f = (m[t] >> a) & 1u;            // a <= first accepted of every chain, so bit a set <=> first accepted == a
                            float c = 0.0f;
                            if (f) {
                                c = ((upm[t] >> a) & 1u) ? -2.0f : 2.0f;                     // -2 * s_a(old)
                                upm[t] ^= 1u << a;
                                if (lane == a) sgn[t] = -sgn[t];
                                todo[t] &= ~((2u << a) - 1u);
                                dirty[t] = true;
                                ++nacc[t];
                            }
                            cc[t] = make_float2(c, c);
                        }
                        const float *row = J + (size_t)(vbase + a) * (size_t)ld + lane * 4;
                        {
                            const float4 r = __ldg(reinterpret_cast<const float4 *>(row + w * 128));
#pragma unroll
                            for (int t = 0; t < T; ++t) {
                                ffma2(Fc[t][0], cc[t], make_float2(r.x, r.y));
                                ffma2(Fc[t][1], cc[t], make_float2(r.z, r.w));
                            }
                        }
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2) {
                            const float4 r = __ldg(reinterpret_cast<const float4 *>(row + w2 * 128));
#pragma unroll
                            for (int t = 0; t < T; ++t) {
                                ffma2(F2[t][w2][0], cc[t], make_float2(r.x, r.y));
                                ffma2(F2[t][w2][1], cc[t], make_float2(r.z, r.w));
                            }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < T; ++t) s4[t] = (s4[t] & ~(1u << k)) | (((upm[t] >> lane) & 1u) << k);
                }
#pragma unroll
                for (int t = 0; t < T; ++t) spins[t] = (spins[t] & ~(15ull << (w * 4))) | ((unsigned long long)s4[t] << (w * 4));
            }
        }
    }

    // ---- write-back: states in natural variable order, 0/1 ----
#pragma unroll
    for (int t = 0; t < T; ++t) {
        if (!clive[t]) continue;
        int8_t *o = p.out + (size_t)cl[t] * (size_t)n;
        for (int w = 0; w < nw_rt; ++w) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = w * 128 + k * 32 + lane;
                if (v < n) o[v] = (int8_t)((spins[t] >> (w * 4 + k)) & 1ull);
            }
        }
        if (p.counters != nullptr && lane == 0) {
            atomicAdd(p.counters + 0, (unsigned long long)nacc[t]);
            atomicAdd(p.counters + 1, (unsigned long long)n * (unsigned long long)ts);
        }
    }
}

template <int NW, int T, int WPC, int MINB>
int launch_multi(const SaParams &p, cudaStream_t st)
{
    auto kern = sa_multi_kernel<NW, T, WPC, MINB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);   // all on-chip memory as L1
    const long long gpp = (p.num_reads + T - 1) / T;
    const long long groups = gpp * p.batch_q;
    const long long blocks = (groups + WPC - 1) / WPC;
    if (blocks > 0x7fffffffLL) {
        qbm_set_error("qbm_sa_sample: too many chains for one launch (%lld)", p.total_chains);
        return QBM_EUNSUPPORTED;
    }
    kern<<<(unsigned)blocks, WPC * 32, 0, st>>>(p, gpp, groups);
    QBM_LAUNCH_OK("sa_multi_kernel");
    return QBM_OK;
}

}  // namespace

// nw = windows of the padded row (sa_variant_nw); chains per warp chosen so that the fields fit the register file
bool sa_multi_supported(int nw, long long num_reads) { return nw >= 2 && num_reads >= 2; }

int sa_multi_launch(const SaParams &p, int nw, cudaStream_t st)
{
    switch (nw) {
        case 2: return launch_multi<2, 4, 8, 1>(p, st);
        case 3: return launch_multi<3, 4, 8, 1>(p, st);
        case 4: return launch_multi<4, 4, 8, 1>(p, st);
        case 5: return launch_multi<5, 4, 8, 1>(p, st);
        case 6: return launch_multi<6, 4, 8, 1>(p, st);
        case 8: return launch_multi<8, 3, 8, 1>(p, st);
        case 10: return launch_multi<10, 2, 8, 1>(p, st);
        case 12: return launch_multi<12, 2, 8, 1>(p, st);
        default: return launch_multi<16, 2, 8, 1>(p, st);
    }
}
