// K1: simulated-annealing QUBO sampler for sm_100a -- one warp per read (chain).
//
// Replaces the Metropolis loop of dwave-neal 0.5.9 (cpu_sa.cpp) that the reference reaches through
// src/qubo/sampler.py:31-33 and src/model/faster_dqbm.py:299-313.  The rule
// is the reference's (fixed sweep order, threshold skip, dE<=0 auto-accept, u < exp(-beta dE) -- evaluated
// in the log domain, dE < -ln(u)/beta, so a proposal that is re-evaluated after a neighbour flipped costs one
// compare); the
// layout is B200-first:
//   * the chain's n local fields live in registers, 4*NW per lane (variable v = w*128 + k*32 + lane
//     is register F[w][k] of `lane`); spins are 4*NW bits per lane
//   * the coupling matrix is stored column-permuted (p128_pos) so the four fields of a lane are one
//     128-bit load per window; rows stream through L1 (read-only path): the sweep order is fixed, so
//     the chains of an SM need row v at about the same time and mostly find it L1-resident (87 % hit
//     rate at n = 2048)
//   * proposals are evaluated 32 at a time (one sub-window of 32 consecutive variables, one per lane):
//     because the uniform for (chain, sweep, v) is a pure function of its index (Philox4x32-10), the
//     first accepted proposal of the sub-window is found with one ballot, every earlier proposal is
//     a rejection that changes nothing, and evaluation restarts right after the flipped variable.
//     The trajectory is therefore exactly the sequential one (oracle/replay_sa.c).
#include "sa_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int NW>
__device__ __forceinline__ void row_update(float (&F)[NW][4], const float *__restrict__ row_lane, const float c)
{
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(row_lane + w2 * 128));
        F[w2][0] = __fmaf_rn(c, r.x, F[w2][0]);
        F[w2][1] = __fmaf_rn(c, r.y, F[w2][1]);
        F[w2][2] = __fmaf_rn(c, r.z, F[w2][2]);
        F[w2][3] = __fmaf_rn(c, r.w, F[w2][3]);
    }
}

// NW = number of 128-variable windows held per lane, KS = sub-windows evaluated per window
// (KS < 4 only for NW == 1, i.e. n <= 32 / 64), WPC = warps (chains) per CTA.
template <int NW, int KS, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB) sa_kernel(const SaParams p)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    long long cl = (long long)blockIdx.x * WPC + warp;
    const bool live = cl < p.total_chains;
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

    float F[NW][4];
    unsigned long long spins = 0ull;              // bit (w*4+k) = spin of variable w*128 + k*32 + lane (1 = up)

    // ---- initial spins and local fields: F_i = h_i ; for j = 0..n-1: F_i = fma(J[j][i], s_j, F_i) ----
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(hq + w * 128 + lane * 4));
        F[w][0] = hv.x; F[w][1] = hv.y; F[w][2] = hv.z; F[w][3] = hv.w;
    }
    for (int w = 0; w < nw_rt; ++w) {
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
            const int jend = min(32, n - jbase);
            for (int jj = 0; jj < jend; ++jj) {
                const float sj = ((wd[k] >> jj) & 1u) ? 1.0f : -1.0f;
                row_update<NW>(F, J + (uint32_t)(jbase + jj) * (uint32_t)ld, sj);
            }
        }
    }

    // ---- annealing ----
    uint32_t nacc = 0;
    uint32_t t = 0;
    for (int b = 0; b < p.num_betas; ++b) {
        const float beta = __ldg(betas + b);
        const float thr = __fdiv_rn(44.36142f, beta);
        for (int s = 0; s < p.sweeps_per_beta; ++s, ++t) {
            for (int w = 0; w < nw_rt; ++w) {
                if (rendezvous) __syncthreads();
                // working copy of this window's four fields (register index must be static); for a single
                // window the fields themselves are the working copy
                float Fc_store[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                float (&Fc)[4] = *((NW == 1) ? &F[0] : &Fc_store);
                if (NW > 1) {
#pragma unroll
                    for (int w2 = 0; w2 < NW; ++w2)
                        if (w2 == w) { Fc[0] = F[w2][0]; Fc[1] = F[w2][1]; Fc[2] = F[w2][2]; Fc[3] = F[w2][3]; }
                }
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
                    unsigned upm = __ballot_sync(FULL, up0);                 // spins of the sub-window (warp-uniform)
                    float sgn = up0 ? -2.0f : 2.0f;                          // dE = sgn * F
                    while (true) {
                        const float dE = __fmul_rn(Fc[k], sgn);
                        if (!have_rng) {
                            const bool pend = (dE > 0.0f) && (dE < thr);
                            if (__ballot_sync(FULL, pend) & todo) {
                                // keep the draw inside the branch: without the barrier the compiler speculates the (pure)
                                // Philox + log above it and every cold sweep pays ~140 instructions per window for nothing.
                                // Measured per instantiation: +7..20 % for NW <= 2 and NW >= 8, -5..12 % for NW = 3..5
                                // (where the hoisted draw overlaps load latency), neutral at NW = 6.
                                uint32_t tt = t;
                                if (NW <= 2 || NW >= 6) asm volatile("" : "+r"(tt));
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
                        const float c = ((upm >> a) & 1u) ? -2.0f : 2.0f;     // -2 * s_a(old)
                        upm ^= 1u << a;
                        if (lane == a) sgn = -sgn;
                        const float *row = J + (uint32_t)(vbase + a) * (uint32_t)ld;
                        if (NW > 1) {
                            const float4 r = __ldg(reinterpret_cast<const float4 *>(row + w * 128));
                            Fc[0] = __fmaf_rn(c, r.x, Fc[0]);
                            Fc[1] = __fmaf_rn(c, r.y, Fc[1]);
                            Fc[2] = __fmaf_rn(c, r.z, Fc[2]);
                            Fc[3] = __fmaf_rn(c, r.w, Fc[3]);
                        }
                        row_update<NW>(F, row, c);
                        todo &= ~((2u << a) - 1u);
                        ++nacc;
                    }
                    s4 = (s4 & ~(1u << k)) | (((upm >> lane) & 1u) << k);
                }
                spins = (spins & ~(15ull << (w * 4))) | ((unsigned long long)s4 << (w * 4));
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
            atomicAdd(p.counters + 1, (unsigned long long)n * (unsigned long long)t);
        }
    }
}

// column-permute + pad one batch of spin models into the workspace
__global__ void sa_permute_kernel(const float *__restrict__ J, const float *__restrict__ h, int n, int ldj, int ld,
                                  float *__restrict__ Jp, float *__restrict__ hp)
{
    // grid: (n + 1, batch_q); row index n = the h vector
    const int row = blockIdx.x;
    const size_t q = blockIdx.y;
    const float *src = (row < n) ? J + (q * (size_t)n + row) * (size_t)ldj : h + q * (size_t)n;
    float *dst = (row < n) ? Jp + (q * (size_t)n + row) * (size_t)ld : hp + q * (size_t)ld;
    for (int pos = threadIdx.x; pos < ld; pos += blockDim.x) {
        // inverse of p128_pos: storage position -> variable
        const int v = (pos & ~127) | (((pos & 3) << 5) | ((pos >> 2) & 31));
        dst[pos] = (v < n) ? src[v] : 0.0f;
    }
}

template <int NW, int KS, int WPC, int MINB>
int launch_sa(const SaParams &p, cudaStream_t st)
{
    auto kern = sa_kernel<NW, KS, WPC, MINB>;
    // all on-chip memory as L1: coupling rows are shared between the chains of an SM through L1
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    const long long blocks = (p.total_chains + WPC - 1) / WPC;
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

}  // namespace

// number of 128-variable windows of the kernel instantiation that serves n (rows are padded to it)
static inline int sa_variant_nw(int n)
{
    const int nw = (n + 127) / 128;
    if (nw <= 6) return nw;
    if (nw <= 8) return 8;
    if (nw <= 10) return 10;
    if (nw <= 12) return 12;
    return 16;
}
static inline int sa_ld(int n) { return sa_variant_nw(n) * 128; }

extern "C" QBM_API size_t qbm_sa_workspace_bytes(int n, long long batch_q)
{
    if (n <= 0 || batch_q <= 0) return 0;
    const size_t ld = (size_t)sa_ld(n);
    return (size_t)batch_q * ((size_t)n + 1) * ld * sizeof(float);
}

extern "C" QBM_API int qbm_sa_sample(const float *J, const float *h, int n, int ldj, long long batch_q,
                             const float *beta, long long beta_stride, int num_betas, int sweeps_per_beta,
                             long long num_reads, uint64_t seed, uint64_t chain_offset,
                             const int8_t *init_states, int8_t *states_out, unsigned long long *counters,
                             void *workspace, size_t workspace_bytes, unsigned flags, void *stream)
{
    QBM_CHECK_ARG(J && h && beta && states_out && workspace, "qbm_sa_sample: null pointer argument");
    QBM_CHECK_ARG(n >= 1 && ldj >= n, "qbm_sa_sample: need n >= 1 and ldj >= n (n=%d ldj=%d)", n, ldj);
    QBM_CHECK_ARG(batch_q >= 1 && num_reads >= 1, "qbm_sa_sample: batch_q and num_reads must be >= 1");
    QBM_CHECK_ARG(num_betas >= 0 && sweeps_per_beta >= 1, "qbm_sa_sample: bad schedule (num_betas=%d sweeps_per_beta=%d)",
                  num_betas, sweeps_per_beta);
    QBM_CHECK_ARG(batch_q <= 65535, "qbm_sa_sample: batch_q > 65535 not supported in one call");
    QBM_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "qbm_sa_sample: workspace must be 16-byte aligned");
    if (n > QBM_SA_MAX_N) {
        qbm_set_error("qbm_sa_sample: n=%d exceeds QBM_SA_MAX_N=%d", n, QBM_SA_MAX_N);
        return QBM_EUNSUPPORTED;
    }
    if (workspace_bytes < qbm_sa_workspace_bytes(n, batch_q)) {
        qbm_set_error("qbm_sa_sample: workspace of %zu bytes, need %zu", workspace_bytes, qbm_sa_workspace_bytes(n, batch_q));
        return QBM_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // kernel choice: the warp-per-chain kernel is the default (faster at every size measured in round 1:
    // 23.3 vs 15.4 G spin-updates/s at n = 2048); flag bit 4 selects the chain-tile kernel (sa_tile.cu)
    const bool tile = (flags & 16u) != 0u && sa_tile_supported(n);
    const int ld = tile ? sa_tile_ld(n) : sa_ld(n);
    float *Jp = reinterpret_cast<float *>(workspace);
    float *hp = Jp + (size_t)batch_q * (size_t)n * (size_t)ld;

    sa_permute_kernel<<<dim3((unsigned)n + 1, (unsigned)batch_q), 128, 0, st>>>(J, h, n, ldj, ld, Jp, hp);
    QBM_LAUNCH_OK("sa_permute_kernel");

    SaParams p;
    p.Jp = Jp; p.hp = hp; p.beta = beta; p.beta_stride = beta_stride; p.num_betas = num_betas;
    p.sweeps_per_beta = sweeps_per_beta; p.n = n; p.ld = ld; p.num_reads = num_reads;
    p.total_chains = batch_q * num_reads; p.seed = seed; p.chain_offset = chain_offset;
    p.init = init_states; p.out = states_out; p.counters = counters; p.flags = flags;
    p.Jnat = J; p.ldj = ldj; p.batch_q = batch_q;
    if (tile) return sa_tile_launch(p, st);

    const int nw = sa_variant_nw(n);
    if ((flags & 32u) && sa_multi_supported(nw, num_reads)) return sa_multi_launch(p, nw, st);
    if (n <= 32) return launch_sa<1, 1, 8, 4>(p, st);
    if (n <= 64) return launch_sa<1, 2, 8, 4>(p, st);
    switch (nw) {
        case 1: return launch_sa<1, 4, 8, 4>(p, st);
        case 2: return launch_sa<2, 4, 8, 4>(p, st);
        case 3: return launch_sa<3, 4, 8, 3>(p, st);
        case 4: return launch_sa<4, 4, 8, 3>(p, st);
        case 5: return launch_sa<5, 4, 8, 3>(p, st);
        case 6: return launch_sa<6, 4, 8, 3>(p, st);
        case 8: return launch_sa<8, 4, 8, 2>(p, st);
        case 10: return launch_sa<10, 4, 16, 1>(p, st);
        case 12: return launch_sa<12, 4, 16, 1>(p, st);
        default: return launch_sa<16, 4, 16, 1>(p, st);
    }
}
