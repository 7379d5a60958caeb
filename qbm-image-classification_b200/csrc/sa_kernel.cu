// K1 dispatch: the warp-per-chain sampler (sa_warp.cuh) and its column-permuted workspace; entry point qbm_sa_sample.
#include "sa_warp.cuh"

using sa_warp::launch_sa;

namespace {

// column-permute + pad one batch of spin models into the workspace
// (rot > 0: row r is stored rotated by its own window, segment j holds window (r / 128 + j) % rot -- the layout of the
// rotating-window kernels)
__global__ void sa_permute_kernel(const float *__restrict__ J, const float *__restrict__ h, int n, int ldj, int ld, int rot,
                                  float *__restrict__ Jp, float *__restrict__ hp)
{
    // grid: (n + 1, batch_q); row index n = the h vector
    const int row = blockIdx.x;
    const size_t q = blockIdx.y;
    const float *src = (row < n) ? J + (q * (size_t)n + row) * (size_t)ldj : h + q * (size_t)n;
    float *dst = (row < n) ? Jp + (q * (size_t)n + row) * (size_t)ld : hp + q * (size_t)ld;
    for (int pos = threadIdx.x; pos < ld; pos += blockDim.x) {
        // inverse of p128_pos: storage position -> variable
        const int seg = pos >> 7;
        const int win = (rot > 0 && row < n) ? ((row >> 7) + seg) % rot : seg;
        const int v = (win << 7) | (((pos & 3) << 5) | ((pos >> 2) & 31));
        dst[pos] = (v < n) ? src[v] : 0.0f;
    }
}

}  // namespace

// number of 128-variable windows of the instantiation that serves n (rows are padded to it): the chains-per-warp kernel
// (and the workspace size, which covers every kernel) ...
static inline int sa_multi_nw(int n)
{
    const int nw = (n + 127) / 128;
    if (nw <= 6) return nw;
    if (nw <= 8) return 8;
    if (nw <= 10) return 10;
    if (nw <= 12) return 12;
    return 16;
}
// ... and the warp-per-chain kernel
static inline int sa_variant_nw(int n)
{
    const int nw = (n + 127) / 128;
    return (nw == 13 || nw == 14) ? 14 : sa_multi_nw(n);
}

extern "C" QBM_API size_t qbm_sa_workspace_bytes(int n, long long batch_q)
{
    if (n <= 0 || batch_q <= 0) return 0;
    const size_t ld = (size_t)sa_multi_nw(n) * 128;
    return (size_t)batch_q * ((size_t)n + 1) * ld * sizeof(float);
}

// two-phase schedule (chain-tile kernel for the hot sweeps, then one warp per chain): possible where both kernels read the
// same permuted rows (3..6 and 8..16 windows), used by default where it measured faster (profiles/r2_two_phase_sizes*.log)
static inline bool sa_two_phase_supported(int n) { return n > 256 && n <= QBM_SA_MAX_N && (n + 127) / 128 != 7; }
static inline bool sa_two_phase_size(int n) { return n > QBM_TWO_PHASE_MIN_N; }

extern "C" QBM_API size_t qbm_sa_workspace_bytes_two_phase(int n, long long batch_q, long long num_reads)
{
    const size_t base = qbm_sa_workspace_bytes(n, batch_q);
    if (base == 0 || num_reads <= 0 || !sa_two_phase_supported(n)) return base;
    const size_t chains = (size_t)batch_q * (size_t)num_reads;
    return base + chains * (size_t)sa_multi_nw(n) * 128 * sizeof(float) + ((chains * sizeof(uint32_t) + 15) / 16) * 16;
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
    const bool multi = !tile && (flags & 32u) != 0u && sa_multi_supported(sa_multi_nw(n), num_reads);
    const int nw = multi ? sa_multi_nw(n) : sa_variant_nw(n);
    const int ld = tile ? sa_tile_ld(n) : nw * 128;
    float *Jp = reinterpret_cast<float *>(workspace);
    float *hp = Jp + (size_t)batch_q * (size_t)n * (size_t)ld;

    // two-phase schedule: in the hot sweeps nearly every chain flips nearly every variable, so the chain-tile kernel (one
    // fetch of a coupling row for 16 chains, 35 instead of 86 clocks per flip and SM at n = 2048) anneals them and hands every
    // chain over after the first sweep that accepts less than hot_fraction of its proposals; the warp-per-chain kernel
    // resumes.  Both kernels follow the same trajectory, so the cut does not change any result.  Needs the larger workspace
    // (qbm_sa_workspace_bytes_two_phase); flag bit 6 switches it off, flag bit 7 forces it wherever it is supported.
    const bool two_phase = !tile && !multi && sa_two_phase_supported(n) && (flags & 64u) == 0u &&
                           (sa_two_phase_size(n) || (flags & 128u) != 0u) && sa_tile_ld(n) == ld &&
                           workspace_bytes >= qbm_sa_workspace_bytes_two_phase(n, batch_q, num_reads);
    // 8..14 windows run the rotating-window shape, whose rows are stored rotated (sa_warp.cuh); the chain-tile kernel reads
    // plain rows, so the two-phase schedule resumes with the plain shape
    const bool rt = !tile && !multi && !two_phase && nw >= 8 && nw <= 14;
    sa_permute_kernel<<<dim3((unsigned)n + 1, (unsigned)batch_q), 128, 0, st>>>(J, h, n, ldj, ld, rt ? nw : 0, Jp, hp);
    QBM_LAUNCH_OK("sa_permute_kernel");

    SaParams p;
    p.Jp = Jp; p.hp = hp; p.beta = beta; p.beta_stride = beta_stride; p.num_betas = num_betas;
    p.sweeps_per_beta = sweeps_per_beta; p.n = n; p.ld = ld; p.num_reads = num_reads;
    p.total_chains = batch_q * num_reads; p.seed = seed; p.chain_offset = chain_offset;
    p.init = init_states; p.out = states_out; p.counters = counters; p.flags = flags;
    p.Jnat = J; p.ldj = ldj; p.batch_q = batch_q;
    p.fields = nullptr; p.sweeps_done = nullptr; p.hot_fraction = 0.0f;
    if (tile) return sa_tile_launch(p, st);

    if (two_phase) {
        const unsigned pct = (flags >> 16) & 0xffu;
        SaParams hot = p;
        hot.fields = hp + (size_t)batch_q * (size_t)ld;
        hot.sweeps_done = reinterpret_cast<uint32_t *>(hot.fields + (size_t)p.total_chains * (size_t)ld);
        hot.hot_fraction = pct ? (float)pct / 100.0f : 0.50f;
        if (const int rc = sa_tile_launch(hot, st)) return rc;
        // the resuming instantiation (RS) starts from fields / sweeps_done / init instead of computing the initial fields
        p.fields = hot.fields; p.sweeps_done = hot.sweeps_done;
        p.init = states_out;
        switch (nw) {
            // (the code shapes of the plain instantiations: unrolled windows, packed FMAs, shuffled coefficient)
            // 3..6 windows: opt-in (flag bit 7).  Measured (profiles/r2_two_phase_small_sizes.log): +8..21 % where n fills its
            // windows and a problem has many reads (n = 384 / 512 / 640 / 768 at 200 reads), -7..9 % at the training shapes
            // (n = 522: 18 % padding; 100 reads = 6.25 tiles of 16 chains), so the default stays one warp per chain there
            case 3: return launch_sa<3, 4, 16, 2, true, true, true, true, false, true>(p, st);
            case 4: return launch_sa<4, 4, 16, 2, true, true, false, true, false, true>(p, st);
            case 5: return launch_sa<5, 4, 16, 1, true, true, true, true, false, true>(p, st);
            case 6: return launch_sa<6, 4, 16, 1, true, false, true, true, false, true>(p, st);
            case 8: return launch_sa<8, 4, 16, 1, false, false, true, false, false, true>(p, st);
            case 10: return launch_sa<10, 4, 16, 1, false, false, true, false, false, true>(p, st);
            case 12: return launch_sa<12, 4, 16, 1, false, false, true, false, false, true>(p, st);
            case 14: return launch_sa<14, 4, 16, 1, false, false, true, false, false, true>(p, st);
            default: return launch_sa<16, 4, 16, 1, false, false, true, false, false, true>(p, st);
        }
    }

    if (multi) return sa_multi_launch(p, nw, st);
    // code shape per instantiation <NW, KS, WPC, MINB, UW, P2, PIN, SH, RT>, each chosen by measurement
    // (profiles/r1d_sa_kernel_*_probe.log): unrolled windows + shuffled coefficient up to 6 windows (+13..32 %), packed FMAs
    // at 3..5 windows; at 8..14 windows the rotating-window shape removes the working copy instead (+8..14 %); at 16 windows
    // (n = 2048, L1-bandwidth-bound with 128 registers) every one of them measured 0..-2 %.  Also measured and not adopted:
    // more registers per thread for fewer resident warps (only 6 windows gain: 2 instead of 3 CTAs per SM), rows padded to 32
    // instead of 128 variables (a float / float2 tail load per row; the time of a flip followed the number of windows, not
    // the bytes), all loads of a row issued before its first FMA, a grid numbering that gives the CTAs of an SM consecutive
    // chains of one problem (r1d_sa_kernel_sm_affine_grid_probe.log: -9..+8 %)
    if (n <= 32) return launch_sa<1, 1, 8, 4, false, false, true, true>(p, st);
    if (n <= 64) return launch_sa<1, 2, 8, 4, false, false, true, true>(p, st);
    switch (nw) {
        case 1: return launch_sa<1, 4, 8, 4, false, false, true, true>(p, st);
        case 2: return launch_sa<2, 4, 8, 4, true, false, true, true>(p, st);
        case 3: return launch_sa<3, 4, 8, 3, true, true, true, true>(p, st);
        case 4: return launch_sa<4, 4, 8, 3, true, true, false, true>(p, st);
        case 5: return launch_sa<5, 4, 8, 3, true, true, true, true>(p, st);
        case 6: return launch_sa<6, 4, 8, 2, true, false, true, true>(p, st);
        case 8: return launch_sa<8, 4, 8, 2, false, false, true, false, true>(p, st);
        case 10: return launch_sa<10, 4, 16, 1, false, false, true, false, true>(p, st);
        case 12: return launch_sa<12, 4, 16, 1, false, false, true, false, true>(p, st);
        case 14: return launch_sa<14, 4, 16, 1, false, false, true, false, true>(p, st);
        default: return launch_sa<16, 4, 16, 1>(p, st);
    }
}
