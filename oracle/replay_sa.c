/*
 * oracle/replay_sa.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * The REPLAY oracle: a plain sequential CPU restatement of the reference's Metropolis rule
 * (dwave-neal 0.5.9 cpu_sa.cpp as pinned in SURVEY.md Appendix A.5: fixed sweep order
 * v = 0..n-1, skip when dE >= 44.36142/beta, accept when dE <= 0, otherwise accept iff
 * exp(-dE*beta) > uniform, evaluated in the log domain) evaluated in the arithmetic of the sm_100a kernel and fed the
 * kernel's own counter-based Philox4x32-10 stream (BASELINE.json north_star: "SA trajectories
 * are bit-exact against a CPU replay of the reference's Metropolis rule fed the kernel's own
 * Philox stream").  It is written independently of csrc/ (no shared headers) from the
 * trajectory specification in DESIGN.md section 3:
 *
 *   spins      s_v in {-1,+1};  x_v = (s_v+1)/2 is what the reference sees (BINARY sample)
 *   init       s_v from init01[v] when given, else bit (v&31) of word ((v>>5)&3) of
 *              philox(ctr = (chain_lo, chain_hi, 0xFFFFFFFF, v>>7), key = (seed_lo, seed_hi))
 *   fields     F_i = h_i ; for j = 0..n-1 : F_i = fmaf(J[j][i], s_j, F_i)          (fp32)
 *   sweep t    beta = betas[t / sweeps_per_beta] (fp32) ; thr = 44.36142f / beta
 *   proposal   dE = (s_v > 0 ? -2 : 2) * F_v
 *              dE >= thr -> skip ; dE <= 0 -> flip ;
 *              else u = philox(ctr = (chain_lo, chain_hi, t, ((v>>7)<<5)|(v&31)))[(v>>5)&3],
 *                   flip iff dE < fminf(thr, neg_log_u32(u) / beta)
 *              (the test u/2^32 < exp(-beta dE) in the log domain: dE < -ln(u/2^32) / beta)
 *   flip       c = (s_v > 0 ? -2 : 2) ; for all j : F_j = fmaf(c, J[v][j], F_j) ; s_v = -s_v
 *
 * Everything is IEEE binary32 with round-to-nearest-even; build with -ffp-contract=off so the
 * only fused operations are the explicit fmaf() calls.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- neg_log_u32: FMA-only -ln(u / 2^32) for a 32-bit uniform (u = 0 -> +inf) ------------------ */
float oracle_neg_log_u32(uint32_t u)
{
    if (u == 0u) return INFINITY;
    const float x = (float)u;                        /* [1, 2^32], round to nearest even */
    int32_t bits;
    memcpy(&bits, &x, sizeof bits);
    int e = (bits >> 23) - 127;
    int32_t mb = (bits & 0x007fffff) | 0x3f800000;
    float m;
    memcpy(&m, &mb, sizeof m);                       /* [1, 2) */
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
    const float f = m - 1.0f;                        /* exact */
    const float z = f * f;
    float y = 7.0376836292e-2f;                      /* Cephes logf polynomial on [sqrt(1/2)-1, sqrt(2)-1] */
    y = fmaf(y, f, -1.1514610310e-1f);
    y = fmaf(y, f, 1.1676998740e-1f);
    y = fmaf(y, f, -1.2420140846e-1f);
    y = fmaf(y, f, 1.4249322787e-1f);
    y = fmaf(y, f, -1.6668057665e-1f);
    y = fmaf(y, f, 2.0000714765e-1f);
    y = fmaf(y, f, -2.4999993993e-1f);
    y = fmaf(y, f, 3.3333331174e-1f);
    y = y * f;
    y = y * z;
    y = fmaf(-0.5f, z, y);
    const float r = f + y;                           /* ln(m) */
    const float E = (float)(32 - e);
    float nl = fmaf(E, 0.693359375f, -r);            /* ln2 high part */
    nl = fmaf(E, -2.12194440e-4f, nl);               /* ln2 low part */
    return nl;
}

static inline int philox_init_spin(uint64_t seed, uint64_t chain, int v)
{
    const uint32_t ctr[4] = { (uint32_t)chain, (uint32_t)(chain >> 32), 0xFFFFFFFFu, (uint32_t)(v >> 7) };
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    return ((o[(v >> 5) & 3] >> (v & 31)) & 1u) ? 1 : -1;
}

/*
 * J: [n, ld] fp32 symmetric spin couplings with zero diagonal; h: [n].
 * init01: nullable [num_chains, n] of 0/1.  states01_out: [num_chains, n] of 0/1.
 * counters (nullable) [3]: accepted flips, threshold skips, uniform draws (summed over chains).
 */
int oracle_replay_sa(int n, int ld, const float *J, const float *h, int num_betas, const float *betas,
                     int sweeps_per_beta, uint64_t seed, uint64_t chain_first, int num_chains,
                     const signed char *init01, signed char *states01_out, uint64_t *counters)
{
    if (n <= 0 || ld < n || num_betas < 0 || sweeps_per_beta < 1) return -2;
    float *F = (float *)malloc(sizeof(float) * (size_t)n);
    signed char *s = (signed char *)malloc((size_t)n);
    if (!F || !s) { free(F); free(s); return -1; }
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint64_t n_acc = 0, n_skip = 0, n_draw = 0;

    for (int c = 0; c < num_chains; ++c) {
        const uint64_t chain = chain_first + (uint64_t)c;
        for (int v = 0; v < n; ++v)
            s[v] = init01 ? (init01[(size_t)c * n + v] ? 1 : -1) : (signed char)philox_init_spin(seed, chain, v);
        for (int i = 0; i < n; ++i) F[i] = h[i];
        for (int j = 0; j < n; ++j) {
            const float sj = (float)s[j];
            const float *row = J + (size_t)j * ld;
            for (int i = 0; i < n; ++i) F[i] = fmaf(row[i], sj, F[i]);
        }
        uint32_t t = 0;
        for (int b = 0; b < num_betas; ++b) {
            const float beta = betas[b];
            const float thr = 44.36142f / beta;
            for (int sw = 0; sw < sweeps_per_beta; ++sw, ++t) {
                for (int v = 0; v < n; ++v) {
                    const float dE = (s[v] > 0 ? -2.0f : 2.0f) * F[v];
                    if (dE >= thr) { ++n_skip; continue; }
                    int flip;
                    if (dE <= 0.0f) {
                        flip = 1;
                    } else {
                        const uint32_t ctr[4] = { (uint32_t)chain, (uint32_t)(chain >> 32), t,
                                                  (uint32_t)(((v >> 7) << 5) | (v & 31)) };
                        uint32_t o[4];
                        oracle_philox4x32_10(ctr, key, o);
                        ++n_draw;
                        const float bound = fminf(thr, oracle_neg_log_u32(o[(v >> 5) & 3]) / beta);
                        flip = dE < bound;
                    }
                    if (flip) {
                        const float cf = (s[v] > 0 ? -2.0f : 2.0f);
                        const float *row = J + (size_t)v * ld;
                        for (int j = 0; j < n; ++j) F[j] = fmaf(cf, row[j], F[j]);
                        s[v] = (signed char)-s[v];
                        ++n_acc;
                    }
                }
            }
        }
        for (int v = 0; v < n; ++v) states01_out[(size_t)c * n + v] = (signed char)(s[v] > 0);
    }
    if (counters) { counters[0] = n_acc; counters[1] = n_skip; counters[2] = n_draw; }
    free(F); free(s);
    return 0;
}

/* QUBO energies x^T Q x in float64, the reference's energy definition:
 * sum_i Q_ii x_i + sum_{i!=j} Q_ij x_i x_j over the full matrix as given. */
int oracle_qubo_energy(int n, const double *Q, long long R, const signed char *x01, double *out)
{
    for (long long r = 0; r < R; ++r) {
        const signed char *x = x01 + (size_t)r * n;
        double e = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!x[i]) continue;
            const double *row = Q + (size_t)i * n;
            double acc = 0.0;
            for (int j = 0; j < n; ++j) if (x[j]) acc += row[j];
            e += acc;
        }
        out[r] = e;
    }
    return 0;
}
