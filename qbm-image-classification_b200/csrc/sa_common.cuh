// Shared declarations of the two SA sampler kernels (sa_kernel.cu: one warp per chain; sa_tile.cu: a
// register tile of T chains per CTA with coupling rows streamed through a TMA ring).
#pragma once
#include "common.cuh"

struct SaParams {
    const float *Jp;          // [batch_q, n, ld]  columns in p128 order, rows zero-padded to ld
    const float *hp;          // [batch_q, ld]     p128 order, zero-padded
    const float *Jnat;        // [batch_q, n, ldj] the caller's natural-order couplings (diagonal blocks of the tile kernel)
    int ldj;
    const float *beta;        // [batch_q or 1, num_betas]
    long long beta_stride;
    int num_betas;
    int sweeps_per_beta;
    int n;
    int ld;
    long long num_reads;
    long long total_chains;   // batch_q * num_reads
    unsigned long long seed;
    unsigned long long chain_offset;
    const int8_t *init;       // nullable [total_chains, n]
    int8_t *out;              // [total_chains, n]
    unsigned long long *counters;
    unsigned flags;
    long long batch_q;
    // two-phase schedule (qbm_sa_sample with a large workspace): the chain-tile kernel anneals the hot sweeps and hands its
    // chains over -- fields in the warp kernel's register layout, sweeps done per chain, spins through `out` -- to the
    // warp-per-chain kernel, which resumes them (same trajectory)
    float *fields;            // nullable [total_chains, ld]
    uint32_t *sweeps_done;    // nullable [tiles]: completed sweeps per tile of 16 chains
    float hot_fraction;       // tile kernel: stop after the first sweep whose accepted fraction is below this (0 = never)
};

// sa_tile.cu
bool sa_tile_supported(int n);
int sa_tile_ld(int n);
int sa_tile_launch(const SaParams &p, cudaStream_t st);

// sa_multi.cu
bool sa_multi_supported(int nw, long long num_reads);
int sa_multi_launch(const SaParams &p, int nw, cudaStream_t st);
