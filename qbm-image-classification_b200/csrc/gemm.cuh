// Epilogue description of the tcgen05 TF32 GEMM (gemm_tcgen05.cu), shared with rbm.cu.
#pragma once
#include "common.cuh"

struct EpiParams {
    float *C; long long ldc;              // nullable: main output
    float *Ct; long long ldct;            // nullable: transposed copy  Ct[n][m]
    float *S; long long lds;              // nullable: Bernoulli(C) sample as 0/1 floats
    float *St; long long ldst;            // nullable: transposed sample
    const float *bias_n;                  // nullable: [N]
    const float *rowtab; const int *ridx; long long ldtab;   // nullable: += rowtab[ridx[m]][n]
    const float *Cin; long long ldcin;    // nullable: += beta * Cin[m][n]   (may alias C)
    float alpha, beta;
    int act;                              // 0 = identity, 1 = sigmoid
    unsigned long long seed; unsigned int stream;   // Philox key / stream id for the sample
    const unsigned int *step_dev;         // nullable: device-resident step counter, stream id += 4 * *step_dev (CUDA-graph replays)
};

// C[M,N] = epi(alpha * A[M,K] . B[N,K]^T), both operands row-major with K contiguous
int qbm_gemm_tf32_launch(const float *A, long long lda, const float *B, long long ldb, int M, int N, int K,
                         const EpiParams &ep, cudaStream_t st);
