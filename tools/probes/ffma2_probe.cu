// Stand-alone probe (not part of the library): FFMA2 issue rate for the chain-tile kernel's row-update shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe.bin ffma2_probe.cu && ./ffma2_probe.bin
// acc[t][j] = fma(c[t], r[j], acc[t][j]); NA packed accumulators per thread = 16 chains x NJ column pairs;
// SCALAR: c as a broadcast 32-bit operand (mov.b64 {c, c}), else a 64-bit register pair; WARPS resident warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int NJ, bool SCALAR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float *out, int iters, float a, float b)
{
    unsigned long long acc[16 * NJ], r[NJ];
    float cs[16];
    unsigned long long cp[16];
#pragma unroll
    for (int i = 0; i < 16 * NJ; ++i) acc[i] = ((unsigned long long)__float_as_uint((float)i) << 32) | __float_as_uint((float)threadIdx.x);
#pragma unroll
    for (int i = 0; i < 16; ++i) { cs[i] = a + i; cp[i] = ((unsigned long long)__float_as_uint(a + i) << 32) | __float_as_uint(a - i); }
#pragma unroll
    for (int i = 0; i < NJ; ++i) r[i] = ((unsigned long long)__float_as_uint(b * i) << 32) | __float_as_uint(b + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            unsigned long long c;
            if (SCALAR) asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(cs[t]));
            else c = cp[t];
#pragma unroll
            for (int j = 0; j < NJ; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[NJ * t + j]) : "l"(c), "l"(r[j]));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16 * NJ; ++i) s ^= acc[i];
    if (s == 0x123456789abcdefull) out[0] = 1.0f;
}

template <int NJ, bool SCALAR, int WARPS>
void run(const char *name, int sms, float *d)
{
    const int iters = 1 << 12;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<NJ, SCALAR, WARPS><<<sms, WARPS * 32>>>(d, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double flops = (double)sms * WARPS * 32 * iters * 16 * NJ * 2 * 2.0;
    // clocks per FFMA2 per scheduler (4 per SM) at 1.965 GHz
    const double clk = best * 1e-3 * 1.965e9 / ((double)iters * 16 * NJ * WARPS / 4);
    printf("%-44s %7.2f TFLOP/s  %5.2f clk per FFMA2 per scheduler  (%s)\n", name, flops / (best * 1e-3) * 1e-12, clk, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d; cudaMalloc(&d, 64);
    run<4, false, 8>("8 warps, 64 acc, pair coefficient", sms, d);
    run<4, true, 8>("8 warps, 64 acc, scalar coefficient", sms, d);
    run<4, true, 4>("4 warps, 64 acc, scalar coefficient", sms, d);
    run<2, false, 16>("16 warps, 32 acc, pair coefficient", sms, d);
    run<2, true, 16>("16 warps, 32 acc, scalar coefficient", sms, d);
    run<2, true, 8>("8 warps, 32 acc, scalar coefficient", sms, d);
    run<2, true, 12>("12 warps, 32 acc, scalar coefficient", sms, d);
    run<1, true, 16>("16 warps, 16 acc, scalar coefficient", sms, d);
    run<1, true, 32>("32 warps, 16 acc, scalar coefficient", sms, d);
    return 0;
}
