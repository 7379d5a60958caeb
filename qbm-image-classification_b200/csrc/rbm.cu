// K5 + step orchestration for the ClassificationRBM (src/ClassificationRBM.py:43-157).
//
// The dense contractions (x.W, h.W^T, x^T.D) run on the tcgen05 TF32 GEMM (gemm_tcgen05.cu) with
// bias / class-bias / sigmoid / Bernoulli / SGD-accumulate fused into its epilogue; everything that is
// O(B*H*C) (softplus sums, p(y|x), the positive-minus-negative phase difference, class-weight and bias
// gradients) is fused elementwise work here.  Storage contract: every matrix is row-major with its
// leading dimension rounded up to a multiple of 4 floats (16-byte rows for TMA): ld(X) = (cols+3)&~3.
#include "gemm.cuh"

namespace {

__host__ __device__ inline long long ld4(long long cols) { return (cols + 3) & ~3LL; }
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float softplus(float z) { return fmaxf(z, 0.0f) + log1pf(__expf(-fabsf(z))); }

constexpr int MAXC = 32;
constexpr int DSL = 8;      // batch slices of the fused discriminative gradient kernel

// p(y|x) per row (ClassificationRBM.py:62-86).  With xin / xt the block also writes row b of the minibatch as column b of x^T (the K-major operand of the W gradient).
__global__ void __launch_bounds__(256) rbm_rows_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ U,
                                                      long long ldu, const float *__restrict__ b_c, const int *__restrict__ y,
                                                      int H, int C, float *__restrict__ P, long long ldp,
                                                      const float *__restrict__ xin = nullptr, long long ldx = 0, int V = 0,
                                                      float *__restrict__ xt = nullptr, long long ldxt = 0,
                                                      unsigned int *__restrict__ tick = nullptr, int nticks = 0)
{
    __shared__ float red[MAXC][8];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *Ab = A + (size_t)b * lda;
    if (b == 0 && tick != nullptr)          // arrival counters of the gradient kernel that follows
        for (int i = tid; i < nticks; i += 256) tick[i] = 0u;
    if (xt != nullptr)
        for (int v = tid; v < V; v += 256) xt[(size_t)v * ldxt + b] = xin[(size_t)b * ldx + v];
    // class-outer loop: one scalar accumulator and ONE inlined softplus (a class-unrolled body is 32 copies of log1pf and
    // spends the launch fetching instructions); the row of A stays in L1 across the classes
    for (int c = 0; c < C; ++c) {
        const float *Uc = U + (size_t)c * ldu;
        float v = 0.0f;
        for (int h = tid; h < H; h += 256) v += softplus(Ab[h] + __ldg(Uc + h));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[c][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
        // softmax over the classes, lane = class; sums in the fixed order of the serial loop they replace
        float v = -INFINITY;
        if (lane < C) {
            v = b_c[lane];
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red[lane][w];
        }
        float mx = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e = lane < C ? __expf(v - mx) : 0.0f;
        float z = 0.0f;
        for (int c = 0; c < C; ++c) z += __shfl_sync(0xffffffffu, e, c);
        if (lane < C) P[(size_t)b * ldp + lane] = e / z;
    }
}

// class-bias update, visible-bias decay, loss (CrossEntropyLoss applied to probabilities, :142) and argmax: one block
__device__ __forceinline__ void rbm_disc_finish(float *__restrict__ b_c, float *__restrict__ b_v, const float *__restrict__ P,
                                                long long ldp, const int *__restrict__ y, int B, int C, int V, float scale,
                                                float sparse, int *__restrict__ pred, float *__restrict__ loss, int update,
                                                float *__restrict__ gbc)
{
    __shared__ float redl[256];
    const int tid = threadIdx.x;
    float l = 0.0f;
    for (int b = tid; b < B; b += blockDim.x) {
        const float *p = P + (size_t)b * ldp;
        float z = 0.0f, best = -1.0f; int arg = 0;
        for (int c = 0; c < C; ++c) { z += __expf(p[c]); if (p[c] > best) { best = p[c]; arg = c; } }
        l += __logf(z) - p[y[b]];
        if (pred != nullptr) pred[b] = arg;
    }
    redl[tid] = l;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) { if (tid < o) redl[tid] += redl[tid + o]; __syncthreads(); }
    if (tid == 0 && loss != nullptr) loss[0] = (update == 2) ? redl[0] : redl[0] / (float)B;     // gradient mode: the sum
    if (!update) return;
    // class-bias gradient: warp w sums class w, w + 8, ... over the batch (lanes stride over b, fixed shuffle tree)
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = warp; c < C; c += 8) {
        float g = 0.0f;
        for (int b = lane; b < B; b += 32) g += (y[b] == c ? 1.0f : 0.0f) - P[(size_t)b * ldp + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == 0) {
            if (update == 2) gbc[c] = g;
            else b_c[c] = b_c[c] + scale * g - sparse;
        }
    }
    if (update != 2 && sparse != 0.0f)
        for (int v = tid; v < V; v += blockDim.x) b_v[v] -= sparse;
}

// The O(B H C) part of the discriminative step after p(y|x) (:106-136, :88-99): the phase difference
//   D[b,h] = o[b,h,y_b] - sum_c p[b,c] o[b,h,c],  o = sigmoid(A[b,h] + U[c,h])           (stored transposed: K-major for x^T.D)
// and the class-weight / hidden-bias gradients  dU[c,h] = sum_b (1[y_b = c] - p[b,c]) o[b,h,c],  db_h[h] = sum_b D[b,h],
// every sigmoid evaluated once.  Grid (h-tiles of 32, DSL batch slices + 1): a block sums its slice (warps over rows, lanes
// over h, rows and warps combined in a fixed order), writes the partial sums, and the last block of an h-tile to arrive adds
// the DSL partials in slice order and applies the update -- deterministic whichever block that is.  The extra row of
// blocks (blockIdx.y == DSL) does the O(B C) rest of the step (rbm_disc_finish).
__global__ void __launch_bounds__(256) rbm_disc_grad_kernel(const float *__restrict__ A, long long lda, float *__restrict__ U,
                                                           long long ldu, float *__restrict__ b_h, const float *__restrict__ P,
                                                           long long ldp, const int *__restrict__ y, float *__restrict__ Dt,
                                                           long long lddt, int B, int H, int C, float scale, float sparse,
                                                           float *__restrict__ gU, float *__restrict__ gbh,
                                                           float *__restrict__ part, unsigned int *__restrict__ tick,
                                                           float *__restrict__ b_c, float *__restrict__ b_v, int V,
                                                           int *__restrict__ pred, float *__restrict__ loss, int fin_update,
                                                           float *__restrict__ gbc)
{
    if ((int)blockIdx.y == DSL) {
        if (blockIdx.x == 0) rbm_disc_finish(b_c, b_v, P, ldp, y, B, C, V, scale, sparse, pred, loss, fin_update, gbc);
        return;
    }
    __shared__ float red[MAXC + 1][8][32];
    __shared__ float pr[32][MAXC];
    __shared__ float dtile[32][33];
    __shared__ int yr[32];
    __shared__ unsigned int last;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ht = blockIdx.x, sl = blockIdx.y;
    const int h = ht * 32 + lane;
    const bool ok = h < H;
    const long long lH = ldu;                                  // partial sums use the leading dimension of U
    const int SB = (B + DSL - 1) / DSL;
    const int bbeg = sl * SB, bend = min(B, bbeg + SB);
    float u[MAXC], gu[MAXC], gb = 0.0f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { u[c] = (ok && c < C) ? U[(size_t)c * ldu + h] : 0.0f; gu[c] = 0.0f; }
    for (int bb = bbeg; bb < bend; bb += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * C; i += 256) {
            const int r = i / C, c = i - r * C;
            pr[r][c] = (bb + r < bend) ? P[(size_t)(bb + r) * ldp + c] : 0.0f;
        }
        if (tid < 32) yr[tid] = (bb + tid < bend) ? y[bb + tid] : -1;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = w + 8 * k, b = bb + r;
            float d = 0.0f;
            if (ok && b < bend) {
                const float a = A[(size_t)b * lda + h];
                const int yb = yr[r];
                float pos = 0.0f, neg = 0.0f;
#pragma unroll
                for (int c = 0; c < MAXC; ++c) {
                    if (c < C) {
                        const float o = sigm(a + u[c]);
                        const float p = pr[r][c];
                        neg += p * o;
                        if (c == yb) pos = o;
                        gu[c] += (c == yb ? o : 0.0f) - p * o;
                    }
                }
                d = pos - neg;
                gb += d;
            }
            dtile[r][lane] = d;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {                          // transposed tile: lanes along the batch
            const int hh = ht * 32 + w + 8 * k, b = bb + lane;
            if (hh < H && b < bend) Dt[(size_t)hh * lddt + b] = dtile[lane][w + 8 * k];
        }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) red[c][w][lane] = gu[c];
    red[C][w][lane] = gb;
    __syncthreads();
    for (int i = tid; i < (C + 1) * 32; i += 256) {
        const int c = i >> 5, hx = i & 31;
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[c][k][hx];
        if (ht * 32 + hx < H) part[((size_t)sl * (MAXC + 1) + c) * lH + ht * 32 + hx] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(tick + ht, 1u) == (unsigned)(DSL - 1)) ? 1u : 0u;
    __syncthreads();
    if (last == 0u) return;
    __threadfence();
    for (int i = tid; i < (C + 1) * 32; i += 256) {
        const int c = i >> 5, hh = ht * 32 + (i & 31);
        if (hh >= H) continue;
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < DSL; ++k) sum += __ldcg(part + ((size_t)k * (MAXC + 1) + c) * lH + hh);
        if (c < C) {
            if (gU != nullptr) gU[(size_t)c * ldu + hh] = sum;             // gradient mode (data-parallel steps): the raw sums
            else U[(size_t)c * ldu + hh] = U[(size_t)c * ldu + hh] + scale * sum;
        } else {
            if (gU != nullptr) gbh[hh] = sum;
            else b_h[hh] = b_h[hh] + scale * sum - sparse;
        }
    }
}

// p(y|h) = exp(h.U^T + b_c) L1-normalised (:54-60) and, when y1 != null, a categorical sample of it
__global__ void __launch_bounds__(128) rbm_class_kernel(const float *__restrict__ Hm, long long ldh, const float *__restrict__ U,
                                                       long long ldu, const float *__restrict__ b_c, int H, int C,
                                                       float *__restrict__ P, long long ldp, int *__restrict__ y1,
                                                       unsigned long long seed, unsigned int stream,
                                                       const unsigned int *__restrict__ step_dev = nullptr,
                                                       const float *__restrict__ xin = nullptr, long long ldx = 0, int V = 0,
                                                       float *__restrict__ xt = nullptr, long long ldxt = 0)
{
    __shared__ float red[MAXC][4];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (xt != nullptr)          // row b of the minibatch as column b of v0^T (the K-major operand of the W gradient)
        for (int v = tid; v < V; v += 128) xt[(size_t)v * ldxt + b] = xin[(size_t)b * ldx + v];
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.0f;
    for (int h = tid; h < H; h += 128) {
        const float hv = Hm[(size_t)b * ldh + h];
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) acc[c] += hv * __ldg(U + (size_t)c * ldu + h);
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
            float v = acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[c][warp] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float e[MAXC], z = 0.0f;
        for (int c = 0; c < C; ++c) { e[c] = __expf(red[c][0] + red[c][1] + red[c][2] + red[c][3] + b_c[c]); z += e[c]; }
        z = fmaxf(z, 1e-12f);                                   // torch.nn.functional.normalize eps
        float cum = 0.0f; int pick = C - 1; bool done = false;
        const unsigned int strm = stream + (step_dev != nullptr ? 4u * __ldg(step_dev) : 0u);
        const Philox4 u4 = philox4x32_10((uint32_t)b, 0u, strm, 0x434C53u, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u = (float)(u4.x >> 8) * 5.9604644775390625e-8f;
        for (int c = 0; c < C; ++c) {
            const float p = e[c] / z;
            if (P != nullptr) P[(size_t)b * ldp + c] = p;
            cum += p;
            if (!done && u < cum) { pick = c; done = true; }
        }
        if (y1 != nullptr) y1[b] = pick;
    }
}

// CD-1 gradients of the small parameters from transposed activations ([., B] rows are contiguous over b):
// one warp per visible / hidden unit, lanes stride over the batch (coalesced), shuffle reduction.
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) rbm_cd_small_update_kernel(const float *__restrict__ v0t, const float *__restrict__ v1t,
                                                                 long long ldvt, const float *__restrict__ p0t,
                                                                 const float *__restrict__ p1t, long long ldpt,
                                                                 const int *__restrict__ y0, const int *__restrict__ y1,
                                                                 float *__restrict__ U, long long ldu, float *__restrict__ b_v,
                                                                 float *__restrict__ b_h, float *__restrict__ b_c, int B, int V,
                                                                 int H, int C, float scale, float sparse, int grad_mode)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i < V) {
        float g = 0.0f;
        for (int b = lane; b < B; b += 32) g += v0t[(size_t)i * ldvt + b] - v1t[(size_t)i * ldvt + b];
        g = warp_sum(g);
        if (lane == 0) b_v[i] = grad_mode ? g : b_v[i] + scale * g - sparse;
    }
    if (i < H) {
        float gh = 0.0f, gu[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) gu[c] = 0.0f;
        for (int b = lane; b < B; b += 32) {
            const float a0 = p0t[(size_t)i * ldpt + b], a1 = p1t[(size_t)i * ldpt + b];
            const int c0 = y0[b], c1 = y1[b];
            gh += a0 - a1;
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
                if (c < C) gu[c] += (c0 == c ? a0 : 0.0f) - (c1 == c ? a1 : 0.0f);
        }
        gh = warp_sum(gh);
        if (lane == 0) b_h[i] = grad_mode ? gh : b_h[i] + scale * gh - sparse;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < C) {
                const float t = warp_sum(gu[c]);
                if (lane == 0) U[(size_t)c * ldu + i] = grad_mode ? t : U[(size_t)c * ldu + i] + scale * t;
            }
        }
    }
    if (i < C) {
        float g = 0.0f;
        for (int b = lane; b < B; b += 32) g += (y0[b] == i ? 1.0f : 0.0f) - (y1[b] == i ? 1.0f : 0.0f);
        g = warp_sum(g);
        if (lane == 0) b_c[i] = grad_mode ? g : b_c[i] + scale * g - sparse;
    }
}

struct Ws {   // carve-up of the caller's workspace (floats)
    float *A, *P, *Dt, *xt, *p0, *p0t, *h0, *v1, *v1t, *p1t, *pc;
    int *y1;
    float *part;             // [DSL][MAXC + 1][ld4(H)] per-slice partial sums of the discriminative gradients
    unsigned int *tick;      // [(H + 31) / 32] arrival counters of the slices of an h-tile
};

size_t ws_floats(int B, int V, int H, int C)
{
    const size_t lB = ld4(B), lV = ld4(V), lH = ld4(H), lC = ld4(C);
    // A[B,lH] P[B,lC] Dt[H,lB] xt[V,lB] p0[B,lH] p0t[H,lB] h0[B,lH] v1[B,lV] v1t[V,lB] p1t[H,lB] pc[B,lC] y1[B]
    return (size_t)B * lH * 3 + (size_t)B * lC * 2 + (size_t)H * lB * 3 + (size_t)V * lB * 2 + (size_t)B * lV + lB + 64 +
           (size_t)DSL * (MAXC + 1) * lH + (((size_t)H + 31) / 32 + 3) / 4 * 4;
}

Ws carve(void *workspace, int B, int V, int H, int C)
{
    const size_t lB = ld4(B), lV = ld4(V), lH = ld4(H), lC = ld4(C);
    float *p = reinterpret_cast<float *>(workspace);
    Ws w;
    auto take = [&](size_t n) { float *r = p; p += (n + 3) & ~size_t(3); return r; };
    w.A = take((size_t)B * lH); w.P = take((size_t)B * lC); w.Dt = take((size_t)H * lB); w.xt = take((size_t)V * lB);
    w.p0 = take((size_t)B * lH); w.p0t = take((size_t)H * lB); w.h0 = take((size_t)B * lH); w.v1 = take((size_t)B * lV);
    w.v1t = take((size_t)V * lB); w.p1t = take((size_t)H * lB); w.pc = take((size_t)B * lC);
    w.y1 = reinterpret_cast<int *>(take(lB));
    w.part = take((size_t)DSL * (MAXC + 1) * lH);
    w.tick = reinterpret_cast<unsigned int *>(take(((size_t)H + 31) / 32));
    return w;
}

// ---- data-parallel steps: gradients in ONE flat buffer (all-reduced as it is), then one fused apply --------------
// layout (floats; every segment starts at a multiple of 4): gW [V, ld4(H)] | gU [C, ld4(H)] | gb_v [V] | gb_h [H] | gb_c [C] | loss sum [1]
struct Grad {
    float *gW, *gU, *gbv, *gbh, *gbc, *loss;
};
size_t grad_floats(int V, int H, int C)
{
    const size_t lH = ld4(H);
    return (size_t)V * lH + (size_t)C * lH + ld4(V) + ld4(H) + ld4(C) + 4;
}
Grad grad_carve(float *g, int V, int H, int C)
{
    const size_t lH = ld4(H);
    Grad r;
    r.gW = g; g += (size_t)V * lH;
    r.gU = g; g += (size_t)C * lH;
    r.gbv = g; g += ld4(V);
    r.gbh = g; g += ld4(H);
    r.gbc = g; g += ld4(C);
    r.loss = g;
    return r;
}

// W += scale gW and W^T refreshed in the same pass (32 x 32 tiles through shared memory); the blocks past the W tiles
// apply the small parameters (param += scale g, the three biases then -= sparse: ClassificationRBM.py:88-99)
__global__ void __launch_bounds__(256) rbm_apply_kernel(float *__restrict__ W, float *__restrict__ Wt, float *__restrict__ U,
                                                       float *__restrict__ b_v, float *__restrict__ b_h, float *__restrict__ b_c,
                                                       const float *__restrict__ gW, const float *__restrict__ gU,
                                                       const float *__restrict__ gbv, const float *__restrict__ gbh,
                                                       const float *__restrict__ gbc, const float *__restrict__ gloss, int V, int H,
                                                       int C, long long lH, long long lV, float scale, float sparse,
                                                       float *__restrict__ loss_out, float loss_scale, int tiles_x, int tiles)
{
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if ((int)blockIdx.x < tiles) {
        const int h0 = ((int)blockIdx.x % tiles_x) * 32, v0 = ((int)blockIdx.x / tiles_x) * 32;
        for (int i = ty; i < 32; i += 8) {
            const int v = v0 + i, h = h0 + tx;
            float w = 0.0f;
            if (v < V && h < H) {
                w = W[(size_t)v * lH + h] + scale * gW[(size_t)v * lH + h];
                W[(size_t)v * lH + h] = w;
            }
            tile[i][tx] = w;
        }
        __syncthreads();
        for (int i = ty; i < 32; i += 8) {
            const int h = h0 + i, v = v0 + tx;
            if (h < H && v < V) Wt[(size_t)h * lV + v] = tile[tx][i];
        }
        return;
    }
    const int e0 = ((int)blockIdx.x - tiles) * 256 + threadIdx.x;
    const int stride = ((int)gridDim.x - tiles) * 256;
    for (int e = e0; e < C * H; e += stride) {
        const int c = e / H, h = e % H;
        U[(size_t)c * lH + h] += scale * gU[(size_t)c * lH + h];
    }
    for (int v = e0; v < V; v += stride) b_v[v] = b_v[v] + scale * gbv[v] - sparse;
    for (int h = e0; h < H; h += stride) b_h[h] = b_h[h] + scale * gbh[h] - sparse;
    for (int c = e0; c < C; c += stride) b_c[c] = b_c[c] + scale * gbc[c] - sparse;
    if (e0 == 0 && loss_out != nullptr) loss_out[0] = gloss[0] * loss_scale;
}

int check_dims(const char *who, int B, int V, int H, int C)
{
    if (B < 1 || V < 1 || H < 1 || C < 1 || C > MAXC) {
        qbm_set_error("%s: need B, V, H >= 1 and 1 <= C <= %d (B=%d V=%d H=%d C=%d)", who, B, V, H, C, MAXC);
        return QBM_EINVAL;
    }
    return QBM_OK;
}

}  // namespace

extern "C" QBM_API size_t qbm_rbm_workspace_bytes(int B, int V, int H, int C)
{
    if (B < 1 || V < 1 || H < 1 || C < 1) return 0;
    return ws_floats(B, V, H, C) * sizeof(float);
}

// test hook: float offsets of the step intermediates inside the workspace, so that tests can replay the CD-1 draws
extern "C" QBM_API int qbm_rbm_workspace_layout(int B, int V, int H, int C, long long *offsets)
{
    if (int rc = check_dims("qbm_rbm_workspace_layout", B, V, H, C)) return rc;
    QBM_CHECK_ARG(offsets, "qbm_rbm_workspace_layout: null pointer argument");
    const Ws w = carve(nullptr, B, V, H, C);
    const float *base = nullptr;
    const float *ptrs[12] = {w.A, w.P, w.Dt, w.xt, w.p0, w.p0t, w.h0, w.v1, w.v1t, w.p1t, w.pc, reinterpret_cast<float *>(w.y1)};
    for (int i = 0; i < 12; ++i) offsets[i] = (long long)(ptrs[i] - base);
    return QBM_OK;
}

// R1 (:43-47): P[B, ld4(H)] = sigmoid(v.W + b_h + U[y]);  Wt = W^T [H, ld4(V)]
extern "C" QBM_API int qbm_rbm_sample_hidden(const float *Wt, const float *U, const float *b_h, const float *v, const int *y,
                                             int B, int V, int H, int C, float *P, void *stream)
{
    if (int rc = check_dims("qbm_rbm_sample_hidden", B, V, H, C)) return rc;
    QBM_CHECK_ARG(Wt && U && b_h && v && y && P, "qbm_rbm_sample_hidden: null pointer argument");
    EpiParams ep = {};
    ep.C = P; ep.ldc = ld4(H); ep.bias_n = b_h; ep.rowtab = U; ep.ridx = y; ep.ldtab = ld4(H); ep.alpha = 1.0f; ep.act = 1;
    return qbm_gemm_tf32_launch(v, ld4(V), Wt, ld4(V), B, H, V, ep, (cudaStream_t)stream);
}

// R2 (:49-52): P[B, ld4(V)] = sigmoid(h.W^T + b_v);  W [V, ld4(H)]
extern "C" QBM_API int qbm_rbm_sample_visible(const float *W, const float *b_v, const float *hid, int B, int V, int H,
                                              float *P, void *stream)
{
    if (int rc = check_dims("qbm_rbm_sample_visible", B, V, H, 1)) return rc;
    QBM_CHECK_ARG(W && b_v && hid && P, "qbm_rbm_sample_visible: null pointer argument");
    EpiParams ep = {};
    ep.C = P; ep.ldc = ld4(V); ep.bias_n = b_v; ep.alpha = 1.0f; ep.act = 1;
    return qbm_gemm_tf32_launch(hid, ld4(H), W, ld4(H), B, V, H, ep, (cudaStream_t)stream);
}

// R3 (:54-60): P[B, ld4(C)] = normalize_L1(exp(h.U^T + b_c))
extern "C" QBM_API int qbm_rbm_sample_class(const float *U, const float *b_c, const float *hid, int B, int H, int C, float *P,
                                            void *stream)
{
    if (int rc = check_dims("qbm_rbm_sample_class", B, 1, H, C)) return rc;
    QBM_CHECK_ARG(U && b_c && hid && P, "qbm_rbm_sample_class: null pointer argument");
    rbm_class_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(hid, ld4(H), U, ld4(H), b_c, H, C, P, ld4(C), nullptr, 0ull, 0u);
    QBM_LAUNCH_OK("rbm_class_kernel");
    return QBM_OK;
}

// p(y|x) (:62-86): P[B, ld4(C)]; workspace as for the steps
extern "C" QBM_API int qbm_rbm_class_given_x(const float *Wt, const float *U, const float *b_h, const float *b_c, const float *x,
                                             int B, int V, int H, int C, float *P, void *workspace, size_t workspace_bytes,
                                             void *stream)
{
    if (int rc = check_dims("qbm_rbm_class_given_x", B, V, H, C)) return rc;
    QBM_CHECK_ARG(Wt && U && b_h && b_c && x && P && workspace, "qbm_rbm_class_given_x: null pointer argument");
    if (workspace_bytes < qbm_rbm_workspace_bytes(B, V, H, C)) { qbm_set_error("qbm_rbm_class_given_x: workspace too small"); return QBM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    Ws w = carve(workspace, B, V, H, C);
    EpiParams ep = {};
    ep.C = w.A; ep.ldc = ld4(H); ep.bias_n = b_h; ep.alpha = 1.0f;
    if (int rc = qbm_gemm_tf32_launch(x, ld4(V), Wt, ld4(V), B, H, V, ep, st)) return rc;
    rbm_rows_kernel<<<B, 256, 0, st>>>(w.A, ld4(H), U, ld4(H), b_c, nullptr, H, C, P, ld4(C));
    QBM_LAUNCH_OK("rbm_rows_kernel");
    return QBM_OK;
}

// R4 + R5 (:101-146, :88-99): one discriminative training step, parameters updated in place.
//   W [V, ld4(H)], Wt [H, ld4(V)] (kept equal to W^T), U [C, ld4(H)], b_v [V], b_h [H], b_c [C],
//   x [B, ld4(V)], y int32 [B]; outputs probs [B, ld4(C)], pred int32 [B] (nullable), loss [1] (nullable)
extern "C" QBM_API int qbm_rbm_disc_step(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *x,
                                         const int *y, int B, int V, int H, int C, float lr, float factor,
                                         float sparse_constant, float *probs, int *pred, float *loss, void *workspace,
                                         size_t workspace_bytes, void *stream)
{
    if (int rc = check_dims("qbm_rbm_disc_step", B, V, H, C)) return rc;
    QBM_CHECK_ARG(W && Wt && U && b_v && b_h && b_c && x && y && probs && workspace, "qbm_rbm_disc_step: null pointer argument");
    if (workspace_bytes < qbm_rbm_workspace_bytes(B, V, H, C)) { qbm_set_error("qbm_rbm_disc_step: workspace too small"); return QBM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    Ws w = carve(workspace, B, V, H, C);
    const long long lB = ld4(B), lV = ld4(V), lH = ld4(H), lC = ld4(C);
    const float scale = factor * lr / (float)B;
    // A = x.W + b_h
    EpiParams e1 = {};
    e1.C = w.A; e1.ldc = lH; e1.bias_n = b_h; e1.alpha = 1.0f;
    if (int rc = qbm_gemm_tf32_launch(x, lV, Wt, lV, B, H, V, e1, st)) return rc;
    // p(y|x) and D^T
    const int nht = (H + 31) / 32;
    rbm_rows_kernel<<<B, 256, 0, st>>>(w.A, lH, U, lH, b_c, y, H, C, probs, lC, x, lV, V, w.xt, lB, w.tick, nht);
    QBM_LAUNCH_OK("rbm_rows_kernel");
    // D^T, class weights / hidden bias (reads the pre-update A, U); the extra row of blocks: class bias, loss, argmax
    rbm_disc_grad_kernel<<<dim3(nht, DSL + 1), 256, 0, st>>>(w.A, lH, U, lH, b_h, probs, lC, y, w.Dt, lB, B, H, C, scale,
                                                             sparse_constant, nullptr, nullptr, w.part, w.tick, b_c, b_v, V,
                                                             pred, loss, 1, nullptr);
    QBM_LAUNCH_OK("rbm_disc_grad_kernel");
    // W += scale * x^T.D   (SGD update fused into the GEMM epilogue, which also writes the K-major copy W^T)
    EpiParams e2 = {};
    e2.C = W; e2.ldc = lH; e2.Cin = W; e2.ldcin = lH; e2.alpha = scale; e2.beta = 1.0f; e2.Ct = Wt; e2.ldct = lV;
    return qbm_gemm_tf32_launch(w.xt, lB, w.Dt, lB, V, H, B, e2, st);
}

// CD-1 step composed from the primitives:
//   h0 ~ Bern(R1(v0,y0)); v1 ~ Bern(R2(h0)); y1 ~ Cat(R3(h0)); ph1 = R1(v1,y1);
//   dW = v0^T ph0 - v1^T ph1; dU = y0^T ph0 - y1^T ph1; db_v = sum(v0-v1); db_h = sum(ph0-ph1); db_c = sum(y0-y1)
static int cd1_step_impl(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *v0,
                         const int *y0, int B, int V, int H, int C, float lr, float sparse_constant,
                         unsigned long long seed, unsigned int step, const unsigned int *step_dev, void *workspace,
                         size_t workspace_bytes, void *stream)
{
    if (int rc = check_dims("qbm_rbm_cd1_step", B, V, H, C)) return rc;
    QBM_CHECK_ARG(W && Wt && U && b_v && b_h && b_c && v0 && y0 && workspace, "qbm_rbm_cd1_step: null pointer argument");
    if (workspace_bytes < qbm_rbm_workspace_bytes(B, V, H, C)) { qbm_set_error("qbm_rbm_cd1_step: workspace too small"); return QBM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    Ws w = carve(workspace, B, V, H, C);
    const long long lB = ld4(B), lV = ld4(V), lH = ld4(H), lC = ld4(C);
    const float scale = lr / (float)B;
    // positive phase: ph0 (+ transposed), h0 ~ Bernoulli(ph0)
    EpiParams e = {};
    e.C = w.p0; e.ldc = lH; e.Ct = w.p0t; e.ldct = lB; e.S = w.h0; e.lds = lH; e.bias_n = b_h; e.rowtab = U; e.ridx = y0;
    e.ldtab = lH; e.alpha = 1.0f; e.act = 1; e.seed = seed; e.stream = step * 4u + 0u; e.step_dev = step_dev;
    if (int rc = qbm_gemm_tf32_launch(v0, lV, Wt, lV, B, H, V, e, st)) return rc;
    // negative phase: v1 ~ Bernoulli(sigmoid(h0.W^T + b_v)) (+ transposed), y1 ~ Cat(p(y|h0))
    EpiParams e2 = {};
    e2.S = w.v1; e2.lds = lV; e2.St = w.v1t; e2.ldst = lB; e2.bias_n = b_v; e2.alpha = 1.0f; e2.act = 1; e2.seed = seed;
    e2.stream = step * 4u + 1u; e2.step_dev = step_dev;
    if (int rc = qbm_gemm_tf32_launch(w.h0, lH, W, lH, B, V, H, e2, st)) return rc;
    rbm_class_kernel<<<B, 128, 0, st>>>(w.h0, lH, U, lH, b_c, H, C, w.pc, lC, w.y1, seed, step * 4u + 2u, step_dev, v0, lV, V, w.xt, lB);
    QBM_LAUNCH_OK("rbm_class_kernel");
    EpiParams e3 = {};
    e3.Ct = w.p1t; e3.ldct = lB; e3.bias_n = b_h; e3.rowtab = U; e3.ridx = w.y1; e3.ldtab = lH; e3.alpha = 1.0f; e3.act = 1;
    if (int rc = qbm_gemm_tf32_launch(w.v1, lV, Wt, lV, B, H, V, e3, st)) return rc;
    // small parameters, then W += scale (v0^T ph0 - v1^T ph1) fused into two GEMM epilogues, then W^T
    const int mx = (V > H ? V : H) > C ? (V > H ? V : H) : C;
    rbm_cd_small_update_kernel<<<(mx + 7) / 8, 256, 0, st>>>(w.xt, w.v1t, lB, w.p0t, w.p1t, lB, y0, w.y1, U, lH, b_v, b_h,
                                                                b_c, B, V, H, C, scale, sparse_constant, 0);
    QBM_LAUNCH_OK("rbm_cd_small_update_kernel");
    EpiParams g1 = {};
    g1.C = W; g1.ldc = lH; g1.Cin = W; g1.ldcin = lH; g1.alpha = scale; g1.beta = 1.0f;
    if (int rc = qbm_gemm_tf32_launch(w.xt, lB, w.p0t, lB, V, H, B, g1, st)) return rc;
    g1.alpha = -scale; g1.Ct = Wt; g1.ldct = lV;              // the second accumulate also writes the K-major copy W^T
    return qbm_gemm_tf32_launch(w.v1t, lB, w.p1t, lB, V, H, B, g1, st);
}

extern "C" QBM_API int qbm_rbm_cd1_step(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *v0,
                                        const int *y0, int B, int V, int H, int C, float lr, float sparse_constant,
                                        unsigned long long seed, unsigned int step, void *workspace, size_t workspace_bytes,
                                        void *stream)
{
    return cd1_step_impl(W, Wt, U, b_v, b_h, b_c, v0, y0, B, V, H, C, lr, sparse_constant, seed, step, nullptr, workspace,
                         workspace_bytes, stream);
}

// the same step with the step counter read from device memory: draws use step + *step_dev, so a captured CUDA graph of
// this call advances through the Philox streams when the caller increments the counter between replays
extern "C" QBM_API int qbm_rbm_cd1_step_dev(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *v0,
                                            const int *y0, int B, int V, int H, int C, float lr, float sparse_constant,
                                            unsigned long long seed, unsigned int step, const unsigned int *step_dev,
                                            void *workspace, size_t workspace_bytes, void *stream)
{
    QBM_CHECK_ARG(step_dev, "qbm_rbm_cd1_step_dev: null step counter");
    return cd1_step_impl(W, Wt, U, b_v, b_h, b_c, v0, y0, B, V, H, C, lr, sparse_constant, seed, step, step_dev, workspace,
                         workspace_bytes, stream);
}

// ---- gradient variants of the two steps + the fused apply: what the data-parallel trainers run ------------------------
// (shard gradient -> ONE all-reduce of the flat buffer -> identical apply on every rank)
extern "C" QBM_API size_t qbm_rbm_grad_count(int V, int H, int C)
{
    if (V < 1 || H < 1 || C < 1) return 0;
    return grad_floats(V, H, C);
}

// R4 (:101-146) without R5: the raw gradient sums of this shard (dW = x^T D, dU, db_h, db_c, db_v = 0) and the loss SUM
extern "C" QBM_API int qbm_rbm_disc_grad(const float *Wt, const float *U, const float *b_h, const float *b_c, const float *x,
                                         const int *y, int B, int V, int H, int C, float *grad, float *probs, int *pred,
                                         void *workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = check_dims("qbm_rbm_disc_grad", B, V, H, C)) return rc;
    QBM_CHECK_ARG(Wt && U && b_h && b_c && x && y && grad && probs && workspace, "qbm_rbm_disc_grad: null pointer argument");
    QBM_CHECK_ARG((reinterpret_cast<uintptr_t>(grad) & 15u) == 0, "qbm_rbm_disc_grad: grad must be 16-byte aligned");
    if (workspace_bytes < qbm_rbm_workspace_bytes(B, V, H, C)) { qbm_set_error("qbm_rbm_disc_grad: workspace too small"); return QBM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    Ws w = carve(workspace, B, V, H, C);
    const Grad g = grad_carve(grad, V, H, C);
    const long long lB = ld4(B), lV = ld4(V), lH = ld4(H), lC = ld4(C);
    EpiParams e1 = {};
    e1.C = w.A; e1.ldc = lH; e1.bias_n = b_h; e1.alpha = 1.0f;
    if (int rc = qbm_gemm_tf32_launch(x, lV, Wt, lV, B, H, V, e1, st)) return rc;
    const int nht = (H + 31) / 32;
    rbm_rows_kernel<<<B, 256, 0, st>>>(w.A, lH, U, lH, b_c, y, H, C, probs, lC, x, lV, V, w.xt, lB, w.tick, nht);
    QBM_LAUNCH_OK("rbm_rows_kernel");
    rbm_disc_grad_kernel<<<dim3(nht, DSL + 1), 256, 0, st>>>(w.A, lH, const_cast<float *>(U), lH, nullptr, probs, lC, y, w.Dt, lB,
                                                             B, H, C, 0.0f, 0.0f, g.gU, g.gbh, w.part, w.tick, nullptr, nullptr,
                                                             V, pred, g.loss, 2, g.gbc);
    QBM_LAUNCH_OK("rbm_disc_grad_kernel");
    QBM_CUDA_OK(cudaMemsetAsync(g.gbv, 0, (size_t)V * sizeof(float), st));        // the discriminative gradient has no b_v term (:138)
    EpiParams e2 = {};
    e2.C = g.gW; e2.ldc = lH; e2.alpha = 1.0f;
    return qbm_gemm_tf32_launch(w.xt, lB, w.Dt, lB, V, H, B, e2, st);
}

// CD-1 (the composition of :43-60) without R5: dW = v0^T ph0 - v1^T ph1, dU, db_v, db_h, db_c of this shard
static int cd1_grad_impl(const float *W, const float *Wt, const float *U, const float *b_v, const float *b_h,
                         const float *b_c, const float *v0, const int *y0, int B, int V, int H, int C,
                         unsigned long long seed, unsigned int step, const unsigned int *step_dev, float *grad,
                         void *workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = check_dims("qbm_rbm_cd1_grad", B, V, H, C)) return rc;
    QBM_CHECK_ARG(W && Wt && U && b_v && b_h && b_c && v0 && y0 && grad && workspace, "qbm_rbm_cd1_grad: null pointer argument");
    QBM_CHECK_ARG((reinterpret_cast<uintptr_t>(grad) & 15u) == 0, "qbm_rbm_cd1_grad: grad must be 16-byte aligned");
    if (workspace_bytes < qbm_rbm_workspace_bytes(B, V, H, C)) { qbm_set_error("qbm_rbm_cd1_grad: workspace too small"); return QBM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    Ws w = carve(workspace, B, V, H, C);
    const Grad g = grad_carve(grad, V, H, C);
    const long long lB = ld4(B), lV = ld4(V), lH = ld4(H), lC = ld4(C);
    EpiParams e = {};
    e.C = w.p0; e.ldc = lH; e.Ct = w.p0t; e.ldct = lB; e.S = w.h0; e.lds = lH; e.bias_n = b_h; e.rowtab = U; e.ridx = y0;
    e.ldtab = lH; e.alpha = 1.0f; e.act = 1; e.seed = seed; e.stream = step * 4u + 0u; e.step_dev = step_dev;
    if (int rc = qbm_gemm_tf32_launch(v0, lV, Wt, lV, B, H, V, e, st)) return rc;
    EpiParams e2 = {};
    e2.S = w.v1; e2.lds = lV; e2.St = w.v1t; e2.ldst = lB; e2.bias_n = b_v; e2.alpha = 1.0f; e2.act = 1; e2.seed = seed;
    e2.stream = step * 4u + 1u; e2.step_dev = step_dev;
    if (int rc = qbm_gemm_tf32_launch(w.h0, lH, W, lH, B, V, H, e2, st)) return rc;
    rbm_class_kernel<<<B, 128, 0, st>>>(w.h0, lH, U, lH, b_c, H, C, w.pc, lC, w.y1, seed, step * 4u + 2u, step_dev, v0, lV, V, w.xt, lB);
    QBM_LAUNCH_OK("rbm_class_kernel");
    EpiParams e3 = {};
    e3.Ct = w.p1t; e3.ldct = lB; e3.bias_n = b_h; e3.rowtab = U; e3.ridx = w.y1; e3.ldtab = lH; e3.alpha = 1.0f; e3.act = 1;
    if (int rc = qbm_gemm_tf32_launch(w.v1, lV, Wt, lV, B, H, V, e3, st)) return rc;
    const int mx = (V > H ? V : H) > C ? (V > H ? V : H) : C;
    rbm_cd_small_update_kernel<<<(mx + 7) / 8, 256, 0, st>>>(w.xt, w.v1t, lB, w.p0t, w.p1t, lB, y0, w.y1, g.gU, lH, g.gbv, g.gbh,
                                                                g.gbc, B, V, H, C, 0.0f, 0.0f, 1);
    QBM_LAUNCH_OK("rbm_cd_small_update_kernel");
    QBM_CUDA_OK(cudaMemsetAsync(g.loss, 0, 4 * sizeof(float), st));
    EpiParams g1 = {};
    g1.C = g.gW; g1.ldc = lH; g1.alpha = 1.0f;
    if (int rc = qbm_gemm_tf32_launch(w.xt, lB, w.p0t, lB, V, H, B, g1, st)) return rc;
    g1.Cin = g.gW; g1.ldcin = lH; g1.alpha = -1.0f; g1.beta = 1.0f;
    return qbm_gemm_tf32_launch(w.v1t, lB, w.p1t, lB, V, H, B, g1, st);
}

extern "C" QBM_API int qbm_rbm_cd1_grad(const float *W, const float *Wt, const float *U, const float *b_v, const float *b_h,
                                        const float *b_c, const float *v0, const int *y0, int B, int V, int H, int C,
                                        unsigned long long seed, unsigned int step, float *grad, void *workspace,
                                        size_t workspace_bytes, void *stream)
{
    return cd1_grad_impl(W, Wt, U, b_v, b_h, b_c, v0, y0, B, V, H, C, seed, step, nullptr, grad, workspace, workspace_bytes, stream);
}

// step counter in device memory (draws keyed by step + *step_dev), for CUDA-graph replays of the data-parallel step
extern "C" QBM_API int qbm_rbm_cd1_grad_dev(const float *W, const float *Wt, const float *U, const float *b_v, const float *b_h,
                                            const float *b_c, const float *v0, const int *y0, int B, int V, int H, int C,
                                            unsigned long long seed, unsigned int step, const unsigned int *step_dev,
                                            float *grad, void *workspace, size_t workspace_bytes, void *stream)
{
    QBM_CHECK_ARG(step_dev, "qbm_rbm_cd1_grad_dev: null step counter");
    return cd1_grad_impl(W, Wt, U, b_v, b_h, b_c, v0, y0, B, V, H, C, seed, step, step_dev, grad, workspace, workspace_bytes, stream);
}

// R5 (:88-99) on the (all-reduced) gradient buffer: param += scale * grad with scale = factor * lr / global batch, the three
// biases -= sparse_constant, W^T refreshed in the same pass; loss_out (nullable) = loss sum * loss_scale
extern "C" QBM_API int qbm_rbm_apply_grad(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *grad,
                                          int V, int H, int C, float scale, float sparse_constant, float *loss_out,
                                          float loss_scale, void *stream)
{
    if (int rc = check_dims("qbm_rbm_apply_grad", 1, V, H, C)) return rc;
    QBM_CHECK_ARG(W && Wt && U && b_v && b_h && b_c && grad, "qbm_rbm_apply_grad: null pointer argument");
    const Grad g = grad_carve(const_cast<float *>(grad), V, H, C);
    const int tx = (H + 31) / 32, ty = (V + 31) / 32;
    const int tiles = tx * ty;
    rbm_apply_kernel<<<tiles + 8, 256, 0, (cudaStream_t)stream>>>(W, Wt, U, b_v, b_h, b_c, g.gW, g.gU, g.gbv, g.gbh, g.gbc, g.loss, V,
                                                                   H, C, ld4(H), ld4(V), scale, sparse_constant, loss_out,
                                                                   loss_scale, tx, tiles);
    QBM_LAUNCH_OK("rbm_apply_kernel");
    return QBM_OK;
}
