/*
 * qbm_b200.h -- C ABI of libqbm_b200.so, the B200 (sm_100a) implementation of the
 * sampling-and-training hot path of Mark-Seebode/QBM-Image-Classification.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless marked "host";
 *     the library allocates nothing and performs no host<->device copies or synchronisation
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and the call returns immediately
 *   - return value: 0 on success, a negative QBM_E* code otherwise; the message is available
 *     from qbm_last_error() (thread-local); nothing is thrown across the ABI
 *   - citations "ref:" are into /root/reference (the interface each entry point replaces)
 */
#ifndef QBM_B200_H
#define QBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define QBM_API __attribute__((visibility("default")))
#else
#define QBM_API
#endif

#define QBM_OK            0
#define QBM_EINVAL       -1   /* bad argument (shape, alignment, null pointer) */
#define QBM_EUNSUPPORTED -2   /* size outside what the kernels are built for */
#define QBM_ECUDA        -3   /* CUDA runtime error (message holds cudaGetErrorString) */
#define QBM_EWORKSPACE   -4   /* caller workspace too small (see *_workspace_bytes) */

#define QBM_SA_MAX_N   2048   /* largest QUBO the SA kernel is instantiated for */
#define QBM_TWO_PHASE_MIN_N 896  /* the sampler's two-phase schedule is the default for n above this */

int         qbm_version(void);              /* ABI version, currently 1 */
const char *qbm_last_error(void);           /* host string, thread-local */
/* sm_count / cc_major / cc_minor of the current device (host out-pointers, nullable). */
int         qbm_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------
 * K0  dense QUBO -> spin model.
 * ref: dimod.BQM(Q, "BINARY") + change_vartype(SPIN) as called from src/qubo/sampler.py:7-8,31
 *      and src/model/faster_dqbm.py:577,619.
 *   Q        [batch, n, n] float64 row-major (any triangle layout; b_ij = Q_ij + Q_ji)
 *   J_out    [batch, n, n] float32 symmetric couplings J_ij = b_ij/4, zero diagonal
 *   h_out    [batch, n]    float32 h_i = Q_ii/2 + sum_j b_ij/4   (accumulated in float64)
 *   offset   [batch]       float64 sum_i Q_ii/2 + sum_{i<j} b_ij/4            (nullable)
 *   range    [batch, 2]    float64 {min non-zero |bias|, max_i (|h_i| + sum_j |J_ij|)}: the two
 *                          numbers neal's legacy _default_ising_beta_range reduces to (A.4);
 *                          {0, 0} when every bias is zero                      (nullable)
 */
int qbm_qubo_to_ising(const double *Q, int n, long long batch, float *J_out, float *h_out,
                      double *offset, double *range, void *stream);
/* Geometric beta schedule from K0's `range` output: hot = ln 2 / range[1], cold = ln 100 / range[0] (neal's legacy
 * _default_ising_beta_range; [0.1, 1.0] when every bias is zero), then np.geomspace(hot, cold, num_betas) cast to
 * float32.  ref: neal/sampler.py as called from src/qubo/sampler.py:31-32.
 *   range [batch, 2] float64, betas_out [batch, num_betas] float32 */
int qbm_beta_schedule(const double *range, long long batch, int num_betas, float *betas_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * K1  simulated-annealing sampler (one warp per read/chain).
 * ref: neal.SimulatedAnnealingSampler.sample -> cpu_sa.cpp, reached through
 *      src/qubo/sampler.py:31-33 (LocalSASampler.sample_Q) and
 *      src/model/faster_dqbm.py:299-313 (Disc_QBM.sample_sa) -- SURVEY.md Appendix A.5.
 *   J, h      spin model of `batch_q` problems: J [batch_q, n, ldj] float32 (symmetric, zero
 *             diagonal, ldj >= n), h [batch_q, n] float32
 *   beta      [batch_q or 1, num_betas] float32 inverse temperatures; beta_stride = elements
 *             between consecutive problems' schedules (0 = one schedule shared by all)
 *   sweeps_per_beta  sweeps at each beta (neal: max(1, num_sweeps // 1000))
 *   num_reads chains per problem; chain g = chain_offset + q*num_reads + r keys the Philox
 *             stream, so results do not depend on how reads are sharded over launches/GPUs
 *   init_states  nullable [batch_q, num_reads, n] int8 0/1; NULL = Philox initial states
 *   states_out   [batch_q, num_reads, n] int8 0/1, read order (what dimod calls record.sample)
 *   counters     nullable uint64[2]: += accepted flips, += proposals
 *   workspace    scratch of at least qbm_sa_workspace_bytes(n, batch_q) bytes, 16-byte aligned.  With at least
 *                qbm_sa_workspace_bytes_two_phase(n, batch_q, num_reads) bytes the sampler may run the two-phase
 *                schedule (by default at n > 896: the chain-tile kernel anneals the hot sweeps, then hands every chain -- fields,
 *                spins, sweep counter -- to the warp-per-chain kernel; identical results, 1.36x (n = 1280) .. 1.8x (n = 2048) faster)
 *   flags        bit 0: make the warps of a CTA rendezvous at every 128-variable window (A-B measurements; off by
 *                       default because it measured slower)
 *                bit 1: chain g = chain_offset + r for every problem (all problems share one random
 *                       stream, as the reference's fixed per-call seed does)
 *                bit 5: use the multi-chain warp kernel (a warp anneals 2-4 chains of one problem and shares their
 *                       coupling-row loads; identical trajectories, n > 128 and num_reads >= 2 only)
 *                bit 6: never use the two-phase schedule; bit 7: use it wherever it is supported (n > 256, not 7 windows) instead of
 *                       only where it is the measured default (n > QBM_TWO_PHASE_MIN_N); bits 16..23: its hand-over
 *                       threshold in percent of accepted proposals per sweep (0 = default 50)
 *                bit 4: use the chain-tile kernel (16 chains per CTA share every coupling row, rows streamed
 *                       by TMA through a shared-memory ring; identical trajectories);
 *                       bits 8..15: its dense/sparse update switch in percent of flipped (chain, variable)
 *                       pairs per 32-variable sub-window (0 = default 40)
 */
size_t qbm_sa_workspace_bytes(int n, long long batch_q);
size_t qbm_sa_workspace_bytes_two_phase(int n, long long batch_q, long long num_reads);
int qbm_sa_sample(const float *J, const float *h, int n, int ldj, long long batch_q,
                  const float *beta, long long beta_stride, int num_betas, int sweeps_per_beta,
                  long long num_reads, uint64_t seed, uint64_t chain_offset,
                  const int8_t *init_states, int8_t *states_out, unsigned long long *counters,
                  void *workspace, size_t workspace_bytes, unsigned flags, void *stream);

/* ------------------------------------------------------------------------------------------
 * K2  batched QUBO energies  E[q, r] = x^T Q_q x  in float64.
 * ref: neal get_state_energy + dimod offset = the BINARY energy in SampleSet.record.energy
 *      (SURVEY.md A.2/A.6).
 *   Q [batch_q, n, n] float64, states [batch_q, R, n] int8 0/1, energy_out [batch_q, R] float64
 */
int qbm_qubo_energy(const double *Q, int n, long long batch_q, const int8_t *states, long long R,
                    double *energy_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * K3  phase statistics: first and second moments of a sample set.
 * ref: the np.average / (block^T @ block) / n_reads reductions in
 *      src/train/train.py:135-253 (get_average_configuration_single) and
 *      src/model/discriminative_qbm.py:696-760 (get_average_configuration).
 *   states [batch_q, R, n] int8 0/1
 *   mean_out   [batch_q, n]    float64  <s_i>
 *   second_out [batch_q, n, n] float64  <s_i s_j> (full symmetric matrix), nullable
 *   workspace  scratch of at least qbm_phase_stats_workspace_bytes(batch_q, R, n) bytes
 * Counts are accumulated as exact integers (bit-plane popcounts) and divided by R once in float64,
 * which is bit-identical to numpy's float64 mean of 0/1 products.
 */
size_t qbm_phase_stats_workspace_bytes(long long batch_q, long long R, int n);
int qbm_phase_stats(const int8_t *states, long long batch_q, long long R, int n, double *mean_out,
                    double *second_out, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * K4  TF32 tensor-core GEMM (tcgen05 + TMEM + TMA):  C[M,N] = act(alpha * A[M,K] . B[N,K]^T + bias_n) + beta * Cin
 * ref: torch.matmul / expand-mul-sum in src/ClassificationRBM.py:44-56,106,118-128.
 *   A [M, lda], B [N, ldb] row-major float32 with K contiguous; lda, ldb multiples of 4, 16-byte aligned
 *   act: 0 identity, 1 sigmoid;  bias_n [N], Cin [M, ldcin] nullable;  C [M, ldc] and/or Ct [N, ldct] (= C^T)
 */
int qbm_gemm_tf32(const float *A, long long lda, const float *B, long long ldb, int M, int N, int K,
                  float alpha, float beta, const float *Cin, long long ldcin, const float *bias_n, int act,
                  float *C, long long ldc, float *Ct, long long ldct, void *stream);

/* ------------------------------------------------------------------------------------------
 * K4/K5  ClassificationRBM (ref: src/ClassificationRBM.py).  Storage contract: every matrix is
 * row-major float32 with leading dimension ld4(cols) = (cols + 3) & ~3:
 *   W [V, ld4(H)]  (:26 weights), Wt [H, ld4(V)] = W^T (kept in sync by the step functions),
 *   U [C, ld4(H)]  (:30 class_weights), b_v [V], b_h [H], b_c [C], x / v [B, ld4(V)], hid [B, ld4(H)],
 *   y int32 [B], probabilities over classes [B, ld4(C)].  C <= 32.
 * qbm_rbm_sample_hidden   :43-47   P[B, ld4(H)] = sigmoid(v W + b_h + U[y])
 * qbm_rbm_sample_visible  :49-52   P[B, ld4(V)] = sigmoid(hid W^T + b_v)
 * qbm_rbm_sample_class    :54-60   P[B, ld4(C)] = L1-normalised exp(hid U^T + b_c)
 * qbm_rbm_class_given_x   :62-86   P[B, ld4(C)] = p(y|x)
 * qbm_rbm_disc_step       :101-146 + :88-99  exact discriminative gradient step, in place;
 *                         probs [B, ld4(C)], pred int32 [B] (nullable), loss float[1] (nullable; the
 *                         reference's CrossEntropyLoss applied to the probabilities, :142)
 * qbm_rbm_cd1_step        CD-1 composed from the three primitives (the reference stores k and the
 *                         primitives but never wires them): h0 ~ Bern(ph0), v1 ~ Bern,
 *                         y1 ~ Cat, dW = v0^T ph0 - v1^T ph1, ...; Philox keyed by (seed, step)
 * workspace: qbm_rbm_workspace_bytes(B, V, H, C) bytes, 16-byte aligned.
 */
size_t qbm_rbm_workspace_bytes(int B, int V, int H, int C);
int qbm_rbm_sample_hidden(const float *Wt, const float *U, const float *b_h, const float *v, const int *y,
                          int B, int V, int H, int C, float *P, void *stream);
int qbm_rbm_sample_visible(const float *W, const float *b_v, const float *hid, int B, int V, int H, float *P,
                           void *stream);
int qbm_rbm_sample_class(const float *U, const float *b_c, const float *hid, int B, int H, int C, float *P,
                         void *stream);
int qbm_rbm_class_given_x(const float *Wt, const float *U, const float *b_h, const float *b_c, const float *x,
                          int B, int V, int H, int C, float *P, void *workspace, size_t workspace_bytes,
                          void *stream);
int qbm_rbm_disc_step(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *x,
                      const int *y, int B, int V, int H, int C, float lr, float factor, float sparse_constant,
                      float *probs, int *pred, float *loss, void *workspace, size_t workspace_bytes, void *stream);
int qbm_rbm_cd1_step(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *v0,
                     const int *y0, int B, int V, int H, int C, float lr, float sparse_constant,
                     unsigned long long seed, unsigned int step, void *workspace, size_t workspace_bytes,
                     void *stream);
/* The same step with the step counter in device memory (draws are keyed by step + *step_dev): a CUDA graph captured around
 * this call advances through the Philox streams when the caller increments the counter between replays. */
int qbm_rbm_cd1_step_dev(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *v0,
                         const int *y0, int B, int V, int H, int C, float lr, float sparse_constant,
                         unsigned long long seed, unsigned int step, const unsigned int *step_dev, void *workspace,
                         size_t workspace_bytes, void *stream);

/* Data-parallel form of the two steps (minibatch rows sharded over the GPUs, ONE all-reduce of the
 * parameter-shaped statistics, identical update on every rank).  The gradient variants leave the parameters untouched and
 * write the raw sums of their shard into one flat buffer of qbm_rbm_grad_count(V, H, C) floats, 16-byte aligned:
 *   [ dW (V x ld4(H)) | dU (C x ld4(H)) | db_v (V) | db_h (H) | db_c (C) | loss sum (1) ]   (segments padded to 4 floats)
 * which the caller all-reduces (sum) as it is; qbm_rbm_apply_grad then performs update_weights (ref: :88-99):
 * param += scale * grad with scale = factor * lr / global batch, the three biases -= sparse_constant, W^T refreshed in
 * the same pass, loss_out (nullable) = loss sum * loss_scale.
 */
size_t qbm_rbm_grad_count(int V, int H, int C);
int qbm_rbm_disc_grad(const float *Wt, const float *U, const float *b_h, const float *b_c, const float *x, const int *y,
                      int B, int V, int H, int C, float *grad, float *probs, int *pred, void *workspace,
                      size_t workspace_bytes, void *stream);
int qbm_rbm_cd1_grad(const float *W, const float *Wt, const float *U, const float *b_v, const float *b_h, const float *b_c,
                     const float *v0, const int *y0, int B, int V, int H, int C, unsigned long long seed, unsigned int step,
                     float *grad, void *workspace, size_t workspace_bytes, void *stream);
/* qbm_rbm_cd1_grad with the step counter in device memory (see qbm_rbm_cd1_step_dev). */
int qbm_rbm_cd1_grad_dev(const float *W, const float *Wt, const float *U, const float *b_v, const float *b_h, const float *b_c,
                     const float *v0, const int *y0, int B, int V, int H, int C, unsigned long long seed, unsigned int step, const unsigned int *step_dev,
                     float *grad, void *workspace, size_t workspace_bytes, void *stream);

/* Peer-memory form of the all-reduce + apply (ranks of one NVLink node, one process per GPU).  Every rank allocates
 * qbm_rbm_peer_bytes(V, H, C) with qbm_peer_alloc (plain cudaMalloc, zeroed: two gradient halves | arrival flags | error
 * word), exports it (64-byte CUDA IPC handle), exchanges the handles by its own means and maps the peers' allocations with
 * qbm_peer_import.  Per step: the gradient kernels write half `parity` (base + parity * qbm_rbm_grad_count floats), then
 * qbm_rbm_apply_grad_peer stores `token` into this rank's slot of every rank's flag array, waits until all slots of its own
 * array have reached `token` (bounded spin: a rank that never arrives sets the error word, see qbm_rbm_peer_error, instead
 * of hanging the GPU), sums the gradients of all ranks in rank order straight from peer memory and applies them like
 * qbm_rbm_apply_grad.  parity must alternate and token (+ *tick_dev when given: CUDA-graph replays) must grow by one
 * every step, on all ranks alike. */
size_t qbm_rbm_peer_bytes(int V, int H, int C);
int qbm_peer_alloc(size_t bytes, void **ptr);
int qbm_peer_free(void *ptr);
int qbm_peer_export(void *ptr, unsigned char *handle64);
int qbm_peer_import(const unsigned char *handle64, void **ptr);
int qbm_peer_close(void *ptr);
int qbm_rbm_peer_error(const void *own_base, int V, int H, int C, unsigned int *flag_out);
int qbm_rbm_apply_grad_peer(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, void *const *peer_bases,
                            int world, int rank, int parity, int V, int H, int C, float scale, float sparse_constant,
                            float *loss_out, float loss_scale, unsigned int token, const unsigned int *tick_dev, void *stream);
int qbm_rbm_apply_grad(float *W, float *Wt, float *U, float *b_v, float *b_h, float *b_c, const float *grad, int V, int H,
                       int C, float scale, float sparse_constant, float *loss_out, float loss_scale, void *stream);

/* ------------------------------------------------------------------------------------------
 * K7 / K8 / K9  the Disc_QBM training step around the sampler, one launch each per minibatch (float64).
 * Parameters live in ONE flat buffer, which is also the layout of the error buffer:
 *   [ b_h (h) | b_o (no) | W_vh ((no+di) x h) | W_vo (di x no) | W_oo (no x no) | W_hh (h x h; absent when restricted) ]
 *   qbm_disc_param_count  number of elements of that buffer
 *   qbm_disc_build_qubo   ref: create_qubo_matrix_from, src/model/faster_dqbm.py:225-284.  X [B, di]; Y [B, no] label
 *                         rows (clamped phase, variables = hidden units) or NULL (unclamped, variables = outputs then
 *                         hidden); Q_out [B, n, n] upper-triangular, divided by beta_eff
 *   qbm_disc_errors       ref: get_average_configuration (src/model/discriminative_qbm.py:696-760; faster_mode = 0) or
 *                         get_average_configuration_batch incl. its quirks (src/model/faster_dqbm.py:754-848; faster_mode
 *                         = 1), summed over the B local images, clamped - unclamped, from the moments of K3:
 *                         mean_c [B, h], second_c [B, h, h] (nullable when unused), mean_u [B, no+h], second_u
 *                         [B, no+h, no+h].  err_out has qbm_disc_param_count + 1 elements; the last one is the NLL sum
 *                         of faster_dqbm.py:972-994 (0 for faster_mode = 0, where the reference has it commented out)
 *   qbm_sgd_apply         ref: faster_dqbm.py:1042-1059.  params[i] -= lr * (err[i] / batch)
 */
long long qbm_disc_param_count(int dim_input, int n_output, int n_hidden, int restricted);
int qbm_disc_build_qubo(const double *params, int dim_input, int n_output, int n_hidden, int restricted, const double *X,
                        const double *Y, long long B, double beta_eff, double *Q_out, void *stream);
int qbm_disc_errors(int dim_input, int n_output, int n_hidden, int restricted, int faster_mode, const double *X,
                    const double *Y, long long B, const double *mean_c, const double *second_c, const double *mean_u,
                    const double *second_u, double *err_out, void *stream);
int qbm_sgd_apply(double *params, const double *err, long long count, double lr, double batch, void *stream);

/* ------------------------------------------------------------------------------------------
 * K6  Conv-Deep inference context for a minibatch: valid convolution with the shared kernel,
 * deterministic p x p pooling (argmin per window) and the input patch of every active unit.
 * ref: src/model/geometry.py:37-53 (conv2d_valid_stride), src/model/layers.py:65-84
 *      (pooled_indices_for_input), src/train/train.py:188-191 (patch gather), reached through
 *      src/model/inference.py:16-44 (prepare_context).
 *   X [B, ih, iw] float64, kernel [k, k] float64 (k*k <= 128)
 *   fmap_out    [B, oh*ow] float64, bit-identical to the reference (numpy's pairwise order)
 *   pooled_out  [B, P] int32 index into the flattened feature map, P = qbm_convdeep_num_pooled(...)
 *               (pool in {0,1}: every conv unit is active, P = oh*ow)
 *   patches_out [B, P, k, k] float64 (nullable)
 */
int qbm_convdeep_num_pooled(int ih, int iw, int k, int stride, int pool);
int qbm_convdeep_context(const double *X, const double *kernel, long long B, int ih, int iw, int k, int stride,
                         int pool, double *fmap_out, int *pooled_out, double *patches_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * K10 / K11  the Conv-Deep training step around the sampler, one launch each per minibatch (float64).
 * Model structure: P pooled conv units, num_layers sequential layers of layer_sizes[] units (host array, at most 8),
 * n_labels outputs; QUBO variables = [pooled | sequential layers | outputs (unclamped only)].
 * Parameters live in ONE flat buffer, which is also the layout of the error buffer:
 *   [ b_conv (1 when shared_bias, else absent) | b_seq | b_out | kernel (k*k) | W_seq[0..L-1] | W_intra[0..L-1]
 *     (absent when restricted) | W_hy (last layer x n_labels) | W_oo (n_labels x n_labels) ]
 *   qbm_convdeep_build_qubo  ref: build_unclamped_qubo / build_clamped_qubo, src/qubo/builder.py:21-110.
 *                            fmap [B, num_conv], pooled [B, P] from qbm_convdeep_context; Y [B, n_labels] label
 *                            vectors (clamped) or NULL (unclamped); Q_out [B, n, n] divided by beta_eff
 *   qbm_convdeep_errors      ref: get_average_configuration_single (src/train/train.py:135-253), clamped - unclamped,
 *                            summed over the B local images, from the moments of K3 (mean_c [B, nh], second_c
 *                            [B, nh, nh], mean_u [B, nh+nl], second_u [B, nh+nl, nh+nl]); round_float32: round the
 *                            moments through float32 like the reference's float32 sample matrix; labels int32 [B];
 *                            err_out has param_count + 1 elements, the last is the loss sum of train.py:45-50
 */
long long qbm_convdeep_param_count(int P, int num_layers, const int *layer_sizes, int n_labels, int kernel_size,
                                   int restricted, int shared_bias);
int qbm_convdeep_build_qubo(const double *params, int P, int num_layers, const int *layer_sizes, int n_labels,
                            int kernel_size, int restricted, int shared_bias, const double *fmap, int num_conv,
                            const int *pooled, const double *Y, long long B, double beta_eff, double *Q_out, void *stream);
int qbm_convdeep_errors(int P, int num_layers, const int *layer_sizes, int n_labels, int kernel_size, int restricted,
                        int shared_bias, int round_float32, int one_hot, const double *patches, const double *Y,
                        const int *labels, long long B, const double *mean_c, const double *second_c, const double *mean_u,
                        const double *second_u, double *err_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Test hooks: run the device versions of the trajectory primitives on `count` inputs so that
 * tests can compare them bit-for-bit with the oracle's independent C restatement.
 *   qbm_test_philox: ctr [count,4] u32, key [count,2] u32 -> out [count,4] u32
 *   qbm_test_neg_log: u [count] u32 -> out [count] f32 (neg_log_u32 of DESIGN.md section 3)
 */
int qbm_test_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out, long long count, void *stream);
int qbm_test_neg_log(const uint32_t *u, float *out, long long count, void *stream);
/*   qbm_rbm_workspace_layout: offsets (in floats, host long long[12]) of the step intermediates inside the RBM workspace,
 *   in this order: A [B,ld4(H)], P [B,ld4(C)], Dt [H,ld4(B)], xt [V,ld4(B)], p0 [B,ld4(H)], p0t [H,ld4(B)], h0 [B,ld4(H)],
 *   v1 [B,ld4(V)], v1t [V,ld4(B)], p1t [H,ld4(B)], pc [B,ld4(C)], y1 int32[B] -- after qbm_rbm_cd1_step they hold ph0, the
 *   draws h0 / v1 / y1, p(y|h0) and ph1^T, which tests/test_gpu_rbm.py replays against oracle.rbm_cd1_step_replay */
int qbm_rbm_workspace_layout(int B, int V, int H, int C, long long *offsets);

/* ------------------------------------------------------------------------------------------
 * Measurement hook (no reference counterpart): on-chip peaks of the current device, the denominators of the
 * sampler's rooflines.  Streams FFMA2 / FFMA, conflict-free LDS.128 and L1-hit LDG.128 loops for a
 * few milliseconds each, timed with CUDA events; SYNCHRONISES `stream`.
 *   out     host double[6]: fp32 TFLOP/s with fma.rn.f32x2, fp32 TFLOP/s with fma.rn.f32, shared-memory TB/s, L1 TB/s,
 *           fp32 TFLOP/s of fma.rn.f32x2 with three distinct register operands at 8 warps per SM (the chain-tile kernel's
 *           row update), coefficient-major and row-major instruction order
 *   scratch device buffer, 16-byte aligned, at least 1 MiB + 16 bytes
 */
int qbm_probe_onchip_peaks(double *out, void *scratch, size_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QBM_B200_H */
