/*
 * sug_b200.h — C ABI of libsug_b200.so, the B200 (sm_100a) implementation of the SUG
 * point-cloud encoder hot path.
 *
 * The reference (SiyuanHuang95/SUG) has no FFI for this path: it is plain PyTorch calls
 * (SURVEY.md §8b).  Each entry point below is therefore what a binding for the named reference
 * function would call; INTEGRATION.md shows the ctypes stub.  Conventions:
 *   - every pointer is a DEVICE pointer unless its name starts with h_;
 *   - tensors are row-major; "point-major" means one row per point, [B*N, C] with a row stride;
 *   - every call is asynchronous on `stream`, allocates nothing, and is CUDA-graph capturable;
 *     scratch memory is passed in as (ws, ws_bytes) and sized by the matching *_ws_bytes();
 *   - return value: 0 on success, a positive cudaError_t on a CUDA failure, a negative SUG_E_*
 *     on a bad argument.  sug_last_error() returns a message for the calling thread.
 */
#ifndef SUG_B200_H
#define SUG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sug_stream_t; /* cudaStream_t */

#define SUG_E_BADARG (-1)
#define SUG_E_WORKSPACE (-2)
#define SUG_E_UNSUPPORTED (-3)

#define SUG_ACT_LEAKY 0 /* LeakyReLU(slope); slope 0 == ReLU */
#define SUG_POOL_MAX 0     /* out[b, c]            = max_n act(bn(y))            (PointNet, Model.py:272-274) */
#define SUG_POOL_MAX_AVG 1 /* out[b, c], [b, Co+c] = max_n, mean_n act(bn(y))    (DGCNN,   Model.py:112-116) */

int sug_version(void);
const char* sug_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * knn(x, k)                                                    reference: model/model_utils.py:178-185
 * x is addressed as x[b*sb + n*sn + c*sc] (elements), so both the reference's [B,C,N] layout
 * (sb=C*N, sn=1, sc=N) and the point-major layout (sb=N*ld, sn=ld, sc=1) are accepted.
 * idx [B,N,k] int32, indices local to the cloud, nearest first (self included), i.e. the order
 * of topk(-|xi-xj|^2).  The N x N matrix is never written to memory.
 * ------------------------------------------------------------------------------------------- */
size_t sug_knn_ws_bytes(int B, int C, int N, int k);
int sug_knn_f32(const float* x, int B, int C, int N, int k, int64_t sb, int64_t sn, int64_t sc,
                int32_t* idx, void* ws, size_t ws_bytes, sug_stream_t stream);

/* Transposed neighbour graph used by the EdgeConv backward: for every point j the list of
 * (i, slot) with idx[i][slot] == j, in unspecified order.  rev_ptr [B, N+1] (offsets local to
 * the cloud), rev_edge [B, N*k] = (i << 8) | slot.  Requires k <= 255. */
int sug_knn_reverse(const int32_t* idx, int B, int N, int k, int32_t* rev_ptr, int32_t* rev_edge,
                    sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * EdgeConv = get_graph_feature -> Conv2d(2C->Cout,1x1,bias=False) -> BatchNorm2d -> LeakyReLU
 *            -> max over k                    reference: model_utils.py:188-210, 8-32; Model.py:88-109
 * Computed as y_ij = a_j + b_i with [a|b] = x * [W1 | W2-W1]^T (one per-point GEMM), a gather over
 * the k neighbours that produces the extreme pre-activation, its slot and the BatchNorm sums, and
 * a monotone finalize out = act(gamma*(ext-mean)*invstd + beta) (min instead of max where
 * gamma < 0).  x [B*N, C] point-major with row stride ldx; w [Cout, 2C] as in the reference's
 * conv.0.weight ([x_j-x_i ; x_i] channel order); out [B*N, Cout] with row stride ldo.
 * training != 0: batch statistics over all B*N*k edges, running stats updated (momentum, unbiased
 * variance); ab / ext / arg / ssum / save_mean_invstd are written for the backward.
 * training == 0: running statistics; ext, arg, ssum, save_mean_invstd may be NULL (ab is scratch).
 * ------------------------------------------------------------------------------------------- */
size_t sug_edgeconv_ws_bytes(int B, int N, int C, int Cout, int k);
/* (training != 0: running_mean / running_var may be NULL -- the momentum update is then left to the caller, who finds
 * the batch mean / invstd in save_mean_invstd; same for sug_mlp_pool_fwd and sug_linear_bn_act_fwd) */
int sug_edgeconv_fwd(const float* x, int64_t ldx, const int32_t* idx, const float* w,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     int B, int N, int C, int Cout, int k, float eps, float momentum, float slope,
                     int training, float* out, int64_t ldo, float* ab, float* ext, uint8_t* arg,
                     float* ssum, float* save_mean_invstd, void* ws, size_t ws_bytes,
                     sug_stream_t stream);

/* Backward of the block above.  gout [B*N, Cout] (row stride ldg).  Outputs: dx [B*N, C] (row
 * stride lddx; accumulate_dx != 0 adds into it), dw [Cout, 2C], dgamma, dbeta [Cout].
 * dab [B*N, 2*Cout] is scratch.  rev_ptr / rev_edge come from sug_knn_reverse. */
int sug_edgeconv_bwd(const float* gout, int64_t ldg, const float* x, int64_t ldx, const int32_t* idx,
                     const int32_t* rev_ptr, const int32_t* rev_edge, const float* w,
                     const float* gamma, const float* beta, const float* ab, const float* ext,
                     const uint8_t* arg, const float* ssum, const float* save_mean_invstd,
                     int B, int N, int C, int Cout, int k, float slope, float* dx, int64_t lddx,
                     int accumulate_dx, float* dw, float* dgamma, float* dbeta, float* dab,
                     void* ws, size_t ws_bytes, sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Shared MLP + BatchNorm + activation + global pool over the N points of each cloud.
 *   DGCNN tail:   Conv1d(512,512,bias=False) -> BatchNorm1d -> leaky_relu(0.2) -> max || avg
 *                                                               reference: Model.py:111-116
 *   PointNet:     conv_2d(128,1024) (+bias) -> BN -> ReLU -> max over N
 *                                                               reference: Model.py:245,272-274
 * x [B*N, Cin] (ldx), w [Cout, Cin], bias [Cout] or NULL.  y [B*N, Cout] holds the linear output
 * (needed by the backward).  out [B, Cout] (max) or [B, 2*Cout] (max || avg).  argext [B, Cout]
 * int32 = point index of the extreme.
 * ------------------------------------------------------------------------------------------- */
size_t sug_mlp_pool_ws_bytes(int B, int N, int Cin, int Cout);
int sug_mlp_pool_fwd(const float* x, int64_t ldx, const float* w, const float* bias,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     int B, int N, int Cin, int Cout, float eps, float momentum, float slope,
                     int pool, int training, float* y, float* out, int32_t* argext,
                     float* save_mean_invstd, void* ws, size_t ws_bytes, sug_stream_t stream);
/* gout has the layout of out.  y is overwritten with dL/dy.  dbias may be NULL. */
int sug_mlp_pool_bwd(const float* gout, const float* x, int64_t ldx, const float* w,
                     const float* bias, const float* gamma, const float* beta, float* y,
                     const int32_t* argext, const float* save_mean_invstd, int B, int N, int Cin,
                     int Cout, float slope, int pool, float* dx, int64_t lddx, int accumulate_dx,
                     float* dw, float* dbias, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                     sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * conv_2d over points without pooling: Conv2d(1x1)(+bias) -> BatchNorm2d -> ReLU / LeakyReLU
 *                                                        reference: model/model_utils.py:8-32
 * (adapt_layer_off.residual, PointNet conv1..conv4, T-Net conv2d1/2).  x [P, Cin] point-major,
 * y [P, Cout] keeps the linear output for the backward (which overwrites it with dL/dy),
 * out [P, Cout] with row stride ldo.
 * ------------------------------------------------------------------------------------------- */
int sug_linear_bn_act_fwd(const float* x, int64_t ldx, const float* w, const float* bias,
                          const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int64_t P, int Cin, int Cout, float eps, float momentum,
                          float slope, int training, float* y, float* out, int64_t ldo,
                          float* save_mean_invstd, void* ws, size_t ws_bytes, sug_stream_t stream);
int sug_linear_bn_act_bwd(const float* gout, int64_t ldg, const float* x, int64_t ldx, const float* w,
                          const float* gamma, const float* beta, float* y,
                          const float* save_mean_invstd, int64_t P, int Cin, int Cout, float slope,
                          float* dx, int64_t lddx, float* dw, float* dbias, float* dgamma,
                          float* dbeta, void* ws, size_t ws_bytes, sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * mix_rbf_mmd2(X, Y, sigma_list, biased=True, sample_weights=w)    reference: model/mmd.py:239-312
 * z [2m, D] = cat(X, Y) with row stride ldz.  weights [m] or NULL (column weights of K_XY,
 * mmd.py:293-297).  Squared norms are taken from the Gram diagonal exactly like mmd.py:245-247.
 * loss: one float.  coef [2m, 2m] receives dL/dG (G = z z^T) for the backward, which is
 * dz = gscale * 2 * coef * z.
 * ------------------------------------------------------------------------------------------- */
size_t sug_mmd_ws_bytes(int m, int D);
int sug_mmd_rbf_fwd(const float* z, int64_t ldz, int m, int D, const float* h_sigmas, int nsig,
                    const float* weights, int biased, float* loss, float* coef, void* ws,
                    size_t ws_bytes, sug_stream_t stream);
int sug_mmd_rbf_bwd(const float* z, int64_t ldz, int m, int D, const float* coef,
                    const float* gloss /* device scalar */, float* dz, int64_t lddz,
                    sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ChamferDistance()(p1, p2)[:2]      reference call sites: model/mmd.py:126-128,169-175
 * (third-party otaheri/chamfer_distance).  p1 [B,N,3], p2 [B,M,3] -> d1 [B,N], d2 [B,M]:
 * squared distance to the nearest point of the other cloud.
 * ------------------------------------------------------------------------------------------- */
int sug_chamfer_f32(const float* p1, const float* p2, int B, int N, int M, float* d1, float* d2,
                    sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Index builders of the self-adaptive node layer        reference: model/point_utils.py:5-165,
 * called from adapt_layer_off.forward, model_utils.py:103-128.  xyz is [B,3,N] (reference layout).
 *   sug_fps:        farthest point sampling from the given start index  (point_utils.py:5-26)
 *   sug_ball_query: lowest-index nsample points within radius, padded with the first hit
 *                                                                        (point_utils.py:86-106)
 *   sug_knn_query:  nsample nearest points of every query, ascending     (point_utils.py:107-108)
 *   sug_three_nn:   k nearest of the M nodes for every point, ascending  (point_utils.py:151-153)
 * All indices int32.  The squared distance is -2 s.d + |s|^2 + |d|^2 as in point_utils.py:112-131.
 * ------------------------------------------------------------------------------------------- */
int sug_fps(const float* xyz, int B, int N, int npoint, const int32_t* start, int32_t* out_idx,
            sug_stream_t stream);
int sug_ball_query(const float* xyz, const float* query, int B, int N, int S, float radius,
                   int nsample, int32_t* out_idx, sug_stream_t stream);
int sug_knn_query(const float* xyz, const float* query, int B, int N, int S, int nsample,
                  int32_t* out_idx, sug_stream_t stream);
int sug_three_nn(const float* xyz, const float* nodes, int B, int N, int M, int k,
                 int32_t* out_idx, sug_stream_t stream);
/* sug_knn_query without the ordering guarantee: the nsample nearest points as a set (what
 * adapt_layer_off needs: the group is max-pooled, model_utils.py:121-123). */
int sug_knn_query_set(const float* xyz, const float* query, int B, int N, int S, int nsample,
                      int32_t* out_idx, sug_stream_t stream);
/* index_points(residual_fea, group_idx).max(-1), model_utils.py:122-123, fused: x [B,N,C] point-major,
 * idx [B,S,K] -> out [B,S,C], arg [B,S,C] (point index of the max).  Backward: dx (zeroed by the
 * caller) += scatter of g to arg. */
int sug_group_max_fwd(const float* x, const int32_t* idx, int B, int N, int S, int K, int C, float* out,
                      int32_t* arg, sug_stream_t stream);
int sug_group_max_bwd(const float* g, const int32_t* arg, int B, int N, int S, int C, float* dx,
                      sug_stream_t stream);
/* torch.sum(index_points(points2, idx) * weight, dim=3), point_utils.py:158-160, fused:
 * f [B,S,C], idx / w [B,N,K] -> out [B,N,C]; backward accumulates into zeroed df [B,S,C], dw [B,N,K]. */
int sug_interp_fwd(const float* f, const int32_t* idx, const float* w, int B, int N, int S, int K, int C,
                   float* out, sug_stream_t stream);
int sug_interp_bwd(const float* g, const float* f, const int32_t* idx, const float* w, int B, int N, int S,
                   int K, int C, float* df, float* dw, sug_stream_t stream);

/* Node offsets (model_utils.py:107-117: tanh(pred_offset(group - centre)) * (group_xyz - centre_xyz),
 * mean over the ball) and the 3-NN inverse-squared-distance interpolation weights
 * (point_utils.py:141-160), each fused into one kernel with its backward.
 *   h [B,N,3] = pred_offset weights applied per point (the conv is linear and bias-free),
 *   xyz [B,3,N], fidx [B,S] centres, gidx [B,S,G] ball members -> out [B,S,3]; dh [B,N,3] zeroed by the caller.
 *   nodes [B,S,3], idx [B,N,K] (K <= 8) -> w [B,N,K]; dnodes [B,S,3] zeroed by the caller.
 * The backward kernels scatter with float atomics (summation order not fixed). */
int sug_node_offset_fwd(const float* h, const float* xyz, const int32_t* fidx, const int32_t* gidx, int B, int N,
                        int S, int G, float* out, sug_stream_t stream);
int sug_node_offset_bwd(const float* gout, const float* h, const float* xyz, const int32_t* fidx,
                        const int32_t* gidx, int B, int N, int S, int G, float* dh, sug_stream_t stream);
int sug_interp_weight_fwd(const float* xyz, const float* nodes, const int32_t* idx, int B, int N, int S, int K,
                          float* w, sug_stream_t stream);
int sug_interp_weight_bwd(const float* gw, const float* xyz, const float* nodes, const int32_t* idx, int B, int N,
                          int S, int K, float* dnodes, sug_stream_t stream);

/* Plain fp32 GEMM used inside the entry points above, exported for tests:
 * C[M,N] = A * B^T (+ bias[n]) with A(m,k) = a[m*sam + k*sak], B(n,k) = b[n*sbn + k*sbk]. */
int sug_gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                 const float* bias, float* c, int64_t ldc, int M, int N, int K, int accumulate,
                 sug_stream_t stream);

/* sug_gemm_f32's problem statement routed through the library's dispatcher: tcgen05 3xTF32 when the
 * operands satisfy the TMA constraints, the exact CUDA-core kernel otherwise.  Used by the host
 * layer for the bias-only linear layers (Conv1d(128,64,1), Model.py:70,101). */
int sug_gemm_auto_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                      const float* bias, float* c, int64_t ldc, int M, int N, int K,
                      sug_stream_t stream);

/* fp32-accurate tensor-core GEMM (tcgen05 kind::tf32, 3-term hi/lo split, TMA-fed, TMEM
 * accumulators): C[M,N] = A * B^T (+ bias).  a_mn_major == 0: a is [M,K] row-major (stride lda);
 * a_mn_major != 0: a is [K,M] row-major (stride lda), i.e. the operand is consumed transposed; same
 * for b / N.  Bases must be 16 B aligned and lda / ldb multiples of 4.  Exported for tests; the
 * entry points above use it internally for every GEMM with K >= 16. */
int sug_gemm_tc_f32(const float* a, int64_t lda, int a_mn_major, const float* b, int64_t ldb,
                    int b_mn_major, const float* bias, float* c, int64_t ldc, int M, int N, int K,
                    sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SDA / MSA glue of model/mmd.py, each one launch:
 *   sug_sda_sem_weights:   prob_weights_soft(..., weighting="mean2one") (mmd.py:134-148, 151-153, 198-201):
 *                          softmax(pred) || onehot(label)*label_weight, batch-normalised, symmetric KL per row,
 *                          scaled by int(1 / mean).  pred_* [m,10], label_* int64 [m] -> w [m].  No gradient
 *                          (the reference detaches the predictions).
 *   sug_soft_mmd_assemble: the operand of soft_mmd (mmd.py:56-66): z [2m, D+num_class] =
 *                          [feat_s | onehot(label_s)*scale ; feat_t | onehot(label_t)*scale].
 * ------------------------------------------------------------------------------------------- */
int sug_sda_sem_weights(const float* pred_s, const float* pred_t, const int64_t* label_s, const int64_t* label_t, int m,
                        int C, float label_weight, float* w, sug_stream_t stream);
int sug_soft_mmd_assemble(const float* feat_s, int64_t lds, const float* feat_t, int64_t ldt, const int64_t* label_s,
                          const int64_t* label_t, int m, int D, int num_class, float scale, float* z,
                          sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Class-weighted focal loss of the trainer (model_utils.py:131-176; gamma = 0 is the weighted cross entropy
 * the SUG configs use):  L = reduce_r alpha_row[r] * ( -(1 - p_r)^gamma log p_r ),  p_r = softmax(preds_r)[labels_r],
 * reduce = mean (mean != 0) or sum.  preds [R,C] row-major, labels int64 [R], alpha_row [R] (the per-row
 * weights, i.e. the reference's alpha after its gather), loss: device scalar.  Backward: gout device scalar.
 * ------------------------------------------------------------------------------------------- */
int sug_focal_loss_fwd(const float* preds, const int64_t* labels, const float* alpha_row, int R, int C, float gamma,
                       int mean, float* loss, sug_stream_t stream);
int sug_focal_loss_bwd(const float* gout, const float* preds, const int64_t* labels, const float* alpha_row, int R,
                       int C, float gamma, int mean, float* dpreds, sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step of the trainer (train_dg_single_gpu.py:191-203, 329-335: three torch.optim.Adam).
 *   sug_adam_f32: one multi-tensor Adam update (L2 weight decay folded into the gradient, bias
 *                 corrected, torch's capturable arithmetic).  p/g/m/v_ptrs and sizes are DEVICE arrays
 *                 with one entry per tensor (addresses as int64); blk_tensor/blk_chunk map every
 *                 thread block to (tensor, chunk of sug_adam_chunk() elements).  `step` (device
 *                 float) is incremented first, `lr` is read from device memory, so the call can be
 *                 captured in a CUDA graph and follows LR schedulers without re-capture.
 * ------------------------------------------------------------------------------------------- */
/* sug_adam_multi_f32: the same update for ALL param groups of an optimizer in one launch (the trainer's optimizer_g has
 * one group per parameter, train_dg_single_gpu.py:191).  Per TENSOR: step_ptrs / lr_ptrs = device addresses of its group's
 * step counter and learning rate, hyper[4*t .. 4*t+3] = (beta1, beta2, eps, weight_decay); group_step_ptrs [n_groups] =
 * the distinct step counters, each incremented once before the update.  torch.optim.Adam semantics per group. */
int sug_adam_multi_f32(const int64_t* p_ptrs, const int64_t* g_ptrs, const int64_t* m_ptrs, const int64_t* v_ptrs,
                       const int64_t* sizes, const int64_t* step_ptrs, const int64_t* lr_ptrs, const float* hyper,
                       const int32_t* blk_tensor, const int32_t* blk_chunk, int n_blocks, long long n_params,
                       const int64_t* group_step_ptrs, int n_groups, sug_stream_t stream);
int sug_adam_chunk(void);
int sug_adam_f32(const int64_t* p_ptrs, const int64_t* g_ptrs, const int64_t* m_ptrs, const int64_t* v_ptrs,
                 const int64_t* sizes, const int32_t* blk_tensor, const int32_t* blk_chunk, int n_blocks,
                 long long n_params, float* step, const float* lr, float beta1, float beta2, float eps,
                 float weight_decay, sug_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Launch accounting used by bench.py (`gpu_launches`, `roofline`).  Every kernel launch of the
 * library is counted per kernel class together with its algorithmic flops / bytes (formulas in
 * DESIGN.md).  sug_prof_enable(mask) additionally brackets the launches of the selected classes
 * with CUDA events on their own stream; sug_prof_collect synchronises the device and returns, per
 * class, the timed milliseconds, the number of timed launches, all launches, flops and bytes.
 * ------------------------------------------------------------------------------------------- */
int sug_prof_num_classes(void);
const char* sug_prof_class_name(int i);
void sug_prof_enable(unsigned mask);
void sug_prof_reset(void);
int sug_prof_collect(double* ms, long long* timed, long long* launches, double* flops, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* SUG_B200_H */
