// Multi-kernel Gaussian MMD (reference: model/mmd.py:239-312, mix_rbf_mmd2 / _mix_rbf_kernel / _mmd2)
// and the Chamfer distance used by the SDA geometric weights (mmd.py:126-128,169-175).
//
// loss = sum_ij c_ij K_ij,  K_ij = sum_sigma exp(-E_ij / (2 sigma^2)),  E_ij = G_ii + G_jj - 2 G_ij,
// G = Z Z^T.  The squared norms are read from the Gram diagonal exactly like the reference, so the
// exponent on the diagonal is exactly 0 (with sigma = 0.01 a 1e-3 error there changes K by e^{+-5}).
//   biased:   c = 1/m^2 on XX and YY, -w_j/m^2 on XY and YX (w = SDA column weights, mmd.py:293-297)
//   unbiased: c = 1/(m(m-1)) off the diagonal of XX / YY, 0 on it
// The forward also emits coef = dL/dG so that the backward is one GEMM: dZ = 2 g coef Z.
#include "common.cuh"

namespace sug {

struct Sigmas {
  float gamma[8];
  int n;
};

__device__ __forceinline__ float mmd_cij(int i, int j, int m, const float* __restrict__ w, int biased) {
  const bool ix = i < m, jx = j < m;
  const float mm = (float)m * (float)m;
  if (ix == jx) {
    if (biased) return 1.f / mm;
    return i == j ? 0.f : 1.f / ((float)m * (float)(m - 1));
  }
  int col = ix ? j - m : i - m;  // index into Y
  float wj = w != nullptr ? __ldg(w + col) : 1.f;
  return -wj / mm;
}

// one block per row i of the [2m, 2m] kernel matrix
__global__ void __launch_bounds__(128)
mmd_coef_kernel(const float* __restrict__ G, int m, Sigmas sg, const float* __restrict__ w, int biased,
                float* __restrict__ coef, double* __restrict__ acc) {
  __shared__ double red_l[4], red_q[4];
  __shared__ float s_qii;
  const int n2 = 2 * m;
  const int i = blockIdx.x;
  const float gii = __ldg(G + (size_t)i * n2 + i);
  double lsum = 0.0, qsum = 0.0;
  for (int j = threadIdx.x; j < n2; j += blockDim.x) {
    const float gjj = __ldg(G + (size_t)j * n2 + j);
    const float gij = __ldg(G + (size_t)i * n2 + j);
    const float E = gii - 2.f * gij + gjj;  // mmd.py:247 operation order
    float K = 0.f, dK = 0.f;
    for (int s = 0; s < sg.n; ++s) {
      float e = expf(-sg.gamma[s] * E);
      K += e;
      dK = fmaf(-sg.gamma[s], e, dK);
    }
    const float c = mmd_cij(i, j, m, w, biased);
    const float q = c * dK;  // dL/dE_ij
    lsum += (double)(c * K);
    qsum += (double)q;
    if (j == i) s_qii = q;
    else coef[(size_t)i * n2 + j] = -2.f * q;
  }
  lsum = warp_sum(lsum);
  qsum = warp_sum(qsum);
  if ((threadIdx.x & 31) == 0) { red_l[threadIdx.x >> 5] = lsum; red_q[threadIdx.x >> 5] = qsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = red_l[0] + red_l[1] + red_l[2] + red_l[3];
    double q = red_q[0] + red_q[1] + red_q[2] + red_q[3];
    // E_ii = G_ii + G_ii - 2 G_ii: d/dG_ii collects the row and the column (Q symmetric)
    coef[(size_t)i * n2 + i] = (float)(2.0 * q - 2.0 * (double)s_qii);
    atomicAdd(acc, l);
  }
}

__global__ void mmd_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss) { *loss = (float)acc[0]; }

__global__ void scale_rows_kernel(float* __restrict__ d, long long ld, int rows, int cols, const float* __restrict__ g,
                                  float mul) {
  const float s = mul * __ldg(g);
  long long total = (long long)rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long r = e / cols;
    int c = (int)(e - r * cols);
    d[r * ld + c] *= s;
  }
}

// Squared distance from every point of p to its nearest point of q; tiles of q in shared memory.
__global__ void __launch_bounds__(128)
chamfer_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int N, int M, float* __restrict__ d1,
               float* __restrict__ d2) {
  __shared__ float sq[128 * 3];
  const int dir = blockIdx.z;
  const float* p = dir == 0 ? p1 : p2;
  const float* q = dir == 0 ? p2 : p1;
  const int np = dir == 0 ? N : M, nq = dir == 0 ? M : N;
  float* d = dir == 0 ? d1 : d2;
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (blockIdx.x * 128 >= np) return;
  const float* pb = p + (size_t)b * np * 3;
  const float* qb = q + (size_t)b * nq * 3;
  float x = 0.f, y = 0.f, z = 0.f;
  if (i < np) { x = pb[3 * i]; y = pb[3 * i + 1]; z = pb[3 * i + 2]; }
  float best = INFINITY;
  for (int j0 = 0; j0 < nq; j0 += 128) {
    const int cnt = min(128, nq - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += 128) sq[e] = qb[(size_t)j0 * 3 + e];
    __syncthreads();
    for (int j = 0; j < cnt; ++j) {
      float dx = x - sq[3 * j], dy = y - sq[3 * j + 1], dz = z - sq[3 * j + 2];
      float dd = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      best = fminf(best, dd);
    }
  }
  if (i < np) d[(size_t)b * np + i] = best;
}

}  // namespace sug

using namespace sug;

extern "C" size_t sug_mmd_ws_bytes(int m, int D) {
  (void)D;
  return align_up(sizeof(float) * 4 * (size_t)m * m, 256) + 512;
}

extern "C" int sug_mmd_rbf_fwd(const float* z, int64_t ldz, int m, int D, const float* h_sigmas, int nsig,
                               const float* weights, int biased, float* loss, float* coef, void* ws, size_t ws_bytes,
                               sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_CHECK_ARG(z && h_sigmas && loss && coef, "mmd_fwd: null pointer");
  SUG_CHECK_ARG(m > 0 && D > 0 && nsig > 0 && nsig <= 8, "mmd_fwd: bad shape m=%d D=%d nsig=%d", m, D, nsig);
  SUG_CHECK_ARG(biased || m > 1, "mmd_fwd: unbiased estimate needs m > 1");
  Workspace W(ws, ws_bytes);
  float* G = W.take<float>(4 * (size_t)m * m);
  double* acc = W.take<double>(2);
  if (!W.ok()) { set_error("mmd_fwd: workspace too small"); return SUG_E_WORKSPACE; }
  Sigmas sg;
  sg.n = nsig;
  for (int s = 0; s < nsig; ++s) sg.gamma[s] = (float)(1.0 / (2.0 * (double)h_sigmas[s] * (double)h_sigmas[s]));
  const int n2 = 2 * m;
  SUG_TRY(gemm_f32(z, ldz, 1, z, ldz, 1, nullptr, G, n2, n2, n2, D, 0, stream));
  SUG_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), stream));
  {
    ProfScope ps(KC_MMD, 12.0 * nsig * n2 * (double)n2, 8.0 * n2 * (double)n2, stream);
    mmd_coef_kernel<<<n2, 128, 0, stream>>>(G, m, sg, weights, biased, coef, acc);
  }
  SUG_LAUNCH_CHECK();
  {
    ProfScope ps(KC_MISC, 0, 0, stream);
    mmd_finalize_kernel<<<1, 1, 0, stream>>>(acc, loss);
  }
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_mmd_rbf_bwd(const float* z, int64_t ldz, int m, int D, const float* coef, const float* gloss,
                               float* dz, int64_t lddz, sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_CHECK_ARG(z && coef && gloss && dz, "mmd_bwd: null pointer");
  const int n2 = 2 * m;
  // dz = coef * z   ([2m, 2m] x [2m, D]),  then scaled by 2 * gloss
  SUG_TRY(gemm_f32(coef, n2, 1, z, 1, ldz, nullptr, dz, lddz, n2, D, n2, 0, stream));
  long long total = (long long)n2 * D;
  ProfScope ps(KC_MISC, 0, 0, stream);
  scale_rows_kernel<<<(int)min((long long)num_sms() * 4, (total + 255) / 256), 256, 0, stream>>>(dz, lddz, n2, D, gloss,
                                                                                               2.f);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_chamfer_f32(const float* p1, const float* p2, int B, int N, int M, float* d1, float* d2,
                               sug_stream_t stream_) {
  SUG_CHECK_ARG(p1 && p2 && d1 && d2, "chamfer: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && M > 0, "chamfer: bad shape");
  dim3 grid(cdiv(N > M ? N : M, 128), B, 2);
  ProfScope ps(KC_CHAMFER, 16.0 * B * (double)N * M, 12.0 * B * (N + M) + 4.0 * B * (N + M), (cudaStream_t)stream_);
  chamfer_kernel<<<grid, 128, 0, (cudaStream_t)stream_>>>(p1, p2, N, M, d1, d2);
  SUG_LAUNCH_CHECK();
  return 0;
}

// ---- class-weighted focal loss (model_utils.py:131-176), one block --------------------------------------
//   L = reduce_r  alpha_r * ( -(1 - p_r)^gamma * log p_r ),   p_r = softmax(preds_r)[label_r]
// forward: scalar loss (sum or mean over the R rows); backward: d preds.
namespace sug {

__device__ __forceinline__ float focal_row(const float* __restrict__ z, int C, int y, float& lse) {
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(z[c] - m);
  lse = m + logf(s);
  return z[y] - lse;  // log p
}

__global__ void focal_fwd_kernel(const float* __restrict__ preds, const long long* __restrict__ labels,
                                 const float* __restrict__ alpha_row, int R, int C, float gamma, int mean,
                                 float* __restrict__ loss) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float lse;
    const float logp = focal_row(preds + (size_t)r * C, C, (int)labels[r], lse);
    const float p = expf(logp);
    const float w = gamma == 0.f ? 1.f : powf(1.f - p, gamma);
    acc += (double)(alpha_row[r] * (-(w * logp)));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(mean ? red[0] / R : red[0]);
}

__global__ void focal_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ preds,
                                 const long long* __restrict__ labels, const float* __restrict__ alpha_row, int R, int C,
                                 float gamma, int mean, float* __restrict__ dpreds) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* z = preds + (size_t)r * C;
  const int y = (int)labels[r];
  float lse;
  const float logp = focal_row(z, C, y, lse);
  const float p = expf(logp);
  // dL/dz_j = alpha * [ gamma (1-p)^(gamma-1) p log p - (1-p)^gamma ] * (delta_jy - s_j)
  float coef;
  if (gamma == 0.f) coef = -1.f;
  else coef = gamma * powf(1.f - p, gamma - 1.f) * p * logp - powf(1.f - p, gamma);
  coef *= alpha_row[r] * (*gout) * (mean ? 1.f / (float)R : 1.f);
  for (int c = 0; c < C; ++c) {
    const float s = expf(z[c] - lse);
    dpreds[(size_t)r * C + c] = coef * ((c == y ? 1.f : 0.f) - s);
  }
}

}  // namespace sug

extern "C" int sug_focal_loss_fwd(const float* preds, const int64_t* labels, const float* alpha_row, int R, int C,
                                  float gamma, int mean, float* loss, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(preds && labels && alpha_row && loss && R > 0 && C > 0, "focal_loss_fwd: bad argument");
  ProfScope ps(KC_MISC, 0, 0, (cudaStream_t)stream);
  focal_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(preds, reinterpret_cast<const long long*>(labels), alpha_row, R, C,
                                                       gamma, mean, loss);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_focal_loss_bwd(const float* gout, const float* preds, const int64_t* labels, const float* alpha_row,
                                  int R, int C, float gamma, int mean, float* dpreds, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(gout && preds && labels && alpha_row && dpreds && R > 0 && C > 0, "focal_loss_bwd: bad argument");
  ProfScope ps(KC_MISC, 0, 0, (cudaStream_t)stream);
  focal_bwd_kernel<<<cdiv(R, 128), 128, 0, (cudaStream_t)stream>>>(gout, preds, reinterpret_cast<const long long*>(labels),
                                                                   alpha_row, R, C, gamma, mean, dpreds);
  SUG_LAUNCH_CHECK();
  return 0;
}

// ---- SDA semantic sample weights (mmd.py:134-148 + 151-153 + 198-201), one block ---------------------------
//   v = [softmax(pred) | onehot(label) * label_weight]          (per row, 2C entries)
//   x = (v_s + 1e-8) / sum_all(v_s + 1e-8),  y likewise for the target batch
//   dist_r = sum_c 0.5 (x log(x/y) - x + y) + 0.5 (y log(y/x) - y + x)
//   "mean2one":  w_r = dist_r * int(1 / mean_r dist_r)          (integer truncation as in the reference)
// and the assembly of the soft-MMD inputs  Z = [feat | onehot(label) * scale]  for both batches (mmd.py:56-66).
namespace sug {

__device__ __forceinline__ void softmax_row(const float* __restrict__ z, int C, float* __restrict__ out) {
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) { out[c] = expf(z[c] - m); s += out[c]; }
  for (int c = 0; c < C; ++c) out[c] = out[c] / s;
}

__device__ double block_sum256(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

template <int C>
__global__ void __launch_bounds__(256)
sda_sem_weights_kernel(const float* __restrict__ pred_s, const float* __restrict__ pred_t,
                       const long long* __restrict__ label_s, const long long* __restrict__ label_t, int m,
                       float label_weight, float* __restrict__ w) {
  __shared__ double red[256];
  const float eps = 1e-8f;
  double ts = 0.0, tt = 0.0;
  for (int r = threadIdx.x; r < m; r += 256) {
    float ps[C], pt[C];
    softmax_row(pred_s + (size_t)r * C, C, ps);
    softmax_row(pred_t + (size_t)r * C, C, pt);
    for (int c = 0; c < C; ++c) {
      ts += (double)(ps[c] + eps) + (double)((c == (int)label_s[r] ? label_weight : 0.f) + eps);
      tt += (double)(pt[c] + eps) + (double)((c == (int)label_t[r] ? label_weight : 0.f) + eps);
    }
  }
  const float Ss = (float)block_sum256(ts, red), St = (float)block_sum256(tt, red);
  double dsum = 0.0;
  for (int r = threadIdx.x; r < m; r += 256) {
    float ps[C], pt[C];
    softmax_row(pred_s + (size_t)r * C, C, ps);
    softmax_row(pred_t + (size_t)r * C, C, pt);
    float d = 0.f;
    for (int c = 0; c < 2 * C; ++c) {
      const float vs = c < C ? ps[c] : ((c - C) == (int)label_s[r] ? label_weight : 0.f);
      const float vt = c < C ? pt[c] : ((c - C) == (int)label_t[r] ? label_weight : 0.f);
      const float x = (vs + eps) / Ss, y = (vt + eps) / St;
      d += (x * logf(x / y) - x + y) * 0.5f + (y * logf(y / x) - y + x) * 0.5f;
    }
    w[r] = d;
    dsum += (double)d;
  }
  const float mean = (float)(block_sum256(dsum, red) / m);
  const float scale = (float)(int)(1.f / mean);
  for (int r = threadIdx.x; r < m; r += 256) w[r] *= scale;
}

__global__ void soft_mmd_assemble_kernel(const float* __restrict__ fs, const float* __restrict__ ft, long long lds,
                                         long long ldt, const long long* __restrict__ label_s,
                                         const long long* __restrict__ label_t, int m, int D, int NC, float scale,
                                         float* __restrict__ z) {
  const int r = blockIdx.x;  // 0 .. 2m-1
  const float* src = r < m ? fs + (size_t)r * lds : ft + (size_t)(r - m) * ldt;
  const int lab = (int)(r < m ? label_s[r] : label_t[r - m]);
  float* dst = z + (size_t)r * (D + NC);
  for (int c = threadIdx.x; c < D; c += blockDim.x) dst[c] = src[c];
  for (int c = threadIdx.x; c < NC; c += blockDim.x) dst[D + c] = c == lab ? scale : 0.f;
}

}  // namespace sug

extern "C" int sug_sda_sem_weights(const float* pred_s, const float* pred_t, const int64_t* label_s,
                                   const int64_t* label_t, int m, int C, float label_weight, float* w, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(pred_s && pred_t && label_s && label_t && w && m > 0, "sda_sem_weights: bad argument");
  SUG_CHECK_ARG(C == 10, "sda_sem_weights: C=%d (the reference's one-hot has 10 classes, common_utils.py:161)", C);
  ProfScope ps(KC_MISC, 0, 0, (cudaStream_t)stream);
  sda_sem_weights_kernel<10><<<1, 256, 0, (cudaStream_t)stream>>>(pred_s, pred_t, reinterpret_cast<const long long*>(label_s),
                                                                  reinterpret_cast<const long long*>(label_t), m, label_weight, w);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_soft_mmd_assemble(const float* feat_s, int64_t lds, const float* feat_t, int64_t ldt,
                                     const int64_t* label_s, const int64_t* label_t, int m, int D, int num_class,
                                     float scale, float* z, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(feat_s && feat_t && label_s && label_t && z && m > 0 && D > 0 && num_class > 0, "soft_mmd_assemble: bad argument");
  ProfScope ps(KC_MISC, 0, 8.0 * m * (2.0 * D + num_class), (cudaStream_t)stream);
  soft_mmd_assemble_kernel<<<2 * m, 256, 0, (cudaStream_t)stream>>>(feat_s, feat_t, lds, ldt,
                                                                   reinterpret_cast<const long long*>(label_s),
                                                                   reinterpret_cast<const long long*>(label_t), m, D,
                                                                   num_class, scale, z);
  SUG_LAUNCH_CHECK();
  return 0;
}
