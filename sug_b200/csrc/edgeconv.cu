// EdgeConv with exact train-mode BatchNorm, without ever forming the [B,2C,N,k] edge tensor.
// Reference: model/model_utils.py:188-210 (get_graph_feature), 8-32 (conv_2d), Model.py:88-109.
//
//   y_ij = W [x_j - x_i ; x_i] = a_j + b_i,   a = W1 x,  b = (W2 - W1) x          (per-point GEMM)
//   ext_i = max_j y_ij (min_j where gamma < 0),  S_i = sum_j y_ij,  sum / sum^2 over all edges
//   out_i = LeakyReLU(gamma (ext_i - mean) invstd + beta)     (BN affine o LeakyReLU is monotone)
//
// Backward: with ghat_i = g_i LReLU'(z_i), G1 = sum ghat, G2 = sum ghat yhat*,
//   dL/dy_ij = scale [ ghat_i [j = j*_i] - G1/M - yhat_ij G2/M ],  scale = gamma invstd, M = B N k
//   dB_i = scale [ ghat_i - k G1/M - (G2/M) invstd (S_i - k mean) ]
//   dA_j = scale [ sum_{(i,s) -> j, arg_i = s} ghat_i - deg_j G1/M
//                  - (G2/M) invstd (deg_j (a_j - mean) + sum_{i -> j} b_i) ]
// evaluated as a gather over the transposed neighbour graph (no atomics, deterministic).
#include "common.cuh"

namespace sug {

__global__ void edge_pack_weight_kernel(const float* __restrict__ w, int C, int Cout, float* __restrict__ wcat) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Cout * C) return;
  int o = e / C, c = e % C;
  float w1 = w[(size_t)o * 2 * C + c];
  float w2 = w[(size_t)o * 2 * C + C + c];
  wcat[(size_t)o * C + c] = w1;
  wcat[(size_t)(Cout + o) * C + c] = w2 - w1;
}

__global__ void edge_unpack_wgrad_kernel(const float* __restrict__ dwcat, int C, int Cout, float* __restrict__ dw) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Cout * C) return;
  int o = e / C, c = e % C;
  float da = dwcat[(size_t)o * C + c];
  float db = dwcat[(size_t)(Cout + o) * C + c];
  dw[(size_t)o * 2 * C + c] = da - db;
  dw[(size_t)o * 2 * C + C + c] = db;
}

// One thread per (point, 4 channels).  TRAIN: writes ext/arg/ssum and accumulates the BN sums.
// !TRAIN: statistics are known (running), so the activation is applied here and `out` written.
template <bool TRAIN>
__global__ void __launch_bounds__(256)
edge_gather_fwd_kernel(const float* __restrict__ ab, const int* __restrict__ idx, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ mean_invstd, int P, int N, int k,
                       int Cout, float slope, float* __restrict__ ext, uint8_t* __restrict__ arg,
                       float* __restrict__ ssum, double* __restrict__ sums, float* __restrict__ out, long long ldo) {
  __shared__ double red[256][8];
  const int CQ = Cout >> 2;
  const int PPB = 256 / CQ;
  const int tid = threadIdx.x;
  const int pl = tid / CQ, c4 = tid - pl * CQ;
  const bool active = pl < PPB;
  const int ld = 2 * Cout;
  float sgn[4], sc[4], sh[4];
  if (active) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float g = __ldg(gamma + 4 * c4 + u);
      sgn[u] = g < 0.f ? -1.f : 1.f;
      if (!TRAIN) {
        float m = __ldg(mean_invstd + 4 * c4 + u), is = __ldg(mean_invstd + Cout + 4 * c4 + u);
        sc[u] = g * is;
        sh[u] = __ldg(beta + 4 * c4 + u) - m * sc[u];
      }
    }
  }
  double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
  for (long long p0 = (long long)blockIdx.x * PPB; p0 < P; p0 += (long long)gridDim.x * PPB) {
    const long long i = p0 + pl;
    if (!active || i >= P) continue;
    const long long cb = (i / N) * N;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(ab + i * ld + Cout) + c4);
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bslot[4] = {0, 0, 0, 0};
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const int* ip = idx + i * k;
    for (int s0 = 0; s0 < k; s0 += 4) {
      float4 a4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int s = min(s0 + u, k - 1);
        long long j = cb + __ldg(ip + s);
        a4[u] = __ldg(reinterpret_cast<const float4*>(ab + j * ld) + c4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (s0 + u >= k) break;
        const float aa[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float y = aa[v] + bb[v];
          float ys = y * sgn[v];
          if (ys > best[v]) { best[v] = ys; bslot[v] = s0 + u; }
          if (TRAIN) {
            s1[v] += y;
            s2[v] = fmaf(y, y, s2[v]);
          }
        }
      }
    }
    float e[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) e[v] = best[v] * sgn[v];
    if (TRAIN) {
      reinterpret_cast<float4*>(ext + i * Cout)[c4] = make_float4(e[0], e[1], e[2], e[3]);
      reinterpret_cast<float4*>(ssum + i * Cout)[c4] = make_float4(s1[0], s1[1], s1[2], s1[3]);
      reinterpret_cast<uchar4*>(arg + i * Cout)[c4] =
          make_uchar4((unsigned char)bslot[0], (unsigned char)bslot[1], (unsigned char)bslot[2], (unsigned char)bslot[3]);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        ds[v] += (double)s1[v];
        dq[v] += (double)s2[v];
      }
    } else {
      float4 o;
      o.x = act_leaky(fmaf(sc[0], e[0], sh[0]), slope);
      o.y = act_leaky(fmaf(sc[1], e[1], sh[1]), slope);
      o.z = act_leaky(fmaf(sc[2], e[2], sh[2]), slope);
      o.w = act_leaky(fmaf(sc[3], e[3], sh[3]), slope);
      *reinterpret_cast<float4*>(out + i * ldo + 4 * c4) = o;
    }
  }
  if (TRAIN) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      red[tid][v] = ds[v];
      red[tid][4 + v] = dq[v];
    }
    __syncthreads();
    if (active && pl == 0) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double a = 0, q = 0;
        for (int r = 0; r < PPB; ++r) {
          a += red[r * CQ + c4][v];
          q += red[r * CQ + c4][4 + v];
        }
        atomicAdd(&sums[4 * c4 + v], a);
        atomicAdd(&sums[Cout + 4 * c4 + v], q);
      }
    }
  }
}

// out = act(gamma (ext - mean) invstd + beta), float4 per thread.
// SUMS: the batch statistics come as fp64 (sum, sum of squares) over `count` samples and are finalised here by every
// block for itself (Cout divisions); block 0 also writes (mean, invstd) for the backward and updates the running
// statistics exactly like nn.BatchNorm (momentum, unbiased variance) -- no separate finalize launch.
template <bool SUMS>
__global__ void __launch_bounds__(256)
bn_act_kernel_t(const float* __restrict__ ext, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ mean_invstd, const double* __restrict__ sums, double count, float eps,
                float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                float* __restrict__ save, long long P, int Cout, float slope, float* __restrict__ out, long long ldo) {
  extern __shared__ float ss[];  // scale[Cout], shift[Cout]
  for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
    float mean_f, is_f;
    if (SUMS) {
      const double mean = sums[c] / count;
      double var = sums[Cout + c] / count - mean * mean;
      if (var < 0.0) var = 0.0;
      mean_f = (float)mean;
      is_f = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {
        save[c] = mean_f;
        save[Cout + c] = is_f;
        if (running_mean != nullptr) {
          const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
          running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
          running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
        }
      }
    } else {
      mean_f = mean_invstd[c];
      is_f = mean_invstd[Cout + c];
    }
    const float sc = gamma[c] * is_f;
    ss[c] = sc;
    ss[Cout + c] = beta[c] - mean_f * sc;
  }
  __syncthreads();
  const int CQ = Cout >> 2;
  const long long total = P * CQ;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long i = e / CQ;
    int c4 = (int)(e - i * CQ);
    float4 v = __ldg(reinterpret_cast<const float4*>(ext + i * Cout) + c4);
    const float* sc = ss + 4 * c4;
    const float* sh = ss + Cout + 4 * c4;
    float4 o;
    o.x = act_leaky(fmaf(sc[0], v.x, sh[0]), slope);
    o.y = act_leaky(fmaf(sc[1], v.y, sh[1]), slope);
    o.z = act_leaky(fmaf(sc[2], v.z, sh[2]), slope);
    o.w = act_leaky(fmaf(sc[3], v.w, sh[3]), slope);
    *reinterpret_cast<float4*>(out + i * ldo + 4 * c4) = o;
  }
}

// launchers shared with pool.cu
int bn_act_launch(const float* y, const float* gamma, const float* beta, const float* mean_invstd, long long P, int Cout,
                  float slope, float* out, long long ldo, cudaStream_t stream) {
  const long long total = P * (Cout >> 2);
  const int g2 = (int)min((long long)num_sms() * 8, (total + 255) / 256);
  ProfScope ps(KC_BN_ACT, 2.0 * P * Cout, 8.0 * P * Cout, stream);
  bn_act_kernel_t<false><<<g2, 256, 2 * Cout * sizeof(float), stream>>>(y, gamma, beta, mean_invstd, nullptr, 0.0, 0.f, 0.f,
                                                                       nullptr, nullptr, nullptr, P, Cout, slope, out, ldo);
  SUG_LAUNCH_CHECK();
  return 0;
}
int bn_act_from_sums_launch(const float* y, const float* gamma, const float* beta, const double* sums, double count, float eps,
                            float momentum, float* running_mean, float* running_var, float* save, long long P, int Cout,
                            float slope, float* out, long long ldo, cudaStream_t stream) {
  const long long total = P * (Cout >> 2);
  const int g2 = (int)min((long long)num_sms() * 8, (total + 255) / 256);
  ProfScope ps(KC_BN_ACT, 2.0 * P * Cout, 8.0 * P * Cout, stream);
  bn_act_kernel_t<true><<<g2, 256, 2 * Cout * sizeof(float), stream>>>(y, gamma, beta, nullptr, sums, count, eps, momentum,
                                                                      running_mean, running_var, save, P, Cout, slope, out, ldo);
  SUG_LAUNCH_CHECK();
  return 0;
}

// ghat = g * act'(z);  G1 += ghat;  G2 += ghat * yhat.
__global__ void __launch_bounds__(256)
edge_bwd_pre_kernel(const float* __restrict__ gout, long long ldg, const float* __restrict__ ext,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ mean_invstd, long long P, int Cout, float slope,
                    float* __restrict__ ghat, double* __restrict__ gsums) {
  __shared__ double red[256][8];
  const int CQ = Cout >> 2;
  const int PPB = 256 / CQ;
  const int tid = threadIdx.x;
  const int pl = tid / CQ, c4 = tid - pl * CQ;
  const bool active = pl < PPB;
  float mean[4], is[4], sc[4], sh[4];
  if (active) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int c = 4 * c4 + u;
      mean[u] = __ldg(mean_invstd + c);
      is[u] = __ldg(mean_invstd + Cout + c);
      sc[u] = __ldg(gamma + c) * is[u];
      sh[u] = __ldg(beta + c) - mean[u] * sc[u];
    }
  }
  double g1[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};
  for (long long p0 = (long long)blockIdx.x * PPB; p0 < P; p0 += (long long)gridDim.x * PPB) {
    const long long i = p0 + pl;
    if (!active || i >= P) continue;
    float4 g = __ldg(reinterpret_cast<const float4*>(gout + i * ldg) + c4);
    float4 e = __ldg(reinterpret_cast<const float4*>(ext + i * Cout) + c4);
    const float gg[4] = {g.x, g.y, g.z, g.w};
    const float ee[4] = {e.x, e.y, e.z, e.w};
    float gh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float z = fmaf(sc[u], ee[u], sh[u]);
      gh[u] = gg[u] * act_leaky_grad(z, slope);
      g1[u] += (double)gh[u];
      g2[u] += (double)(gh[u] * ((ee[u] - mean[u]) * is[u]));
    }
    reinterpret_cast<float4*>(ghat + i * Cout)[c4] = make_float4(gh[0], gh[1], gh[2], gh[3]);
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    red[tid][v] = g1[v];
    red[tid][4 + v] = g2[v];
  }
  __syncthreads();
  if (active && pl == 0) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double a = 0, q = 0;
      for (int r = 0; r < PPB; ++r) {
        a += red[r * CQ + c4][v];
        q += red[r * CQ + c4][4 + v];
      }
      atomicAdd(&gsums[4 * c4 + v], a);
      atomicAdd(&gsums[Cout + 4 * c4 + v], q);
    }
  }
}

// dA | dB per point from the transposed graph.
__global__ void __launch_bounds__(256)
edge_bwd_main_kernel(const float* __restrict__ ab, const float* __restrict__ ghat, const uint8_t* __restrict__ arg,
                     const float* __restrict__ ssum, const int* __restrict__ rev_ptr, const int* __restrict__ rev_edge,
                     const float* __restrict__ gamma, const float* __restrict__ mean_invstd,
                     const double* __restrict__ gsums, long long P, int N, int k, int Cout, float* __restrict__ dab,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int CQ = Cout >> 2;
  const int PPB = 256 / CQ;
  const int tid = threadIdx.x;
  const int pl = tid / CQ, c4 = tid - pl * CQ;
  if (pl >= PPB) return;
  const int ld = 2 * Cout;
  const double Md = (double)P * (double)k;
  float mean[4], is[4], sc[4], c1[4], c2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int c = 4 * c4 + u;
    mean[u] = __ldg(mean_invstd + c);
    is[u] = __ldg(mean_invstd + Cout + c);
    sc[u] = __ldg(gamma + c) * is[u];
    double G1 = gsums[c], G2 = gsums[Cout + c];
    c1[u] = (float)(G1 / Md);
    c2[u] = (float)(G2 / Md) * is[u];
    if (blockIdx.x == 0 && pl == 0) {
      dbeta[c] = (float)G1;
      dgamma[c] = (float)G2;
    }
  }
  const float kf = (float)k;
  for (long long p0 = (long long)blockIdx.x * PPB; p0 < P; p0 += (long long)gridDim.x * PPB) {
    const long long j = p0 + pl;
    if (j >= P) continue;
    const long long bidx = j / N;
    const long long cb = bidx * N;
    const int jl = (int)(j - cb);
    const int* rp = rev_ptr + bidx * (N + 1);
    const int lo = __ldg(rp + jl), hi = __ldg(rp + jl + 1);
    const int* re = rev_edge + cb * k;
    float T[4] = {0, 0, 0, 0}, Gs[4] = {0, 0, 0, 0};
    for (int t = lo; t < hi; ++t) {
      int pk = __ldg(re + t);
      long long i = cb + (pk >> 8);
      unsigned s = (unsigned)(pk & 255);
      float4 b4 = __ldg(reinterpret_cast<const float4*>(ab + i * ld + Cout) + c4);
      float4 g4 = __ldg(reinterpret_cast<const float4*>(ghat + i * Cout) + c4);
      uchar4 a4 = __ldg(reinterpret_cast<const uchar4*>(arg + i * Cout) + c4);
      T[0] += b4.x; T[1] += b4.y; T[2] += b4.z; T[3] += b4.w;
      Gs[0] += (a4.x == s) ? g4.x : 0.f;
      Gs[1] += (a4.y == s) ? g4.y : 0.f;
      Gs[2] += (a4.z == s) ? g4.z : 0.f;
      Gs[3] += (a4.w == s) ? g4.w : 0.f;
    }
    const float deg = (float)(hi - lo);
    float4 a4 = __ldg(reinterpret_cast<const float4*>(ab + j * ld) + c4);
    float4 gh = __ldg(reinterpret_cast<const float4*>(ghat + j * Cout) + c4);
    float4 S4 = __ldg(reinterpret_cast<const float4*>(ssum + j * Cout) + c4);
    const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
    const float gg[4] = {gh.x, gh.y, gh.z, gh.w};
    const float SS[4] = {S4.x, S4.y, S4.z, S4.w};
    float dA[4], dB[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dB[u] = sc[u] * (gg[u] - kf * c1[u] - c2[u] * (SS[u] - kf * mean[u]));
      dA[u] = sc[u] * (Gs[u] - deg * c1[u] - c2[u] * (deg * (aa[u] - mean[u]) + T[u]));
    }
    reinterpret_cast<float4*>(dab + j * ld)[c4] = make_float4(dA[0], dA[1], dA[2], dA[3]);
    reinterpret_cast<float4*>(dab + j * ld + Cout)[c4] = make_float4(dB[0], dB[1], dB[2], dB[3]);
  }
}

// ---- shared-memory staged forward gather ------------------------------------------------------------
// One CTA per (cloud, chunk of CH channels).  The cloud's `a` rows for the chunk are staged once in
// shared memory, pre-multiplied by sign(gamma) so that "extreme" is always a max.  Because
// y_ij = a_j + b_i with b_i constant over the k neighbours, the arg-max over y is the arg-max over a,
// and the BatchNorm sums factor as
//     sum y   = sum_i (S_i + k b_i),                         S_i = sum_{j in N(i)} a_j,  Q_i = sum_{j in N(i)} a_j^2
//     sum y^2 = sum_i (Q_i + 2 b_i S_i + k b_i^2)
// so the per-edge work is one conflict-free LDS.128 and, per channel, compare + 2 selects + 1 add.
// Lanes own 4 channels; a warp serves 32*4/CH points per step; neighbour lists are staged per warp.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// KB: bits of the arg-slot code embedded in the keys (0 = exact compare / select path)
template <int CH, bool TRAIN, int KB>
__global__ void __launch_bounds__(512, 1)
edge_gather_smem_kernel(const float* __restrict__ ab, const int* __restrict__ idx,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ mean_invstd, int N, int k, int Cout, float slope,
                        float* __restrict__ ext, uint8_t* __restrict__ arg, float* __restrict__ ssum,
                        double* __restrict__ sums, float* __restrict__ out, long long ldo) {
  constexpr int Q = CH / 4;      // float4 lanes per point
  constexpr int PPW = 32 / Q;    // points per warp step
  extern __shared__ __align__(16) float smem_f[];
  float* As = smem_f;                                        // [N][CH]
  int* widx = reinterpret_cast<int*>(smem_f + (size_t)N * CH);  // [16 warps][2][PPW * k]
  __shared__ double red[16][Q][8];                           // stats partials per (warp, quad)
  const int c0 = blockIdx.x * CH;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = lane % Q, sub = lane / Q;
  const int ld = 2 * Cout;
  const long long cb = (long long)b * N;
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + (tid % Q));
  // ---- stage the signed A chunk (four independent loads in flight per thread) ----
  for (int e0 = tid; e0 < N * Q; e0 += 4 * 512) {
    float4 vv[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int e = e0 + w * 512;
      const int n = e / Q, qq = e - n * Q;
      if (e < N * Q) vv[w] = __ldg(reinterpret_cast<const float4*>(ab + (cb + n) * ld + c0) + qq);
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int e = e0 + w * 512;
      if (e >= N * Q) break;
      const int n = e / Q, qq = e - n * Q;
      float4 v = vv[w];
      v.x = g4.x < 0.f ? -v.x : v.x;
      v.y = g4.y < 0.f ? -v.y : v.y;
      v.z = g4.z < 0.f ? -v.z : v.z;
      v.w = g4.w < 0.f ? -v.w : v.w;
      reinterpret_cast<float4*>(As + (size_t)n * CH)[qq] = v;
    }
  }
  // per-lane constants for the channels this lane owns in the gather phase
  const float4 gq = __ldg(reinterpret_cast<const float4*>(gamma + c0) + q);
  const float sg[4] = {gq.x < 0.f ? -1.f : 1.f, gq.y < 0.f ? -1.f : 1.f, gq.z < 0.f ? -1.f : 1.f, gq.w < 0.f ? -1.f : 1.f};
  float sc[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0};
  if (!TRAIN) {
    const float gg[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 4 * q + u;
      const float m = __ldg(mean_invstd + c), is = __ldg(mean_invstd + Cout + c);
      sc[u] = gg[u] * is;
      sh[u] = __ldg(beta + c) - m * sc[u];
    }
  }
  __syncthreads();
  double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
  // neighbour lists: PPW consecutive points => PPW*k consecutive ints, double buffered per warp -- the lists of the
  // NEXT step are requested before the current step's gather and parked in shared memory after it, so their global
  // latency is hidden behind ~20 neighbours of work
  int* wbuf = widx + warp * (2 * PPW * k);
  const float kf = (float)k;
  constexpr int NPF = (PPW * 64 + 31) / 32;  // prefetch registers per lane (k <= 64; larger k re-loads synchronously)
  const bool can_pf = PPW * k <= NPF * 32;
  {
    const int p0 = warp * PPW;
    if (p0 < N) {
      const int cnt = min(PPW, N - p0) * k;
      for (int e = lane; e < cnt; e += 32) wbuf[e] = __ldg(idx + (cb + p0) * k + e);
    }
    __syncwarp();
  }
  int cur = 0;
  for (int p0 = warp * PPW; p0 < N; p0 += 16 * PPW, cur ^= 1) {
    int* wi = wbuf + cur * (PPW * k);
    const int pn = p0 + 16 * PPW;
    const int cntn = pn < N ? min(PPW, N - pn) * k : 0;
    int pf[NPF];
    if (can_pf) {
#pragma unroll
      for (int u = 0; u < NPF; ++u) pf[u] = (lane + 32 * u < cntn) ? __ldg(idx + (cb + pn) * k + lane + 32 * u) : 0;
    }
    const int i = p0 + sub;
    if (i < N) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(ab + (cb + i) * ld + Cout + c0) + q);
      float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int bslot[4] = {0, 0, 0, 0};
      float s1[4] = {0, 0, 0, 0};
      float q1[4] = {0, 0, 0, 0};  // sum of a_j^2 over the neighbours: FMA-pipe work, the ALU pipe bounds the loop
      const int* myi = wi + sub * k;
      if (KB == 0) {
        // exact arg: compare + two selects per neighbour and channel (any k <= 255)
#pragma unroll 4
        for (int s = 0; s < k; ++s) {
          const int j = myi[s];
          const float4 a4 = reinterpret_cast<const float4*>(As + (size_t)j * CH)[q];
          const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool gt = aa[u] > best[u];
            best[u] = gt ? aa[u] : best[u];
            bslot[u] = gt ? s : bslot[u];
            if (TRAIN) {
              s1[u] += aa[u];
              q1[u] = fmaf(aa[u], aa[u], q1[u]);
            }
          }
        }
      } else {
        // The ALU pipe bounds this loop (measured: 59 % ALU, 20 % FMA pipe).  Neighbours are taken in pairs: the exact
        // maximum with one 3-input FMNMX3 per pair, and the arg slot from a second FMNMX3 over KEYS = the value with
        // its KB lowest mantissa bits replaced by (2^KB - 1 - slot) -- one LOP3 per neighbour.  2 ALU operations per
        // neighbour and channel instead of 3; `ext` stays the exact maximum, the recorded slot is the maximum's or
        // that of a neighbour within 2^KB ulp of it (a near-tie: either choice routes the gradient legitimately).
        constexpr unsigned KM = (1u << KB) - 1u;
        float kbest[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int s = 0;
#pragma unroll 2
        for (; s + 1 < k; s += 2) {
          const int j0 = myi[s], j1 = myi[s + 1];
          const float4 a4 = reinterpret_cast<const float4*>(As + (size_t)j0 * CH)[q];
          const float4 c4v = reinterpret_cast<const float4*>(As + (size_t)j1 * CH)[q];
          const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
          const float cc[4] = {c4v.x, c4v.y, c4v.z, c4v.w};
          const unsigned code0 = KM - (unsigned)s, code1 = code0 - 1u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            best[u] = fmax3(best[u], aa[u], cc[u]);
            if (TRAIN) {
              const float k0 = __uint_as_float((__float_as_uint(aa[u]) & ~KM) | code0);
              const float k1 = __uint_as_float((__float_as_uint(cc[u]) & ~KM) | code1);
              kbest[u] = fmax3(kbest[u], k0, k1);
              s1[u] += aa[u] + cc[u];
              q1[u] = fmaf(cc[u], cc[u], fmaf(aa[u], aa[u], q1[u]));
            }
          }
        }
        if (s < k) {
          const float4 a4 = reinterpret_cast<const float4*>(As + (size_t)myi[s] * CH)[q];
          const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
          const unsigned code0 = KM - (unsigned)s;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            best[u] = fmaxf(best[u], aa[u]);
            if (TRAIN) {
              kbest[u] = fmaxf(kbest[u], __uint_as_float((__float_as_uint(aa[u]) & ~KM) | code0));
              s1[u] += aa[u];
              q1[u] = fmaf(aa[u], aa[u], q1[u]);
            }
          }
        }
        if (TRAIN) {
#pragma unroll
          for (int u = 0; u < 4; ++u) bslot[u] = (int)(KM - (__float_as_uint(kbest[u]) & KM));
        }
      }
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
      float e4[4], S4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        e4[u] = best[u] * sg[u] + bb[u];
        const float sa = s1[u] * sg[u];  // S^a_i (unsigned)
        S4[u] = fmaf(kf, bb[u], sa);
        if (TRAIN) {
          ds[u] += (double)S4[u];
          dq[u] += (double)(q1[u] + 2.f * bb[u] * sa + kf * bb[u] * bb[u]);
        }
      }
      if (TRAIN) {
        reinterpret_cast<float4*>(ext + (cb + i) * Cout + c0)[q] = make_float4(e4[0], e4[1], e4[2], e4[3]);
        reinterpret_cast<float4*>(ssum + (cb + i) * Cout + c0)[q] = make_float4(S4[0], S4[1], S4[2], S4[3]);
        reinterpret_cast<uchar4*>(arg + (cb + i) * Cout + c0)[q] =
            make_uchar4((unsigned char)bslot[0], (unsigned char)bslot[1], (unsigned char)bslot[2], (unsigned char)bslot[3]);
      } else {
        float4 o;
        o.x = act_leaky(fmaf(sc[0], e4[0], sh[0]), slope);
        o.y = act_leaky(fmaf(sc[1], e4[1], sh[1]), slope);
        o.z = act_leaky(fmaf(sc[2], e4[2], sh[2]), slope);
        o.w = act_leaky(fmaf(sc[3], e4[3], sh[3]), slope);
        *reinterpret_cast<float4*>(out + (cb + i) * ldo + c0 + 4 * q) = o;
      }
    }
    {
      int* wn = wbuf + (cur ^ 1) * (PPW * k);
      if (can_pf) {
#pragma unroll
        for (int u = 0; u < NPF; ++u)
          if (lane + 32 * u < cntn) wn[lane + 32 * u] = pf[u];
      } else {
        for (int e = lane; e < cntn; e += 32) wn[e] = __ldg(idx + (cb + pn) * k + e);
      }
      __syncwarp();
    }
  }
  if (TRAIN) {
    // thread t owns quad (t % Q) in both phases (512 % Q == 0 and 32 % Q == 0): fold the lanes of
    // a warp that share a quad with shuffles, then the 16 warps through shared memory
    double v[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = ds[u];
      v[4 + u] = dq[u];
    }
#pragma unroll
    for (int o = 16; o >= Q; o >>= 1)
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
    if (lane < Q) {
#pragma unroll
      for (int u = 0; u < 8; ++u) red[warp][lane][u] = v[u];
    }
    __syncthreads();
    if (tid < CH) {
      const int qq = tid >> 2, u = tid & 3;
      double a = 0.0, s2 = 0.0;
      for (int w = 0; w < 16; ++w) {
        a += red[w][qq][u];
        s2 += red[w][qq][4 + u];
      }
      atomicAdd(&sums[c0 + tid], a);
      atomicAdd(&sums[Cout + c0 + tid], s2);
    }
  }
}

// ---- backward, fast path -------------------------------------------------------------------------------
// Two launches instead of (pre, main) over the transposed graph with 9 B of L2 gathers per edge and channel:
//   route:  ghat = g act'(z);  G1, G2;  and the arg-routed term  Gs[j*_i[c], c] += ghat_i[c]  as ONE scatter
//           of P*Cout values (red.global.add.f32 into the dA half of `dab`, zeroed before), instead of a masked
//           gather of k*P*Cout (ghat, arg) pairs;
//   main:   per (cloud, channel chunk) CTA the chunk's `b` rows are staged in shared memory once; T_j = sum of
//           b over the reverse edges of j is gathered from there (the only per-edge work left: one LDS.128 and
//           four adds per edge and four channels), a warp works on one destination row at a time with its
//           32 / (CH/4) sub-groups taking different edges -- hub rows (in-degrees in the thousands in feature
//           space) no longer serialise on one thread -- and rows are handed out dynamically.
// The float atomics make the summation order of Gs (hence dA, dX, dW1) run-dependent at the ulp level, like
// every scatter-add backward of the reference's own PyTorch path.
__global__ void __launch_bounds__(256)
edge_bwd_route_kernel(const float* __restrict__ gout, long long ldg, const float* __restrict__ ext,
                      const uint8_t* __restrict__ arg, const int* __restrict__ idx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ mean_invstd, long long P, int N, int k,
                      int Cout, float slope, float* __restrict__ ghat, float* __restrict__ gs, long long ldgs,
                      double* __restrict__ gsums) {
  __shared__ double red[8][256];
  constexpr int R = 2;  // rows in flight per thread (the arg -> idx -> atomic chain is two dependent loads deep); the
                        // kernel is bound by the L2 atomic units (~140 G red/s measured), not by this
  const int CQ = Cout >> 2;
  const int PPB = 256 / CQ;  // rows per block step (CQ <= 256)
  const int tid = threadIdx.x;
  const int pl = tid / CQ, c4 = tid - pl * CQ;
  const bool active = pl < PPB;
  float mean[4], is[4], sc[4], sh[4];
  if (active) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = 4 * c4 + u;
      mean[u] = __ldg(mean_invstd + c);
      is[u] = __ldg(mean_invstd + Cout + c);
      sc[u] = __ldg(gamma + c) * is[u];
      sh[u] = __ldg(beta + c) - mean[u] * sc[u];
    }
  }
  float g1[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};   // fp32 over this thread's few rows, fp64 from there on
  if (active) {
    const long long stride = (long long)gridDim.x * PPB;
    for (long long ib = (long long)blockIdx.x * PPB + pl; ib < P; ib += R * stride) {
      float4 g[R], e[R];
      uchar4 a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long i = ib + r * stride;
        if (i < P) {
          g[r] = __ldg(reinterpret_cast<const float4*>(gout + i * ldg) + c4);
          e[r] = __ldg(reinterpret_cast<const float4*>(ext + i * Cout) + c4);
          a[r] = __ldg(reinterpret_cast<const uchar4*>(arg + i * Cout) + c4);
        }
      }
      int jj[R][4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long i = ib + r * stride;
        if (i < P) {
          const int* ip = idx + i * k;
          jj[r][0] = __ldg(ip + a[r].x);
          jj[r][1] = __ldg(ip + a[r].y);
          jj[r][2] = __ldg(ip + a[r].z);
          jj[r][3] = __ldg(ip + a[r].w);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long i = ib + r * stride;
        if (i >= P) break;
        const float gg[4] = {g[r].x, g[r].y, g[r].z, g[r].w};
        const float ee[4] = {e[r].x, e[r].y, e[r].z, e[r].w};
        float* grow = gs + (i / N) * N * ldgs + 4 * c4;
        float gh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float z = fmaf(sc[u], ee[u], sh[u]);
          gh[u] = gg[u] * act_leaky_grad(z, slope);
          g1[u] += gh[u];
          g2[u] = fmaf(gh[u], (ee[u] - mean[u]) * is[u], g2[u]);
          atomicAdd(grow + (long long)jj[r][u] * ldgs + u, gh[u]);
        }
        reinterpret_cast<float4*>(ghat + i * Cout)[c4] = make_float4(gh[0], gh[1], gh[2], gh[3]);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    red[v][tid] = (double)g1[v];
    red[4 + v][tid] = (double)g2[v];
  }
  __syncthreads();
  if (active && pl == 0) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double a = 0, q = 0;
      for (int r = 0; r < PPB; ++r) {
        a += red[v][r * CQ + c4];
        q += red[4 + v][r * CQ + c4];
      }
      atomicAdd(&gsums[4 * c4 + v], a);
      atomicAdd(&gsums[Cout + 4 * c4 + v], q);
    }
  }
}

template <int CH>
__global__ void __launch_bounds__(512, 1)
edge_bwd_main_smem_kernel(const float* __restrict__ ab, const float* __restrict__ ghat, const float* __restrict__ ssum,
                          const int* __restrict__ rev_ptr, const int* __restrict__ rev_edge,
                          const float* __restrict__ gamma, const float* __restrict__ mean_invstd,
                          const double* __restrict__ gsums, long long P, int N, int k, int Cout,
                          float* __restrict__ dab, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int Q = CH / 4;      // float4 lanes per row
  constexpr int NSUB = 32 / Q;   // edge sub-groups of a warp == rows per batch
  extern __shared__ __align__(16) float smem_f[];
  float* Bs = smem_f;                                                         // [N][CH]
  int* ptr_s = reinterpret_cast<int*>(smem_f + (size_t)N * CH);               // [N+1]  (padded to a multiple of 4)
  unsigned short* src_s = reinterpret_cast<unsigned short*>(ptr_s + ((N + 4) & ~3));  // [N*k] source point of every reverse edge
  __shared__ int next_row;
  const int c0 = blockIdx.x * CH;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int q = lane % Q, sub = lane / Q;
  const int ld = 2 * Cout;
  const long long cb = (long long)b * N;
  if (tid == 0) next_row = 0;
  // the whole per-edge loop runs out of shared memory: b rows of the chunk, the cloud's CSR offsets and sources
  // (eight independent loads in flight per thread: the staging is latency-, not bandwidth-bound)
  for (int e0 = tid; e0 < N * Q; e0 += 8 * 512) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * 512;
      const int n = e / Q, qq = e - n * Q;
      if (e < N * Q) v[u] = __ldg(reinterpret_cast<const float4*>(ab + (cb + n) * ld + Cout + c0) + qq);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * 512;
      const int n = e / Q, qq = e - n * Q;
      if (e < N * Q) reinterpret_cast<float4*>(Bs + (size_t)n * CH)[qq] = v[u];
    }
  }
  for (int e = tid; e <= N; e += 512) ptr_s[e] = __ldg(rev_ptr + (long long)b * (N + 1) + e);
  {
    const int* reb = rev_edge + cb * k;
    const int E = N * k;
    for (int e0 = tid; e0 < E; e0 += 8 * 512) {
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (e0 + u * 512 < E) ? __ldg(reb + e0 + u * 512) : 0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (e0 + u * 512 < E) src_s[e0 + u * 512] = (unsigned short)(v[u] >> 8);
    }
  }
  const double Md = (double)P * (double)k;
  float mean[4], sc[4], c1[4], c2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = c0 + 4 * q + u;
    mean[u] = __ldg(mean_invstd + c);
    const float is = __ldg(mean_invstd + Cout + c);
    sc[u] = __ldg(gamma + c) * is;
    const double G1 = gsums[c], G2 = gsums[Cout + c];
    c1[u] = (float)(G1 / Md);
    c2[u] = (float)(G2 / Md) * is;
    if (b == 0 && tid < Q) {
      dbeta[c] = (float)G1;
      dgamma[c] = (float)G2;
    }
  }
  __syncthreads();
  const float kf = (float)k;
  for (;;) {
    int r0 = 0;
    if (lane == 0) r0 = atomicAdd(&next_row, NSUB);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if (r0 >= N) break;
    // this lane finalises row r0 + sub: its operands are requested now and arrive while the batch is gathered
    const int jl = r0 + sub;
    const long long j = cb + min(jl, N - 1);
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(ab + j * ld + c0) + q);
    const float4 gh = __ldg(reinterpret_cast<const float4*>(ghat + j * Cout + c0) + q);
    const float4 S4 = __ldg(reinterpret_cast<const float4*>(ssum + j * Cout + c0) + q);
    float4* dap = reinterpret_cast<float4*>(dab + j * ld + c0) + q;
    const float4 G4 = *dap;  // the routed sums of the route kernel
    float T[4] = {0.f, 0.f, 0.f, 0.f};  // sub-group s ends up with the row r0 + s
    float degf = 0.f;
#pragma unroll 1
    for (int rr = 0; rr < NSUB; ++rr) {
      const int jr = r0 + rr;
      if (jr >= N) break;
      const int lo = ptr_s[jr], hi = ptr_s[jr + 1];
      float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int t = lo + sub; t < hi; t += NSUB) {
        const int i = src_s[t];
        const float4 v = reinterpret_cast<const float4*>(Bs + (size_t)i * CH)[q];
        t4[0] += v.x; t4[1] += v.y; t4[2] += v.z; t4[3] += v.w;
      }
#pragma unroll
      for (int o = Q; o < 32; o <<= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) t4[u] += __shfl_xor_sync(0xffffffffu, t4[u], o);
      }
      if (sub == rr) {
#pragma unroll
        for (int u = 0; u < 4; ++u) T[u] = t4[u];
        degf = (float)(hi - lo);
      }
    }
    if (jl < N) {
      const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
      const float gg[4] = {gh.x, gh.y, gh.z, gh.w};
      const float SS[4] = {S4.x, S4.y, S4.z, S4.w};
      const float Gs[4] = {G4.x, G4.y, G4.z, G4.w};
      float dA[4], dB[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dB[u] = sc[u] * (gg[u] - kf * c1[u] - c2[u] * (SS[u] - kf * mean[u]));
        dA[u] = sc[u] * (Gs[u] - degf * c1[u] - c2[u] * (degf * (aa[u] - mean[u]) + T[u]));
      }
      *dap = make_float4(dA[0], dA[1], dA[2], dA[3]);
      reinterpret_cast<float4*>(dab + j * ld + Cout + c0)[q] = make_float4(dB[0], dB[1], dB[2], dB[3]);
    }
  }
}

static size_t bwd_smem_bytes(int N, int CH, int k) {
  return (size_t)N * CH * 4 + (size_t)((N + 4) & ~3) * 4 + (size_t)N * k * 2 + 16;
}
// chunk width of the shared-memory backward: the widest of 32 / 16 / 8 / 4 that divides Cout and whose staged
// b rows fit; 0 = fall back to the global-memory kernels
static int bwd_chunk(int N, int Cout, int k) {
  const size_t cap = 220 * 1024;
  if (N > 65535) return 0;  // 16-bit edge sources
  for (int ch = 32; ch >= 4; ch >>= 1)
    if (Cout % ch == 0 && bwd_smem_bytes(N, ch, k) <= cap) return ch;
  return 0;
}

static int gather_grid(long long P, int Cout) {
  int ppb = 256 / (Cout >> 2);
  long long blocks = (P + ppb - 1) / ppb;
  long long cap = (long long)num_sms() * 8;
  return (int)(blocks < cap ? blocks : cap);
}

static size_t smem_bytes_gather(int N, int CH, int k) {
  return (size_t)N * CH * 4 + (size_t)2 * 16 * (128 / CH) * k * 4;  // A chunk + double-buffered neighbour lists
}
// Channel-chunk width of the shared-memory gather: 32 when that still fills the GPU, 16 for narrow
// layers, 0 (global-memory gather) when a cloud's chunk does not fit in shared memory.
static int smem_chunk(int B, int N, int Cout, int k) {
  const size_t cap = 200 * 1024;
  if (Cout % 32 == 0 && smem_bytes_gather(N, 32, k) <= cap && (long long)B * (Cout / 32) >= num_sms()) return 32;
  if (Cout % 16 == 0 && smem_bytes_gather(N, 16, k) <= cap) return 16;
  if (Cout % 32 == 0 && smem_bytes_gather(N, 32, k) <= cap) return 32;
  return 0;
}
static int smem_attr(const void* fn, size_t bytes) { return ensure_dyn_smem(fn, bytes); }

}  // namespace sug

using namespace sug;

extern "C" size_t sug_edgeconv_ws_bytes(int B, int N, int C, int Cout, int k) {
  (void)k;
  size_t P = (size_t)B * N;
  size_t w = align_up(sizeof(float) * 2 * (size_t)Cout * C, 256);
  size_t s = align_up(sizeof(double) * 2 * (size_t)Cout, 256);
  size_t gh = align_up(sizeof(float) * P * Cout, 256);
  return 2 * w + 2 * s + gh + align_up(sizeof(int) * P, 256) + 2048;
}

static int edge_check(int B, int N, int C, int Cout, int k) {
  SUG_CHECK_ARG(B > 0 && N > 0 && C > 0, "edgeconv: bad shape B=%d N=%d C=%d", B, N, C);
  SUG_CHECK_ARG(Cout % 4 == 0 && Cout >= 4 && Cout <= 1024, "edgeconv: Cout=%d must be a multiple of 4 in [4,1024]", Cout);
  SUG_CHECK_ARG(k > 0 && k <= 255, "edgeconv: k=%d out of range", k);
  return 0;
}

extern "C" int sug_edgeconv_fwd(const float* x, int64_t ldx, const int32_t* idx, const float* w, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, int B, int N, int C,
                                int Cout, int k, float eps, float momentum, float slope, int training, float* out,
                                int64_t ldo, float* ab, float* ext, uint8_t* arg, float* ssum,
                                float* save_mean_invstd, void* ws, size_t ws_bytes, sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_TRY(edge_check(B, N, C, Cout, k));
  SUG_CHECK_ARG(x && idx && w && gamma && beta && out && ab, "edgeconv_fwd: null pointer");
  SUG_CHECK_ARG(ldo % 4 == 0 && ((uintptr_t)out % 16) == 0, "edgeconv_fwd: out must be 16B aligned with ldo %% 4 == 0");
  SUG_CHECK_ARG(training || (running_mean && running_var), "edgeconv_fwd: eval mode needs the running statistics");
  if (training) SUG_CHECK_ARG(ext && arg && ssum && save_mean_invstd, "edgeconv_fwd: training needs ext/arg/ssum/save");
  const long long P = (long long)B * N;
  Workspace W(ws, ws_bytes);
  float* wcat = W.take<float>(2 * (size_t)Cout * C);
  double* sums = W.take<double>(2 * (size_t)Cout);
  float* mi_eval = W.take<float>(2 * (size_t)Cout);
  if (!W.ok()) { set_error("edgeconv_fwd: workspace too small (%zu B)", ws_bytes); return SUG_E_WORKSPACE; }

  {
    ProfScope ps(KC_MISC, 0, 0, stream);
    edge_pack_weight_kernel<<<cdiv((long long)Cout * C, 256), 256, 0, stream>>>(w, C, Cout, wcat);
  }
  SUG_LAUNCH_CHECK();
  SUG_TRY(gemm_f32(x, ldx, 1, wcat, C, 1, nullptr, ab, 2 * Cout, (int)P, 2 * Cout, C, 0, stream));
  const int grid = gather_grid(P, Cout);
  if (training) {
    SUG_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * Cout, stream));
    {
      ProfScope ps(KC_EDGE_FWD, 4.0 * P * k * Cout, (double)P * (8.0 * Cout + 4.0 * k + 9.0 * Cout), stream);
      const int CH = smem_chunk(B, N, Cout, k);
      if (CH != 0) {
        const size_t sm = smem_bytes_gather(N, CH, k);
#define SUG_GATHER_T(CH_, KB_)                                                                               \
  do {                                                                                                       \
    SUG_TRY(smem_attr((const void*)edge_gather_smem_kernel<CH_, true, KB_>, sm));                            \
    edge_gather_smem_kernel<CH_, true, KB_><<<dim3(Cout / CH_, B), 512, sm, stream>>>(                       \
        ab, idx, gamma, beta, nullptr, N, k, Cout, slope, ext, arg, ssum, sums, nullptr, 0);                 \
  } while (0)
        if (CH == 32) {
          if (k <= 32) SUG_GATHER_T(32, 5);
          else if (k <= 64) SUG_GATHER_T(32, 6);
          else SUG_GATHER_T(32, 0);
        } else {
          if (k <= 32) SUG_GATHER_T(16, 5);
          else if (k <= 64) SUG_GATHER_T(16, 6);
          else SUG_GATHER_T(16, 0);
        }
#undef SUG_GATHER_T
      } else {
        edge_gather_fwd_kernel<true><<<grid, 256, 0, stream>>>(ab, idx, gamma, beta, nullptr, (int)P, N, k, Cout, slope,
                                                               ext, arg, ssum, sums, nullptr, 0);
      }
    }
    SUG_LAUNCH_CHECK();
    SUG_TRY(bn_act_from_sums_launch(ext, gamma, beta, sums, (double)P * k, eps, momentum, running_mean, running_var,
                                    save_mean_invstd, P, Cout, slope, out, ldo, stream));
  } else {
    SUG_TRY(bn_eval_stats(running_mean, running_var, Cout, eps, mi_eval, stream));
    {
      ProfScope ps(KC_EDGE_FWD, 2.0 * P * k * Cout, (double)P * (8.0 * Cout + 4.0 * k + 4.0 * Cout), stream);
      const int CH = smem_chunk(B, N, Cout, k);
      const size_t sm = CH ? smem_bytes_gather(N, CH, k) : 0;
      if (CH == 32) {
        SUG_TRY(smem_attr((const void*)edge_gather_smem_kernel<32, false, 5>, sm));
        edge_gather_smem_kernel<32, false, 5><<<dim3(Cout / 32, B), 512, sm, stream>>>(
            ab, idx, gamma, beta, mi_eval, N, k, Cout, slope, nullptr, nullptr, nullptr, nullptr, out, ldo);
      } else if (CH == 16) {
        SUG_TRY(smem_attr((const void*)edge_gather_smem_kernel<16, false, 5>, sm));
        edge_gather_smem_kernel<16, false, 5><<<dim3(Cout / 16, B), 512, sm, stream>>>(
            ab, idx, gamma, beta, mi_eval, N, k, Cout, slope, nullptr, nullptr, nullptr, nullptr, out, ldo);
      } else {
        edge_gather_fwd_kernel<false><<<grid, 256, 0, stream>>>(ab, idx, gamma, beta, mi_eval, (int)P, N, k, Cout, slope,
                                                                nullptr, nullptr, nullptr, nullptr, out, ldo);
      }
    }
    SUG_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int sug_edgeconv_bwd(const float* gout, int64_t ldg, const float* x, int64_t ldx, const int32_t* idx,
                                const int32_t* rev_ptr, const int32_t* rev_edge, const float* w, const float* gamma,
                                const float* beta, const float* ab, const float* ext, const uint8_t* arg,
                                const float* ssum, const float* save_mean_invstd, int B, int N, int C, int Cout, int k,
                                float slope, float* dx, int64_t lddx, int accumulate_dx, float* dw, float* dgamma,
                                float* dbeta, float* dab, void* ws, size_t ws_bytes, sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_TRY(edge_check(B, N, C, Cout, k));
  SUG_CHECK_ARG(gout && x && rev_ptr && rev_edge && w && gamma && beta && ab && ext && arg && ssum &&
                    save_mean_invstd && dw && dgamma && dbeta && dab,
                "edgeconv_bwd: null pointer");
  SUG_CHECK_ARG(ldg % 4 == 0 && ((uintptr_t)gout % 16) == 0, "edgeconv_bwd: gout must be 16B aligned with ldg %% 4 == 0");
  const long long P = (long long)B * N;
  Workspace W(ws, ws_bytes);
  float* wcat = W.take<float>(2 * (size_t)Cout * C);
  float* dwcat = W.take<float>(2 * (size_t)Cout * C);
  double* gsums = W.take<double>(2 * (size_t)Cout);
  float* ghat = W.take<float>((size_t)P * Cout);
  if (!W.ok()) { set_error("edgeconv_bwd: workspace too small (%zu B)", ws_bytes); return SUG_E_WORKSPACE; }

  SUG_CUDA(cudaMemsetAsync(gsums, 0, sizeof(double) * 2 * Cout, stream));
  const int grid = gather_grid(P, Cout);
  const int CHB = (Cout <= 1024 && idx != nullptr) ? bwd_chunk(N, Cout, k) : 0;
  if (CHB != 0) {
    // routed sums accumulate in the dA half of dab
    SUG_CUDA(cudaMemset2DAsync(dab, sizeof(float) * 2 * Cout, 0, sizeof(float) * Cout, (size_t)P, stream));
    {
      ProfScope ps(KC_EDGE_BWD_PRE, 6.0 * P * Cout, (double)P * (17.0 * Cout + 4.0 * k), stream);
      edge_bwd_route_kernel<<<grid, 256, 0, stream>>>(gout, ldg, ext, arg, idx, gamma, beta, save_mean_invstd, P, N, k,
                                                      Cout, slope, ghat, dab, 2LL * Cout, gsums);
    }
    SUG_LAUNCH_CHECK();
    {
      ProfScope ps(KC_EDGE_BWD_MAIN, (double)P * k * Cout, (double)P * (28.0 * Cout + 4.0 * k), stream);
      const size_t sm = bwd_smem_bytes(N, CHB, k);
      const dim3 g2(Cout / CHB, B);
#define SUG_BWD_MAIN(CH_)                                                                                          \
  do {                                                                                                             \
    SUG_TRY(smem_attr((const void*)edge_bwd_main_smem_kernel<CH_>, sm));                                           \
    edge_bwd_main_smem_kernel<CH_><<<g2, 512, sm, stream>>>(ab, ghat, ssum, rev_ptr, rev_edge, gamma,              \
                                                            save_mean_invstd, gsums, P, N, k, Cout, dab, dgamma,   \
                                                            dbeta);                                                \
  } while (0)
      if (CHB == 32) SUG_BWD_MAIN(32);
      else if (CHB == 16) SUG_BWD_MAIN(16);
      else if (CHB == 8) SUG_BWD_MAIN(8);
      else SUG_BWD_MAIN(4);
#undef SUG_BWD_MAIN
    }
    SUG_LAUNCH_CHECK();
  } else {
    {
      ProfScope ps(KC_EDGE_BWD_PRE, 6.0 * P * Cout, 12.0 * P * Cout, stream);
      edge_bwd_pre_kernel<<<grid, 256, 0, stream>>>(gout, ldg, ext, gamma, beta, save_mean_invstd, P, Cout, slope, ghat,
                                                    gsums);
    }
    SUG_LAUNCH_CHECK();
    {
      ProfScope ps(KC_EDGE_BWD_MAIN, 3.0 * P * k * Cout, (double)P * (8.0 * Cout + 9.0 * Cout + 4.0 * k + 8.0 * Cout), stream);
      edge_bwd_main_kernel<<<grid, 256, 0, stream>>>(ab, ghat, arg, ssum, rev_ptr, rev_edge, gamma, save_mean_invstd,
                                                     gsums, P, N, k, Cout, dab, dgamma, dbeta);
    }
    SUG_LAUNCH_CHECK();
  }
  // dWcat = dab^T x   ([2Cout, P] x [P, C])
  SUG_TRY(gemm_f32(dab, 1, 2 * Cout, x, 1, ldx, nullptr, dwcat, C, 2 * Cout, C, (int)P, 0, stream));
  {
    ProfScope ps(KC_MISC, 0, 0, stream);
    edge_unpack_wgrad_kernel<<<cdiv((long long)Cout * C, 256), 256, 0, stream>>>(dwcat, C, Cout, dw);
  }
  SUG_LAUNCH_CHECK();
  if (dx != nullptr) {
    {
      ProfScope ps(KC_MISC, 0, 0, stream);
      edge_pack_weight_kernel<<<cdiv((long long)Cout * C, 256), 256, 0, stream>>>(w, C, Cout, wcat);
    }
    SUG_LAUNCH_CHECK();
    // dx = dab * Wcat   ([P, 2Cout] x [2Cout, C])
    SUG_TRY(gemm_f32(dab, 2 * Cout, 1, wcat, 1, C, nullptr, dx, lddx, (int)P, C, 2 * Cout, accumulate_dx, stream));
  }
  return 0;
}
