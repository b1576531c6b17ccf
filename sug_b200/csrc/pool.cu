// Shared MLP (1x1 conv) + BatchNorm + activation + global pool over the points of a cloud.
//   DGCNN tail  : Conv1d(512,512) -> BatchNorm1d -> leaky_relu(0.2) -> max || avg   (Model.py:111-116)
//   PointNet    : conv_2d(128,1024) -> BN -> ReLU -> max over N                      (Model.py:245,272-274)
// The [B*N, Cout] linear output y is written once by the GEMM; the activated tensor never exists:
// the pool kernel reads y, applies the monotone BN-affine + activation on the fly and keeps, per
// (cloud, channel), the extreme value (max, or min where gamma < 0), its point index and the sum
// of activations.  The backward rebuilds dL/dz from (gmax, gavg, argext), reduces the two
// BatchNorm sums and overwrites y with dL/dy in place.
#include "common.cuh"

namespace sug {

constexpr int PCQ = 32;  // channel quads per block  (128 channels, 512 B per row segment)
constexpr int PRL = 8;   // row lanes per block

// per-column sum / sum of squares of y [P, Cout]
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ y, long long P, int Cout, double* __restrict__ sums) {
  __shared__ double red[256][8];
  const int tid = threadIdx.x;
  const int c4l = tid % PCQ, rl = tid / PCQ;
  const int c4 = blockIdx.x * PCQ + c4l;
  const bool active = 4 * c4 < Cout;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (active) {
    // fp32 partials over short runs, flushed to fp64
    for (long long r0 = (long long)blockIdx.y * PRL + rl; r0 < P; r0 += (long long)gridDim.y * PRL * 16) {
      float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
#pragma unroll 4
      for (int t = 0; t < 16; ++t) {
        long long r = r0 + (long long)t * gridDim.y * PRL;
        if (r >= P) break;
        float4 v = __ldg(reinterpret_cast<const float4*>(y + r * Cout) + c4);
        fs[0] += v.x; fs[1] += v.y; fs[2] += v.z; fs[3] += v.w;
        fq[0] = fmaf(v.x, v.x, fq[0]); fq[1] = fmaf(v.y, v.y, fq[1]);
        fq[2] = fmaf(v.z, v.z, fq[2]); fq[3] = fmaf(v.w, v.w, fq[3]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s[u] += (double)fs[u]; q[u] += (double)fq[u]; }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) { red[tid][u] = s[u]; red[tid][4 + u] = q[u]; }
  __syncthreads();
  if (active && rl == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double a = 0, b = 0;
      for (int r = 0; r < PRL; ++r) { a += red[r * PCQ + c4l][u]; b += red[r * PCQ + c4l][4 + u]; }
      atomicAdd(&sums[4 * c4 + u], a);
      atomicAdd(&sums[Cout + 4 * c4 + u], b);
    }
  }
}

// grid (Cout/128, B, nsplit).  out[b, c] = act(scale*ext+shift); POOL_MAX_AVG: out[b, Cout+c] = mean_n act(.)
// nsplit > 1 (few clouds, many points: LiDAR-scale inference): every z-slice reduces its share of the N points and
// writes (signed extreme, arg, sum) partials; pool_combine_kernel finishes.
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ mean_invstd, int N, int Cout, float slope, int pool,
                float* __restrict__ out, int* __restrict__ argext, float* __restrict__ part_ext,
                int* __restrict__ part_arg, float* __restrict__ part_sum) {
  __shared__ float s_ext[PRL][PCQ * 4];
  __shared__ int s_arg[PRL][PCQ * 4];
  __shared__ float s_sum[PRL][PCQ * 4];
  const int tid = threadIdx.x;
  const int c4l = tid % PCQ, rl = tid / PCQ;
  const int c4 = blockIdx.x * PCQ + c4l;
  const int b = blockIdx.y;
  const int nsplit = gridDim.z, zs = blockIdx.z;
  const int per = (N + nsplit - 1) / nsplit;
  const int n_lo = zs * per, n_hi = min(N, n_lo + per);
  const bool active = 4 * c4 < Cout;
  float sc[4], sh[4], sg[4];
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int barg[4] = {0, 0, 0, 0};
  float asum[4] = {0, 0, 0, 0};
  if (active) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int c = 4 * c4 + u;
      float g = __ldg(gamma + c);
      sc[u] = g * __ldg(mean_invstd + Cout + c);
      sh[u] = __ldg(beta + c) - __ldg(mean_invstd + c) * sc[u];
      sg[u] = g < 0.f ? -1.f : 1.f;
    }
    const float* yb = y + (long long)b * N * Cout;
    for (int n = n_lo + rl; n < n_hi; n += PRL) {
      float4 v = __ldg(reinterpret_cast<const float4*>(yb + (long long)n * Cout) + c4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float t = vv[u] * sg[u];
        if (t > best[u]) { best[u] = t; barg[u] = n; }
        asum[u] += act_leaky(fmaf(sc[u], vv[u], sh[u]), slope);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    s_ext[rl][4 * c4l + u] = best[u];
    s_arg[rl][4 * c4l + u] = barg[u];
    s_sum[rl][4 * c4l + u] = asum[u];
  }
  __syncthreads();
  if (tid < PCQ * 4) {
    const int c = blockIdx.x * PCQ * 4 + tid;
    if (c < Cout) {
      float bv = s_ext[0][tid];
      int ba = s_arg[0][tid];
      float sm = s_sum[0][tid];
      for (int r = 1; r < PRL; ++r) {
        float v = s_ext[r][tid];
        int a = s_arg[r][tid];
        if (v > bv || (v == bv && a < ba)) { bv = v; ba = a; }
        sm += s_sum[r][tid];
      }
      if (nsplit > 1) {
        const long long o = ((long long)b * nsplit + zs) * Cout + c;
        part_ext[o] = bv;
        part_arg[o] = ba;
        part_sum[o] = sm;
        return;
      }
      float g = __ldg(gamma + c);
      float scl = g * __ldg(mean_invstd + Cout + c);
      float shf = __ldg(beta + c) - __ldg(mean_invstd + c) * scl;
      float e = g < 0.f ? -bv : bv;
      const int OW = pool == SUG_POOL_MAX_AVG ? 2 * Cout : Cout;
      out[(long long)b * OW + c] = act_leaky(fmaf(scl, e, shf), slope);
      if (pool == SUG_POOL_MAX_AVG) out[(long long)b * OW + Cout + c] = sm / (float)N;
      if (argext != nullptr) argext[(long long)b * Cout + c] = ba;
    }
  }
}

__global__ void pool_combine_kernel(const float* __restrict__ part_ext, const int* __restrict__ part_arg,
                                    const float* __restrict__ part_sum, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean_invstd, int B, int N,
                                    int Cout, int nsplit, float slope, int pool, float* __restrict__ out,
                                    int* __restrict__ argext) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * Cout) return;
  const int b = e / Cout, c = e - b * Cout;
  float bv = -INFINITY, sm = 0.f;
  int ba = 0;
  for (int z = 0; z < nsplit; ++z) {  // slices are in ascending point order: ties keep the lower index
    const long long o = ((long long)b * nsplit + z) * Cout + c;
    const float v = part_ext[o];
    if (v > bv) { bv = v; ba = part_arg[o]; }
    sm += part_sum[o];
  }
  const float g = __ldg(gamma + c);
  const float scl = g * __ldg(mean_invstd + Cout + c);
  const float shf = __ldg(beta + c) - __ldg(mean_invstd + c) * scl;
  const float ev = g < 0.f ? -bv : bv;
  const int OW = pool == SUG_POOL_MAX_AVG ? 2 * Cout : Cout;
  out[(long long)b * OW + c] = act_leaky(fmaf(scl, ev, shf), slope);
  if (pool == SUG_POOL_MAX_AVG) out[(long long)b * OW + Cout + c] = sm / (float)N;
  if (argext != nullptr) argext[(long long)b * Cout + c] = ba;
}

// dz_nc = act'(z_nc) * (gavg_bc / N + [n == argext_bc] gmax_bc)
__device__ __forceinline__ float pool_dz(float yv, float sc, float sh, float slope, float gavg_n, float gmax, bool is_arg) {
  float z = fmaf(sc, yv, sh);
  float d = act_leaky_grad(z, slope);
  return d * (gavg_n + (is_arg ? gmax : 0.f));
}

// grid (Cout/128, B): G1 += dz, G2 += dz * yhat
__global__ void __launch_bounds__(256)
pool_bwd_pre_kernel(const float* __restrict__ y, const float* __restrict__ gout, const int* __restrict__ argext,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ mean_invstd, int N, int Cout, float slope, int pool,
                    double* __restrict__ gsums) {
  __shared__ double red[256][8];
  const int tid = threadIdx.x;
  const int c4l = tid % PCQ, rl = tid / PCQ;
  const int c4 = blockIdx.x * PCQ + c4l;
  const int b = blockIdx.y;
  const bool active = 4 * c4 < Cout;
  double g1[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};
  if (active) {
    const int OW = pool == SUG_POOL_MAX_AVG ? 2 * Cout : Cout;
    float sc[4], sh[4], mean[4], is[4], gm[4], ga[4];
    int ar[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int c = 4 * c4 + u;
      mean[u] = __ldg(mean_invstd + c);
      is[u] = __ldg(mean_invstd + Cout + c);
      sc[u] = __ldg(gamma + c) * is[u];
      sh[u] = __ldg(beta + c) - mean[u] * sc[u];
      gm[u] = __ldg(gout + (long long)b * OW + c);
      ga[u] = pool == SUG_POOL_MAX_AVG ? __ldg(gout + (long long)b * OW + Cout + c) / (float)N : 0.f;
      ar[u] = __ldg(argext + (long long)b * Cout + c);
    }
    const float* yb = y + (long long)b * N * Cout;
    float f1[4] = {0, 0, 0, 0}, f2[4] = {0, 0, 0, 0};
    int run = 0;
    for (int n = rl; n < N; n += PRL) {
      float4 v = __ldg(reinterpret_cast<const float4*>(yb + (long long)n * Cout) + c4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float dz = pool_dz(vv[u], sc[u], sh[u], slope, ga[u], gm[u], n == ar[u]);
        f1[u] += dz;
        f2[u] = fmaf(dz, (vv[u] - mean[u]) * is[u], f2[u]);
      }
      if (++run == 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { g1[u] += (double)f1[u]; g2[u] += (double)f2[u]; f1[u] = 0.f; f2[u] = 0.f; }
        run = 0;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { g1[u] += (double)f1[u]; g2[u] += (double)f2[u]; }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) { red[tid][u] = g1[u]; red[tid][4 + u] = g2[u]; }
  __syncthreads();
  if (active && rl == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double a = 0, q = 0;
      for (int r = 0; r < PRL; ++r) { a += red[r * PCQ + c4l][u]; q += red[r * PCQ + c4l][4 + u]; }
      atomicAdd(&gsums[4 * c4 + u], a);
      atomicAdd(&gsums[Cout + 4 * c4 + u], q);
    }
  }
}

// y <- dL/dy = scale (dz - G1/M - yhat G2/M), in place.  grid (Cout/128, B)
__global__ void __launch_bounds__(256)
pool_bwd_main_kernel(float* __restrict__ y, const float* __restrict__ gout, const int* __restrict__ argext,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ mean_invstd, const double* __restrict__ gsums, int B, int N, int Cout,
                     float slope, int pool, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int tid = threadIdx.x;
  const int c4l = tid % PCQ, rl = tid / PCQ;
  const int c4 = blockIdx.x * PCQ + c4l;
  const int b = blockIdx.y;
  if (4 * c4 >= Cout) return;
  const int OW = pool == SUG_POOL_MAX_AVG ? 2 * Cout : Cout;
  const double Md = (double)B * (double)N;
  float sc[4], sh[4], mean[4], is[4], gm[4], ga[4], c1[4], c2[4];
  int ar[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int c = 4 * c4 + u;
    mean[u] = __ldg(mean_invstd + c);
    is[u] = __ldg(mean_invstd + Cout + c);
    sc[u] = __ldg(gamma + c) * is[u];
    sh[u] = __ldg(beta + c) - mean[u] * sc[u];
    gm[u] = __ldg(gout + (long long)b * OW + c);
    ga[u] = pool == SUG_POOL_MAX_AVG ? __ldg(gout + (long long)b * OW + Cout + c) / (float)N : 0.f;
    ar[u] = __ldg(argext + (long long)b * Cout + c);
    double G1 = gsums[c], G2 = gsums[Cout + c];
    c1[u] = (float)(G1 / Md);
    c2[u] = (float)(G2 / Md);
    if (b == 0 && rl == 0) { dbeta[c] = (float)G1; dgamma[c] = (float)G2; }
  }
  float* yb = y + (long long)b * N * Cout;
  for (int n = rl; n < N; n += PRL) {
    float4* p = reinterpret_cast<float4*>(yb + (long long)n * Cout) + c4;
    float4 v = *p;
    const float vv[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float dz = pool_dz(vv[u], sc[u], sh[u], slope, ga[u], gm[u], n == ar[u]);
      float yh = (vv[u] - mean[u]) * is[u];
      o[u] = sc[u] * (dz - c1[u] - yh * c2[u]);
    }
    *p = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---- pointwise (no pooling) BatchNorm + activation backward ----------------------------------------
// dz = g * act'(z);  G1 += dz;  G2 += dz * yhat      (y = linear output, z = scale*y + shift)
__global__ void __launch_bounds__(256)
pw_bwd_pre_kernel(const float* __restrict__ y, const float* __restrict__ gout, long long ldg,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ mean_invstd, long long P, int Cout, float slope,
                  double* __restrict__ gsums) {
  __shared__ double red[256][8];
  const int tid = threadIdx.x;
  const int c4l = tid % PCQ, rl = tid / PCQ;
  const int c4 = blockIdx.x * PCQ + c4l;
  const bool active = 4 * c4 < Cout;
  double g1[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};
  if (active) {
    float sc[4], sh[4], mean[4], is[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int c = 4 * c4 + u;
      mean[u] = __ldg(mean_invstd + c);
      is[u] = __ldg(mean_invstd + Cout + c);
      sc[u] = __ldg(gamma + c) * is[u];
      sh[u] = __ldg(beta + c) - mean[u] * sc[u];
    }
    float f1[4] = {0, 0, 0, 0}, f2[4] = {0, 0, 0, 0};
    int run = 0;
    for (long long r = (long long)blockIdx.y * PRL + rl; r < P; r += (long long)gridDim.y * PRL) {
      float4 v = __ldg(reinterpret_cast<const float4*>(y + r * Cout) + c4);
      float4 g = __ldg(reinterpret_cast<const float4*>(gout + r * ldg) + c4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float dz = gg[u] * act_leaky_grad(fmaf(sc[u], vv[u], sh[u]), slope);
        f1[u] += dz;
        f2[u] = fmaf(dz, (vv[u] - mean[u]) * is[u], f2[u]);
      }
      if (++run == 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { g1[u] += (double)f1[u]; g2[u] += (double)f2[u]; f1[u] = 0.f; f2[u] = 0.f; }
        run = 0;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { g1[u] += (double)f1[u]; g2[u] += (double)f2[u]; }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) { red[tid][u] = g1[u]; red[tid][4 + u] = g2[u]; }
  __syncthreads();
  if (active && rl == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double a = 0, q = 0;
      for (int r = 0; r < PRL; ++r) { a += red[r * PCQ + c4l][u]; q += red[r * PCQ + c4l][4 + u]; }
      atomicAdd(&gsums[4 * c4 + u], a);
      atomicAdd(&gsums[Cout + 4 * c4 + u], q);
    }
  }
}

// y <- dL/dy = scale (dz - G1/M - yhat G2/M), in place
__global__ void __launch_bounds__(256)
pw_bwd_main_kernel(float* __restrict__ y, const float* __restrict__ gout, long long ldg,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ mean_invstd, const double* __restrict__ gsums, long long P, int Cout,
                   float slope, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int tid = threadIdx.x;
  const int c4l = tid % PCQ, rl = tid / PCQ;
  const int c4 = blockIdx.x * PCQ + c4l;
  if (4 * c4 >= Cout) return;
  const double Md = (double)P;
  float sc[4], sh[4], mean[4], is[4], c1[4], c2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int c = 4 * c4 + u;
    mean[u] = __ldg(mean_invstd + c);
    is[u] = __ldg(mean_invstd + Cout + c);
    sc[u] = __ldg(gamma + c) * is[u];
    sh[u] = __ldg(beta + c) - mean[u] * sc[u];
    double G1 = gsums[c], G2 = gsums[Cout + c];
    c1[u] = (float)(G1 / Md);
    c2[u] = (float)(G2 / Md);
    if (blockIdx.y == 0 && rl == 0) { dbeta[c] = (float)G1; dgamma[c] = (float)G2; }
  }
  for (long long r = (long long)blockIdx.y * PRL + rl; r < P; r += (long long)gridDim.y * PRL) {
    float4* p = reinterpret_cast<float4*>(y + r * Cout) + c4;
    float4 v = *p;
    float4 g = __ldg(reinterpret_cast<const float4*>(gout + r * ldg) + c4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const float gg[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float dz = gg[u] * act_leaky_grad(fmaf(sc[u], vv[u], sh[u]), slope);
      o[u] = sc[u] * (dz - c1[u] - (vv[u] - mean[u]) * is[u] * c2[u]);
    }
    *p = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace sug

using namespace sug;

// ---- linear (1x1 conv over points) + BatchNorm + activation --------------------------------------
extern "C" int sug_linear_bn_act_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, int64_t P, int Cin,
                                     int Cout, float eps, float momentum, float slope, int training, float* y,
                                     float* out, int64_t ldo, float* save_mean_invstd, void* ws, size_t ws_bytes,
                                     sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_CHECK_ARG(x && w && gamma && beta && y && out, "linear_bn_act_fwd: null pointer");
  SUG_CHECK_ARG(training || (running_mean && running_var), "linear_bn_act_fwd: eval mode needs the running statistics");
  SUG_CHECK_ARG(P > 0 && Cin > 0 && Cout > 0 && Cout % 4 == 0 && ldo % 4 == 0, "linear_bn_act_fwd: bad shape");
  Workspace W(ws, ws_bytes);
  double* sums = W.take<double>(2 * (size_t)Cout);
  float* mi_eval = W.take<float>(2 * (size_t)Cout);
  if (!W.ok()) { set_error("linear_bn_act_fwd: workspace too small"); return SUG_E_WORKSPACE; }
  const int gx = cdiv(Cout, PCQ * 4);
  const float* mi = save_mean_invstd;
  if (training) {
    SUG_CHECK_ARG(save_mean_invstd != nullptr, "linear_bn_act_fwd: training needs save_mean_invstd");
    SUG_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * Cout, stream));
    bool fused = false;  // BatchNorm sums from the GEMM epilogue when the tensor-core kernel runs
    SUG_TRY(gemm_f32_colstats(x, ldx, w, Cin, bias, y, Cout, (int)P, Cout, Cin, sums, &fused, stream));
    if (!fused) {
      int gy = (int)min((long long)num_sms() * 4 / gx + 1, (long long)(P + PRL - 1) / PRL);
      ProfScope ps(KC_COLSTATS, 3.0 * P * Cout, 4.0 * P * Cout, stream);
      col_stats_kernel<<<dim3(gx, gy), 256, 0, stream>>>(y, P, Cout, sums);
    }
    SUG_LAUNCH_CHECK();
    SUG_TRY(bn_act_from_sums_launch(y, gamma, beta, sums, (double)P, eps, momentum, running_mean, running_var,
                                    save_mean_invstd, P, Cout, slope, out, ldo, stream));
    return 0;
  }
  SUG_TRY(gemm_f32(x, ldx, 1, w, Cin, 1, bias, y, Cout, (int)P, Cout, Cin, 0, stream));
  SUG_TRY(bn_eval_stats(running_mean, running_var, Cout, eps, mi_eval, stream));
  mi = mi_eval;
  SUG_TRY(bn_act_launch(y, gamma, beta, mi, P, Cout, slope, out, ldo, stream));
  return 0;
}

extern "C" int sug_linear_bn_act_bwd(const float* gout, int64_t ldg, const float* x, int64_t ldx, const float* w,
                                     const float* gamma, const float* beta, float* y, const float* save_mean_invstd,
                                     int64_t P, int Cin, int Cout, float slope, float* dx, int64_t lddx, float* dw,
                                     float* dbias, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                                     sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_CHECK_ARG(gout && x && w && gamma && beta && y && save_mean_invstd && dw && dgamma && dbeta,
                "linear_bn_act_bwd: null pointer");
  SUG_CHECK_ARG(P > 0 && Cin > 0 && Cout > 0 && Cout % 4 == 0 && ldg % 4 == 0, "linear_bn_act_bwd: bad shape");
  Workspace W(ws, ws_bytes);
  double* gsums = W.take<double>(2 * (size_t)Cout);
  if (!W.ok()) { set_error("linear_bn_act_bwd: workspace too small"); return SUG_E_WORKSPACE; }
  SUG_CUDA(cudaMemsetAsync(gsums, 0, sizeof(double) * 2 * Cout, stream));
  const int gx = cdiv(Cout, PCQ * 4);
  const int gy = (int)min((long long)num_sms() * 4 / gx + 1, (long long)(P + PRL - 1) / PRL);
  {
    ProfScope ps(KC_POOL_BWD, 8.0 * P * Cout, 8.0 * P * Cout, stream);
    pw_bwd_pre_kernel<<<dim3(gx, gy), 256, 0, stream>>>(y, gout, ldg, gamma, beta, save_mean_invstd, P, Cout, slope, gsums);
  }
  SUG_LAUNCH_CHECK();
  {
    ProfScope ps(KC_POOL_BWD, 10.0 * P * Cout, 12.0 * P * Cout, stream);
    pw_bwd_main_kernel<<<dim3(gx, gy), 256, 0, stream>>>(y, gout, ldg, gamma, beta, save_mean_invstd, gsums, P, Cout,
                                                         slope, dgamma, dbeta);
  }
  SUG_LAUNCH_CHECK();
  if (dbias != nullptr) SUG_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, stream));
  SUG_TRY(gemm_f32(y, 1, Cout, x, 1, ldx, nullptr, dw, Cin, Cout, Cin, (int)P, 0, stream));
  if (dx != nullptr) SUG_TRY(gemm_f32(y, Cout, 1, w, 1, Cin, nullptr, dx, lddx, (int)P, Cin, Cout, 0, stream));
  return 0;
}

extern "C" size_t sug_mlp_pool_ws_bytes(int B, int N, int Cin, int Cout) {
  (void)Cin;
  // statistics + (few clouds, many points: the reduction over N is split) partial (extreme, arg, sum) slices
  const long long gx = cdiv(Cout, PCQ * 4), blocks = gx * B, sms2 = 2LL * num_sms();
  size_t parts = 0;
  if (blocks < sms2) {
    const long long nsplit = min((sms2 + blocks - 1) / blocks, (long long)(N + 255) / 256);
    if (nsplit > 1) parts = 3 * align_up(sizeof(float) * (size_t)B * Cout * (size_t)nsplit, 256);
  }
  return align_up(sizeof(double) * 2 * (size_t)Cout, 256) + align_up(sizeof(float) * 2 * (size_t)Cout, 256) + parts + 1024;
}

extern "C" int sug_mlp_pool_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, int B, int N, int Cin,
                                int Cout, float eps, float momentum, float slope, int pool, int training, float* y,
                                float* out, int32_t* argext, float* save_mean_invstd, void* ws, size_t ws_bytes,
                                sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SUG_CHECK_ARG(x && w && gamma && beta && y && out, "mlp_pool_fwd: null pointer");
  SUG_CHECK_ARG(training || (running_mean && running_var), "mlp_pool_fwd: eval mode needs the running statistics");
  SUG_CHECK_ARG(B > 0 && N > 0 && Cin > 0 && Cout > 0 && Cout % 4 == 0, "mlp_pool_fwd: bad shape");
  SUG_CHECK_ARG(pool == SUG_POOL_MAX || pool == SUG_POOL_MAX_AVG, "mlp_pool_fwd: bad pool mode %d", pool);
  if (training) SUG_CHECK_ARG(argext && save_mean_invstd, "mlp_pool_fwd: training needs argext/save");
  const long long P = (long long)B * N;
  Workspace W(ws, ws_bytes);
  double* sums = W.take<double>(2 * (size_t)Cout);
  float* mi_eval = W.take<float>(2 * (size_t)Cout);
  if (!W.ok()) { set_error("mlp_pool_fwd: workspace too small"); return SUG_E_WORKSPACE; }
  const int gx = cdiv(Cout, PCQ * 4);
  const float* mi = save_mean_invstd;
  if (training) {
    SUG_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * Cout, stream));
    bool fused = false;  // BatchNorm sums from the GEMM epilogue when the tensor-core kernel runs
    SUG_TRY(gemm_f32_colstats(x, ldx, w, Cin, bias, y, Cout, (int)P, Cout, Cin, sums, &fused, stream));
    if (!fused) {
      int gy = (int)min((long long)num_sms() * 4 / gx + 1, (P + PRL - 1) / PRL);
      ProfScope ps(KC_COLSTATS, 3.0 * P * Cout, 4.0 * P * Cout, stream);
      col_stats_kernel<<<dim3(gx, gy), 256, 0, stream>>>(y, P, Cout, sums);
    }
    SUG_LAUNCH_CHECK();
    SUG_TRY(bn_finalize_stats(sums, Cout, (double)P, eps, momentum, running_mean, running_var, save_mean_invstd,
                              stream));
  } else {
    SUG_TRY(gemm_f32(x, ldx, 1, w, Cin, 1, bias, y, Cout, (int)P, Cout, Cin, 0, stream));
    SUG_TRY(bn_eval_stats(running_mean, running_var, Cout, eps, mi_eval, stream));
    mi = mi_eval;
  }
  // few clouds with many points: split N over grid.z so that the reduction still fills the GPU
  int nsplit = 1;
  if ((long long)gx * B < 2LL * num_sms()) {
    nsplit = (int)min((long long)(2 * num_sms() + gx * B - 1) / ((long long)gx * B), (long long)(N + 255) / 256);
    if (nsplit < 1) nsplit = 1;
  }
  float* part_ext = nullptr;
  int* part_arg = nullptr;
  float* part_sum = nullptr;
  if (nsplit > 1) {
    part_ext = W.take<float>((size_t)B * nsplit * Cout);
    part_arg = W.take<int>((size_t)B * nsplit * Cout);
    part_sum = W.take<float>((size_t)B * nsplit * Cout);
    if (!W.ok()) nsplit = 1;  // workspace of an older caller: fall back to the unsplit reduction
  }
  {
    ProfScope ps(KC_POOL_FWD, 5.0 * P * Cout, 4.0 * P * Cout, stream);
    pool_fwd_kernel<<<dim3(gx, B, nsplit), 256, 0, stream>>>(y, gamma, beta, mi, N, Cout, slope, pool, out, argext, part_ext,
                                                             part_arg, part_sum);
    if (nsplit > 1)
      pool_combine_kernel<<<cdiv((long long)B * Cout, 256), 256, 0, stream>>>(part_ext, part_arg, part_sum, gamma, beta, mi, B, N,
                                                                             Cout, nsplit, slope, pool, out, argext);
  }
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_mlp_pool_bwd(const float* gout, const float* x, int64_t ldx, const float* w, const float* bias,
                                const float* gamma, const float* beta, float* y, const int32_t* argext,
                                const float* save_mean_invstd, int B, int N, int Cin, int Cout, float slope, int pool,
                                float* dx, int64_t lddx, int accumulate_dx, float* dw, float* dbias, float* dgamma,
                                float* dbeta, void* ws, size_t ws_bytes, sug_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  (void)bias;
  SUG_CHECK_ARG(gout && x && w && gamma && beta && y && argext && save_mean_invstd && dw && dgamma && dbeta,
                "mlp_pool_bwd: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && Cin > 0 && Cout > 0 && Cout % 4 == 0, "mlp_pool_bwd: bad shape");
  const long long P = (long long)B * N;
  Workspace W(ws, ws_bytes);
  double* gsums = W.take<double>(2 * (size_t)Cout);
  if (!W.ok()) { set_error("mlp_pool_bwd: workspace too small"); return SUG_E_WORKSPACE; }
  SUG_CUDA(cudaMemsetAsync(gsums, 0, sizeof(double) * 2 * Cout, stream));
  const int gx = cdiv(Cout, PCQ * 4);
  {
    ProfScope ps(KC_POOL_BWD, 8.0 * P * Cout, 4.0 * P * Cout, stream);
    pool_bwd_pre_kernel<<<dim3(gx, B), 256, 0, stream>>>(y, gout, argext, gamma, beta, save_mean_invstd, N, Cout,
                                                         slope, pool, gsums);
  }
  SUG_LAUNCH_CHECK();
  {
    ProfScope ps(KC_POOL_BWD, 10.0 * P * Cout, 8.0 * P * Cout, stream);
    pool_bwd_main_kernel<<<dim3(gx, B), 256, 0, stream>>>(y, gout, argext, gamma, beta, save_mean_invstd, gsums, B, N,
                                                          Cout, slope, pool, dgamma, dbeta);
  }
  SUG_LAUNCH_CHECK();
  // a per-channel constant added before train-mode BatchNorm has zero gradient
  if (dbias != nullptr) SUG_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, stream));
  // dw = dy^T x   ([Cout, P] x [P, Cin])
  SUG_TRY(gemm_f32(y, 1, Cout, x, 1, ldx, nullptr, dw, Cin, Cout, Cin, (int)P, 0, stream));
  if (dx != nullptr)
    SUG_TRY(gemm_f32(y, Cout, 1, w, 1, Cin, nullptr, dx, lddx, (int)P, Cin, Cout, accumulate_dx, stream));
  return 0;
}
