// Exact-fp32 CUDA-core GEMM: C[M,N] (+)= A * B^T (+ bias).
//
// Used for the xyz layer (K = 3 / 6, where tensor cores have nothing to do), for the narrow
// weight-gradient reductions, and as the in-library fp32 reference of the tcgen05 path.
// 128x128x8 tiles, 256 threads, 8x8 outputs per thread (split 4+4 so that every LDS.128 of a
// quarter-warp is contiguous), double-buffered shared memory with register prefetch.
// Operands are addressed with generic (row, k) strides so the same kernel serves
//   X * W^T            (both k-contiguous)            forward per-point GEMM
//   dY * W             (A k-contiguous, B n-contiguous) input gradient
//   dY^T * X           (both k-strided, split-K)        weight gradient
#include "common.cuh"

namespace sug {

constexpr int GBM = 128, GBN = 128, GBK = 8, GPAD = 4;

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2)
gemm_simt_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm,
                 long long sbn, long long sbk, const float* __restrict__ bias, float* __restrict__ C,
                 long long ldc, int M, int N, int K, int kchunk, int accumulate, int use_atomic) {
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (A_KC) { kk = e % GBK; mm = e / GBK; } else { mm = e % GBM; kk = e / GBM; }
      int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < M && gk < kend) ? __ldg(A + (long long)gm * sam + (long long)gk * sak) : 0.f;
      int nn, kb;
      if (B_KC) { kb = e % GBK; nn = e / GBK; } else { nn = e % GBN; kb = e / GBN; }
      int gn = n0 + nn, gkb = k0 + kb;
      rb[i] = (gn < N && gkb < kend) ? __ldg(Bm + (long long)gn * sbn + (long long)gkb * sbk) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (A_KC) { kk = e % GBK; mm = e / GBK; } else { mm = e % GBM; kk = e / GBM; }
      As[buf][kk][mm] = ra[i];
      int nn, kb;
      if (B_KC) { kb = e % GBK; nn = e / GBK; } else { nn = e % GBN; kb = e / GBN; }
      Bs[buf][kb][nn] = rb[i];
    }
  };

  if (kbeg < kend) {
    load_tile(kbeg);
    store_tile(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += GBK, buf ^= 1) {
    const bool has_next = k0 + GBK < kend;
    if (has_next) load_tile(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) store_tile(buf ^ 1);
    __syncthreads();
  }

  const bool add_bias = bias != nullptr && blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (add_bias) v += __ldg(bias + gn);
      float* p = C + (long long)gm * ldc + gn;
      if (use_atomic) atomicAdd(p, v);
      else if (accumulate) *p += v;
      else *p = v;
    }
  }
}

int gemm_simt_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                  const float* bias, float* c, int64_t ldc, int M, int N, int K, int accumulate,
                  cudaStream_t stream) {
  SUG_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  SUG_CHECK_ARG(a && b && c, "gemm: null operand");
  const int tm = cdiv(M, GBM), tn = cdiv(N, GBN);
  const long long tiles = (long long)tm * tn;
  // few output tiles and a long reduction (the 128 x 128 Gram matrix of the MMD over 512 ... 4106 features, narrow
  // weight gradients): split K until about two CTAs per SM are in flight, 64 k per CTA at least
  int splits = 1;
  if (tiles < num_sms() && K >= 256) {
    splits = (int)min((long long)cdiv(K, 64), (2LL * num_sms() + tiles - 1) / tiles);
    if (splits < 1) splits = 1;
  }
  int kchunk = cdiv(cdiv(K, splits), GBK) * GBK;
  splits = cdiv(K, kchunk);
  const int use_atomic = splits > 1;
  if (use_atomic && !accumulate)
    SUG_CUDA(cudaMemset2DAsync(c, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream));
  dim3 grid(tn, tm, splits);
  ProfScope ps(KC_GEMM, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
  const bool akc = (sak == 1), bkc = (sbk == 1);
#define SUG_GEMM_LAUNCH(AK, BK_)                                                                             \
  gemm_simt_kernel<AK, BK_><<<grid, 256, 0, stream>>>(a, sam, sak, b, sbn, sbk, bias, c, ldc, M, N, K, kchunk, \
                                                      accumulate, use_atomic)
  if (akc && bkc) SUG_GEMM_LAUNCH(true, true);
  else if (akc) SUG_GEMM_LAUNCH(true, false);
  else if (bkc) SUG_GEMM_LAUNCH(false, true);
  else SUG_GEMM_LAUNCH(false, false);
#undef SUG_GEMM_LAUNCH
  SUG_LAUNCH_CHECK();
  return 0;
}


// ---- skinny shapes (HBM-bound, no tile reuse to exploit) -------------------------------------------
// K <= 8 (the xyz layers: K = 3 / 6): every thread owns one row m, keeps A(m, 0..K) in registers and
// streams the N outputs with float4 stores; B (N x K) sits in shared memory.
__global__ void __launch_bounds__(256)
skinny_k_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm, long long sbn,
                long long sbk, const float* __restrict__ bias, float* __restrict__ C, long long ldc, int M, int N, int K) {
  extern __shared__ float sB[];  // [K][N] + bias[N]
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    int n = e % N, k = e / N;
    sB[k * N + n] = __ldg(Bm + (long long)n * sbn + (long long)k * sbk);
  }
  for (int n = threadIdx.x; n < N; n += blockDim.x) sB[K * N + n] = bias ? __ldg(bias + n) : 0.f;
  __syncthreads();
  const bool vec = (N % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  // a warp covers 32 consecutive n-quads of one row at a time => fully coalesced 512 B stores
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (long long m = (long long)blockIdx.x * 8 + wib; m < M; m += (long long)gridDim.x * 8) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = k < K ? __ldg(A + m * sam + (long long)k * sak) : 0.f;
    if (vec) {
      for (int n4 = lane; n4 < N / 4; n4 += 32) {
        float4 o = *reinterpret_cast<const float4*>(sB + K * N + 4 * n4);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < K) {
            const float4 b = *reinterpret_cast<const float4*>(sB + k * N + 4 * n4);
            o.x = fmaf(a[k], b.x, o.x); o.y = fmaf(a[k], b.y, o.y); o.z = fmaf(a[k], b.z, o.z); o.w = fmaf(a[k], b.w, o.w);
          }
        }
        *reinterpret_cast<float4*>(C + m * ldc + 4 * n4) = o;
      }
    } else {
      for (int n = lane; n < N; n += 32) {
        float o = sB[K * N + n];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < K) o = fmaf(a[k], sB[k * N + n], o);
        C[m * ldc + n] = o;
      }
    }
  }
}

// N <= 8, A k-contiguous, moderate K (e.g. the 64 -> 3 offset head): one thread per row m, B in smem.
__global__ void __launch_bounds__(256)
skinny_n_kernel(const float* __restrict__ A, long long sam, const float* __restrict__ Bm, long long sbn, long long sbk,
                const float* __restrict__ bias, float* __restrict__ C, long long ldc, int M, int N, int K) {
  extern __shared__ __align__(16) float sB[];  // [N][K]
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    int k = e % K, n = e / K;
    sB[n * K + k] = __ldg(Bm + (long long)n * sbn + (long long)k * sbk);
  }
  __syncthreads();
  const bool vec = (K % 4 == 0) && (sam % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = (bias && n < N) ? __ldg(bias + n) : 0.f;
    const float* ar = A + m * sam;
    if (vec) {
      for (int k = 0; k < K; k += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ar + k));
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          if (n < N) {
            const float4 w = *reinterpret_cast<const float4*>(sB + n * K + k);
            acc[n] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[n]))));
          }
        }
      }
    } else {
      for (int k = 0; k < K; ++k) {
        const float a = __ldg(ar + k);
#pragma unroll
        for (int n = 0; n < 8; ++n)
          if (n < N) acc[n] = fmaf(a, sB[n * K + k], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n < N) C[m * ldc + n] = acc[n];
  }
}

// min(M, N) <= 8 with a long reduction (weight gradients of the xyz layer / offset head):
// out(w, s) = sum_k Wd(w, k) * Sm(s, k).  A block covers TW wide indices x (256 / TW) interleaved
// k-slices, reduces the slices through shared memory and issues one atomic per output and block.
__global__ void __launch_bounds__(256)
skinny_reduce_kernel(const float* __restrict__ Wd, long long sww, long long swk, const float* __restrict__ Sm, long long sss,
                     long long ssk, float* __restrict__ C, long long cw, long long cs, int W, int S, int K, int kchunk,
                     int TW) {
  __shared__ float red[256][8];
  const int KS = 256 / TW;
  const int wl = threadIdx.x % TW, ks = threadIdx.x / TW;
  const int w = blockIdx.x * TW + wl;
  const int k0 = blockIdx.y * kchunk, k1 = min(K, k0 + kchunk);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (w < W) {
    // the operands stream from HBM exactly once: keep 4 independent rows in flight per thread
    int k = k0 + ks;
    for (; k + 3 * KS < k1; k += 4 * KS) {
      float a[4], sv[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldg(Wd + (long long)w * sww + (long long)(k + u * KS) * swk);
#pragma unroll
        for (int s = 0; s < 8; ++s) sv[u][s] = s < S ? __ldg(Sm + (long long)s * sss + (long long)(k + u * KS) * ssk) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int s = 0; s < 8; ++s) acc[s] = fmaf(a[u], sv[u][s], acc[s]);
    }
    for (; k < k1; k += KS) {
      const float a = __ldg(Wd + (long long)w * sww + (long long)k * swk);
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < S) acc[s] = fmaf(a, __ldg(Sm + (long long)s * sss + (long long)k * ssk), acc[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) red[threadIdx.x][s] = acc[s];
  __syncthreads();
  if (ks == 0 && w < W) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < S) {
        float t = 0.f;
        for (int r = 0; r < KS; ++r) t += red[r * TW + wl][s];
        atomicAdd(C + (long long)w * cw + (long long)s * cs, t);
      }
    }
  }
}

static int gemm_skinny(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk, const float* bias,
                       float* c, int64_t ldc, int M, int N, int K, cudaStream_t stream, bool* handled) {
  *handled = true;
  if (K <= 8 && (size_t)(K + 1) * N * sizeof(float) <= 40 * 1024) {
    ProfScope ps(KC_GEMM, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
    const int grid = (int)min((long long)num_sms() * 8, ((long long)M + 7) / 8);
    skinny_k_kernel<<<grid, 256, (size_t)(K + 1) * N * sizeof(float), stream>>>(a, sam, sak, b, sbn, sbk, bias, c, ldc, M, N, K);
    SUG_LAUNCH_CHECK();
    return 0;
  }
  if (N <= 8 && sak == 1 && K <= 2048 && M >= 4096) {
    ProfScope ps(KC_GEMM, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
    const int grid = (int)min((long long)num_sms() * 8, ((long long)M + 255) / 256);
    skinny_n_kernel<<<grid, 256, (size_t)N * K * sizeof(float), stream>>>(a, sam, b, sbn, sbk, bias, c, ldc, M, N, K);
    SUG_LAUNCH_CHECK();
    return 0;
  }
  if ((N <= 8 || M <= 8) && K >= 4096 && bias == nullptr && max(M, N) <= 4096) {
    ProfScope ps(KC_GEMM, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
    SUG_CUDA(cudaMemset2DAsync(c, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream));
    const bool wide_is_m = N <= 8;
    const int W = wide_is_m ? M : N, S = wide_is_m ? N : M;
    int TW = 32;
    while (TW < W && TW < 256) TW <<= 1;
    const int gx = cdiv(W, TW);
    int splits = max(1, min(cdiv(K, 256), (2 * num_sms()) / gx));
    const int kchunk = cdiv(K, splits);
    splits = cdiv(K, kchunk);
    if (wide_is_m)
      skinny_reduce_kernel<<<dim3(gx, splits), 256, 0, stream>>>(a, sam, sak, b, sbn, sbk, c, ldc, 1, W, S, K, kchunk, TW);
    else
      skinny_reduce_kernel<<<dim3(gx, splits), 256, 0, stream>>>(b, sbn, sbk, a, sam, sak, c, 1, ldc, W, S, K, kchunk, TW);
    SUG_LAUNCH_CHECK();
    return 0;
  }
  *handled = false;
  return 0;
}

// Dispatcher used by every entry point: the tcgen05 3xTF32 kernel whenever the operands satisfy
// the TMA constraints (unit stride on one axis, 16 B aligned base, leading dimension % 4 == 0) and
// the reduction is long enough to feed the tensor core; the CUDA-core kernel otherwise (xyz layers
// with K = 3, ragged leading dimensions such as the 4106-wide MMD features, accumulate mode).
int gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
             const float* bias, float* c, int64_t ldc, int M, int N, int K, int accumulate,
             cudaStream_t stream) {
  auto tma_ok = [](const float* p, int64_t s_row, int64_t s_k, int* mn, int64_t* ld) {
    if (s_k == 1) { *mn = 0; *ld = s_row; }
    else if (s_row == 1) { *mn = 1; *ld = s_k; }
    else return false;
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (*ld % 4) == 0 && *ld > 0;
  };
  if (!accumulate) {
    bool handled = false;
    SUG_TRY(gemm_skinny(a, sam, sak, b, sbn, sbk, bias, c, ldc, M, N, K, stream, &handled));
    if (handled) return 0;
  }
  int a_mn = 0, b_mn = 0;
  int64_t lda = 0, ldb = 0;
  if (!accumulate && K >= 16 && N >= 8 && tma_ok(a, sam, sak, &a_mn, &lda) && tma_ok(b, sbn, sbk, &b_mn, &ldb))
    return gemm_tc_f32(a, lda, a_mn, b, ldb, b_mn, bias, c, ldc, M, N, K, stream);
  return gemm_simt_f32(a, sam, sak, b, sbn, sbk, bias, c, ldc, M, N, K, accumulate, stream);
}

int gemm_f32_colstats(const float* a, int64_t lda, const float* b, int64_t ldb, const float* bias, float* c, int64_t ldc,
                      int M, int N, int K, double* colsums, bool* fused, cudaStream_t stream) {
  *fused = false;
  const bool tc_ok = K >= 16 && N >= 8 && M > 8 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && lda % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(b) & 15) == 0 && ldb % 4 == 0;
  if (tc_ok) return gemm_tc_stats_f32(a, lda, 0, b, ldb, 0, bias, c, ldc, M, N, K, colsums, fused, stream);
  return gemm_f32(a, lda, 1, b, ldb, 1, bias, c, ldc, M, N, K, 0, stream);
}

}  // namespace sug

// Same problem statement as sug_gemm_f32 but through the dispatcher (tensor cores when possible).
extern "C" int sug_gemm_auto_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                                 const float* bias, float* c, int64_t ldc, int M, int N, int K, sug_stream_t stream) {
  return sug::gemm_f32(a, sam, sak, b, sbn, sbk, bias, c, ldc, M, N, K, 0, (cudaStream_t)stream);
}

extern "C" int sug_gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                            const float* bias, float* c, int64_t ldc, int M, int N, int K, int accumulate,
                            sug_stream_t stream) {
  return sug::gemm_simt_f32(a, sam, sak, b, sbn, sbk, bias, c, ldc, M, N, K, accumulate, (cudaStream_t)stream);
}
