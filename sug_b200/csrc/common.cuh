// Shared helpers for libsug_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>
#include <math.h>

#include "../../include/sug_b200.h"

namespace sug {

void set_error(const char* fmt, ...);

#define SUG_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      sug::set_error(__VA_ARGS__);              \
      return SUG_E_BADARG;                      \
    }                                           \
  } while (0)

#define SUG_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      sug::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                     __LINE__);                                                          \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define SUG_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      sug::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__,   \
                     __LINE__);                                                              \
      return (int)_e;                                                                        \
    }                                                                                        \
  } while (0)

#define SUG_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();
// Opt a kernel in to `bytes` of dynamic shared memory (cudaFuncAttributeMaxDynamicSharedMemorySize) on the
// CURRENT device; remembered per (kernel, device), so a process that drives several GPUs configures each.
int ensure_dyn_smem(const void* fn, size_t bytes);

// ---- launch accounting / optional per-kernel-class CUDA-event timing (bench.py roofline) -------
enum KClass {
  KC_GEMM = 0, KC_KNN, KC_KNN_REV, KC_EDGE_FWD, KC_BN_ACT, KC_EDGE_BWD_PRE, KC_EDGE_BWD_MAIN, KC_COLSTATS,
  KC_POOL_FWD, KC_POOL_BWD, KC_MMD, KC_CHAMFER, KC_ADAPT, KC_MISC, KC_GEMM_TC, KC_KNN_TC, KC_NUM
};
// Counts one kernel launch of class `cls` with its algorithmic flops / bytes (DESIGN.md states the
// formulas); when timing of the class is enabled, brackets the launch with CUDA events on `stream`.
struct ProfScope {
  int cls;
  cudaStream_t stream;
  int slot;
  bool capturing;
  ProfScope(int cls, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
};

// Simple bump allocator over the caller's workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t o = align_up(off, 256);
    size_t need = o + count * sizeof(T);
    if (base == nullptr || need > size) {
      off = (size_t)-1;
      return nullptr;
    }
    off = need;
    return (T*)(base + o);
  }
  bool ok() const { return off != (size_t)-1; }
};

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float act_leaky(float z, float slope) { return z > 0.f ? z : z * slope; }
__device__ __forceinline__ float act_leaky_grad(float z, float slope) { return z > 0.f ? 1.f : slope; }

// internal launchers shared between translation units -----------------------------------------
// C[M,N] (+)= A B^T (+bias) with A(m,k) = a[m*sam + k*sak], B(n,k) = b[n*sbn + k*sbk]: dispatches to the
// tcgen05 3xTF32 kernel (gemm_tc.cu) or the exact-fp32 CUDA-core kernel (gemm.cu).
int gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
             const float* bias, float* c, int64_t ldc, int M, int N, int K, int accumulate,
             cudaStream_t stream);
int gemm_simt_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                  const float* bias, float* c, int64_t ldc, int M, int N, int K, int accumulate,
                  cudaStream_t stream);
int gemm_tc_f32(const float* a, int64_t lda, int a_mn, const float* b, int64_t ldb, int b_mn, const float* bias,
                float* c, int64_t ldc, int M, int N, int K, cudaStream_t stream);
int gemm_tc_stats_f32(const float* a, int64_t lda, int a_mn, const float* b, int64_t ldb, int b_mn, const float* bias,
                      float* c, int64_t ldc, int M, int N, int K, double* colsums, bool* fused, cudaStream_t stream);
// gemm_f32 for a K-major A and B (x W^T + bias) that also accumulates the BatchNorm column statistics of C into
// colsums[2*N] (fp64, zeroed by the caller) when the tensor-core epilogue can do it; *fused reports whether it did.
int gemm_f32_colstats(const float* a, int64_t lda, const float* b, int64_t ldb, const float* bias, float* c, int64_t ldc,
                      int M, int N, int K, double* colsums, bool* fused, cudaStream_t stream);

// Per-channel BatchNorm batch statistics -> (mean, invstd), running-stat update.
// sums: [2*C] doubles (sum, sum of squares) over `count` samples.
int bn_finalize_stats(const double* sums, int C, double count, float eps, float momentum,
                      float* running_mean, float* running_var, float* save_mean_invstd,
                      cudaStream_t stream);
// out = act(BN(y)) over [P, Cout] rows (edgeconv.cu): statistics given as (mean, invstd), or finalised from fp64 sums by
// the kernel itself (which then also writes `save` = (mean, invstd) and updates the running statistics).
int bn_act_launch(const float* y, const float* gamma, const float* beta, const float* mean_invstd, long long P, int Cout,
                  float slope, float* out, long long ldo, cudaStream_t stream);
int bn_act_from_sums_launch(const float* y, const float* gamma, const float* beta, const double* sums, double count, float eps,
                            float momentum, float* running_mean, float* running_var, float* save, long long P, int Cout,
                            float slope, float* out, long long ldo, cudaStream_t stream);
// (mean, invstd) from running statistics (eval mode).
int bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                  float* save_mean_invstd, cudaStream_t stream);

}  // namespace sug
