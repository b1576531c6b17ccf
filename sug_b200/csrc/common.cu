// Error reporting, device queries and the small BatchNorm-statistics kernels.
#include <stdarg.h>

#include "common.cuh"

namespace sug {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// mean/invstd from fp64 sums; running stats follow nn.BatchNorm (momentum, unbiased variance).
__global__ void bn_finalize_stats_kernel(const double* __restrict__ sums, int C, double count, float eps,
                                         float momentum, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, float* __restrict__ save) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  save[c] = (float)mean;
  save[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                     float* __restrict__ save) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save[c] = rm[c];
  save[C + c] = 1.0f / sqrtf(rv[c] + eps);
}

int bn_finalize_stats(const double* sums, int C, double count, float eps, float momentum, float* running_mean,
                      float* running_var, float* save, cudaStream_t stream) {
  bn_finalize_stats_kernel<<<cdiv(C, 128), 128, 0, stream>>>(sums, C, count, eps, momentum, running_mean,
                                                              running_var, save);
  SUG_LAUNCH_CHECK();
  return 0;
}

int bn_eval_stats(const float* rm, const float* rv, int C, float eps, float* save, cudaStream_t stream) {
  bn_eval_stats_kernel<<<cdiv(C, 128), 128, 0, stream>>>(rm, rv, C, eps, save);
  SUG_LAUNCH_CHECK();
  return 0;
}

}  // namespace sug

extern "C" int sug_version(void) { return 100; }
extern "C" const char* sug_last_error(void) { return sug::g_err; }
