// Error reporting, device queries and the small BatchNorm-statistics kernels.
#include <stdarg.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

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

int ensure_dyn_smem(const void* fn, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  SUG_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& have = done[std::make_pair(fn, dev)];
  if (bytes > have) {
    SUG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
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

// ---- profiling -----------------------------------------------------------------------------------
static const char* kClassNames[KC_NUM] = {"gemm_simt", "knn_simt", "knn_reverse", "edge_gather_fwd", "bn_act",
                                          "edge_bwd_pre", "edge_bwd_main", "col_stats", "pool_fwd", "pool_bwd",
                                          "mmd", "chamfer", "adapt_index", "misc", "gemm_tc", "knn_tc"};
struct ProfState {
  std::mutex mu;
  unsigned mask = 0;
  long long launches[KC_NUM] = {0};
  double flops[KC_NUM] = {0};
  double bytes[KC_NUM] = {0};
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<int> ev_cls;
  size_t used = 0;               // pairs in use
};
static ProfState g_prof;

ProfScope::ProfScope(int c, double fl, double by, cudaStream_t s) : cls(c), stream(s), slot(-1), capturing(false) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.launches[c] += 1;
  g_prof.flops[c] += fl;
  g_prof.bytes[c] += by;
  if (g_prof.mask & (1u << c)) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    capturing = st != cudaStreamCaptureStatusNone;
    if (g_prof.used < 400000) {
      if (g_prof.ev.size() < 2 * (g_prof.used + 1)) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        g_prof.ev.push_back(a);
        g_prof.ev.push_back(b);
        g_prof.ev_cls.push_back(c);
      }
      slot = (int)g_prof.used++;
      g_prof.ev_cls[slot] = c;
      // inside a stream capture the events become external event-record nodes of the graph, so every
      // replay re-times the launch (bench.py reads the last replay)
      if (capturing) cudaEventRecordWithFlags(g_prof.ev[2 * slot], s, cudaEventRecordExternal);
      else cudaEventRecord(g_prof.ev[2 * slot], s);
    }
  }
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  if (capturing) cudaEventRecordWithFlags(g_prof.ev[2 * slot + 1], stream, cudaEventRecordExternal);
  else cudaEventRecord(g_prof.ev[2 * slot + 1], stream);
}

int bn_finalize_stats(const double* sums, int C, double count, float eps, float momentum, float* running_mean,
                      float* running_var, float* save, cudaStream_t stream) {
  ProfScope ps(KC_MISC, 0, 0, stream);
  bn_finalize_stats_kernel<<<cdiv(C, 128), 128, 0, stream>>>(sums, C, count, eps, momentum, running_mean,
                                                              running_var, save);
  SUG_LAUNCH_CHECK();
  return 0;
}

int bn_eval_stats(const float* rm, const float* rv, int C, float eps, float* save, cudaStream_t stream) {
  ProfScope ps(KC_MISC, 0, 0, stream);
  bn_eval_stats_kernel<<<cdiv(C, 128), 128, 0, stream>>>(rm, rv, C, eps, save);
  SUG_LAUNCH_CHECK();
  return 0;
}

}  // namespace sug

extern "C" int sug_version(void) { return 100; }
extern "C" const char* sug_last_error(void) { return sug::g_err; }

// ---- profiling C ABI -----------------------------------------------------------------------------
extern "C" int sug_prof_num_classes(void) { return sug::KC_NUM; }
extern "C" const char* sug_prof_class_name(int i) { return (i >= 0 && i < sug::KC_NUM) ? sug::kClassNames[i] : ""; }
extern "C" void sug_prof_enable(unsigned mask) {
  std::lock_guard<std::mutex> lk(sug::g_prof.mu);
  sug::g_prof.mask = mask;
}
extern "C" void sug_prof_reset(void) {
  std::lock_guard<std::mutex> lk(sug::g_prof.mu);
  for (int i = 0; i < sug::KC_NUM; ++i) {
    sug::g_prof.launches[i] = 0;
    sug::g_prof.flops[i] = 0;
    sug::g_prof.bytes[i] = 0;
  }
  sug::g_prof.used = 0;
}
// Synchronises the device, then fills per-class totals: timed milliseconds, timed launch count,
// all launches, algorithmic flops and bytes (arrays of sug_prof_num_classes() entries).
extern "C" int sug_prof_collect(double* ms, long long* timed, long long* launches, double* flops, double* bytes) {
  SUG_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(sug::g_prof.mu);
  for (int i = 0; i < sug::KC_NUM; ++i) {
    ms[i] = 0;
    timed[i] = 0;
    launches[i] = sug::g_prof.launches[i];
    flops[i] = sug::g_prof.flops[i];
    bytes[i] = sug::g_prof.bytes[i];
  }
  for (size_t p = 0; p < sug::g_prof.used; ++p) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sug::g_prof.ev[2 * p], sug::g_prof.ev[2 * p + 1]) == cudaSuccess) {
      ms[sug::g_prof.ev_cls[p]] += t;
      timed[sug::g_prof.ev_cls[p]] += 1;
    }
  }
  return 0;
}
