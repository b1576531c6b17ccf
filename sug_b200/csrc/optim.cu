// Multi-tensor Adam for the trainer step (reference: torch.optim.Adam as built in
// train_dg_single_gpu.py:191-203 -- L2 weight decay folded into the gradient, bias-corrected moments).
//
// One launch updates every tensor of an optimizer: the host passes device tables of pointers and
// sizes plus a (block -> tensor, chunk) map, so a step is  28 B / parameter  of HBM traffic in one
// streaming pass instead of a dozen foreach passes.  The step counter and the learning rate live in
// device memory, which makes the update capturable in a CUDA graph and lets LR schedulers change the
// rate without re-capturing.
#include "common.cuh"

namespace sug {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = ADAM_THREADS * 16;  // elements per block

__global__ void adam_tick_kernel(float* step) { *step += 1.f; }

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const long long* __restrict__ p_ptrs, const long long* __restrict__ g_ptrs,
            const long long* __restrict__ m_ptrs, const long long* __restrict__ v_ptrs,
            const long long* __restrict__ sizes, const int* __restrict__ blk_tensor,
            const int* __restrict__ blk_chunk, const float* __restrict__ step_p, const float* __restrict__ lr_p,
            float beta1, float beta2, float eps, float wd) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = sizes[t];
  const long long lo = (long long)blk_chunk[blockIdx.x] * ADAM_CHUNK;
  const long long hi = min(n, lo + ADAM_CHUNK);
  float* __restrict__ p = reinterpret_cast<float*>(p_ptrs[t]);
  const float* __restrict__ g = reinterpret_cast<const float*>(g_ptrs[t]);
  float* __restrict__ m = reinterpret_cast<float*>(m_ptrs[t]);
  float* __restrict__ v = reinterpret_cast<float*>(v_ptrs[t]);
  const float step = *step_p, lr = *lr_p;
  // same arithmetic as torch's capturable single-tensor Adam
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = fmaf(wd, pp, gg);
    mm = fmaf(beta1, mm, (1.f - beta1) * gg);  // lerp form: m + (g - m)(1 - b1)
    vv = fmaf(beta2, vv, (1.f - beta2) * gg * gg);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * (mm / denom);
  };
  if (vec) {
    for (long long i = lo + threadIdx.x * 4; i < hi; i += ADAM_THREADS * 4) {
      if (i + 4 <= hi) {
        float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
               vv = *reinterpret_cast<float4*>(v + i);
        const float4 gg = *reinterpret_cast<const float4*>(g + i);
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        *reinterpret_cast<float4*>(p + i) = pp;
        *reinterpret_cast<float4*>(m + i) = mm;
        *reinterpret_cast<float4*>(v + i) = vv;
      } else {
        for (long long j = i; j < hi; ++j) upd(p[j], g[j], m[j], v[j]);
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += ADAM_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}

// ---- all param groups of an optimizer in one launch -------------------------------------------------------------
// train_dg_single_gpu.py:191 builds optimizer_g with ONE PARAM GROUP PER PARAMETER (~40 groups): per-group launches
// were 50 launches per step.  Here every tensor carries pointers to ITS group's step counter and learning rate and its
// group's hyper-parameters, so the semantics stay per group (torch.optim.Adam) while a step is two launches.
__global__ void adam_tick_multi_kernel(const long long* __restrict__ step_ptrs, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) *reinterpret_cast<float*>(step_ptrs[i]) += 1.f;
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_multi_kernel(const long long* __restrict__ p_ptrs, const long long* __restrict__ g_ptrs,
                  const long long* __restrict__ m_ptrs, const long long* __restrict__ v_ptrs,
                  const long long* __restrict__ sizes, const long long* __restrict__ step_ptrs,
                  const long long* __restrict__ lr_ptrs, const float* __restrict__ hyper,
                  const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = sizes[t];
  const long long lo = (long long)blk_chunk[blockIdx.x] * ADAM_CHUNK;
  const long long hi = min(n, lo + ADAM_CHUNK);
  float* __restrict__ p = reinterpret_cast<float*>(p_ptrs[t]);
  const float* __restrict__ g = reinterpret_cast<const float*>(g_ptrs[t]);
  float* __restrict__ m = reinterpret_cast<float*>(m_ptrs[t]);
  float* __restrict__ v = reinterpret_cast<float*>(v_ptrs[t]);
  const float step = *reinterpret_cast<const float*>(step_ptrs[t]), lr = *reinterpret_cast<const float*>(lr_ptrs[t]);
  const float beta1 = hyper[4 * t], beta2 = hyper[4 * t + 1], eps = hyper[4 * t + 2], wd = hyper[4 * t + 3];
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = fmaf(wd, pp, gg);
    mm = fmaf(beta1, mm, (1.f - beta1) * gg);
    vv = fmaf(beta2, vv, (1.f - beta2) * gg * gg);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * (mm / denom);
  };
  if (vec) {
    for (long long i = lo + threadIdx.x * 4; i < hi; i += ADAM_THREADS * 4) {
      if (i + 4 <= hi) {
        float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
               vv = *reinterpret_cast<float4*>(v + i);
        const float4 gg = *reinterpret_cast<const float4*>(g + i);
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        *reinterpret_cast<float4*>(p + i) = pp;
        *reinterpret_cast<float4*>(m + i) = mm;
        *reinterpret_cast<float4*>(v + i) = vv;
      } else {
        for (long long j = i; j < hi; ++j) upd(p[j], g[j], m[j], v[j]);
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += ADAM_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace sug

extern "C" int sug_adam_multi_f32(const int64_t* p_ptrs, const int64_t* g_ptrs, const int64_t* m_ptrs, const int64_t* v_ptrs,
                                  const int64_t* sizes, const int64_t* step_ptrs, const int64_t* lr_ptrs, const float* hyper,
                                  const int32_t* blk_tensor, const int32_t* blk_chunk, int n_blocks, long long n_params,
                                  const int64_t* group_step_ptrs, int n_groups, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(p_ptrs && g_ptrs && m_ptrs && v_ptrs && sizes && step_ptrs && lr_ptrs && hyper && blk_tensor && blk_chunk &&
                    group_step_ptrs,
                "adam_multi: null pointer");
  SUG_CHECK_ARG(n_blocks >= 0 && n_groups > 0, "adam_multi: bad counts %d %d", n_blocks, n_groups);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(KC_MISC, 12.0 * n_params, 28.0 * n_params, st);
  adam_tick_multi_kernel<<<cdiv(n_groups, 128), 128, 0, st>>>(reinterpret_cast<const long long*>(group_step_ptrs), n_groups);
  if (n_blocks > 0)
    adam_multi_kernel<<<n_blocks, ADAM_THREADS, 0, st>>>(
        reinterpret_cast<const long long*>(p_ptrs), reinterpret_cast<const long long*>(g_ptrs),
        reinterpret_cast<const long long*>(m_ptrs), reinterpret_cast<const long long*>(v_ptrs),
        reinterpret_cast<const long long*>(sizes), reinterpret_cast<const long long*>(step_ptrs),
        reinterpret_cast<const long long*>(lr_ptrs), hyper, blk_tensor, blk_chunk);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_adam_chunk(void) { return sug::ADAM_CHUNK; }

extern "C" int sug_adam_f32(const int64_t* p_ptrs, const int64_t* g_ptrs, const int64_t* m_ptrs, const int64_t* v_ptrs,
                            const int64_t* sizes, const int32_t* blk_tensor, const int32_t* blk_chunk, int n_blocks,
                            long long n_params, float* step, const float* lr, float beta1, float beta2, float eps,
                            float weight_decay, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(p_ptrs && g_ptrs && m_ptrs && v_ptrs && sizes && blk_tensor && blk_chunk && step && lr, "adam: null pointer");
  SUG_CHECK_ARG(n_blocks >= 0, "adam: bad block count %d", n_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(KC_MISC, 12.0 * n_params, 28.0 * n_params, st);
  adam_tick_kernel<<<1, 1, 0, st>>>(step);
  if (n_blocks > 0)
    adam_kernel<<<n_blocks, ADAM_THREADS, 0, st>>>(
        reinterpret_cast<const long long*>(p_ptrs), reinterpret_cast<const long long*>(g_ptrs),
        reinterpret_cast<const long long*>(m_ptrs), reinterpret_cast<const long long*>(v_ptrs),
        reinterpret_cast<const long long*>(sizes), blk_tensor, blk_chunk, step, lr, beta1, beta2, eps, weight_decay);
  SUG_LAUNCH_CHECK();
  return 0;
}
