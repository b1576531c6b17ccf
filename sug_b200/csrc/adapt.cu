// Index builders of the self-adaptive node layer (reference: model/point_utils.py:5-165, called from
// adapt_layer_off.forward, model/model_utils.py:103-128).  The reference runs a 64-iteration Python
// loop with two boolean-mask host syncs per iteration plus two full sorts of [B,64,1024]; here each
// step is one launch with no host round trip.  Only indices are produced; the differentiable
// gathers / weights around them stay in the autograd graph of the host module.
//
// Floating-point forms follow the reference so that selections agree except on genuine ties:
//   FPS:        d = (dx*dx + dy*dy) + dz*dz on explicit differences  (point_utils.py:21)
//   the others: d = ((-2 s.d) + |s|^2) + |d|^2                        (point_utils.py:127-130)
#include "common.cuh"

namespace sug {

__device__ __forceinline__ float sq3(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fmaf(az, bz, fmaf(ay, by, __fmul_rn(ax, bx)));
}
__device__ __forceinline__ float sqdist_expanded(float dot, float nq, float np) {
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), nq), np);
}

// ---- farthest point sampling: one CTA per cloud ------------------------------------------------
__global__ void __launch_bounds__(1024)
fps_kernel(const float* __restrict__ xyz, int N, int npoint, const int* __restrict__ start, int* __restrict__ out) {
  extern __shared__ float sm[];
  float* px = sm;
  float* py = sm + N;
  float* pz = sm + 2 * N;
  float* dist = sm + 3 * N;
  __shared__ float wv[32];
  __shared__ int wi[32];
  __shared__ int s_far;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* xb = xyz + (size_t)b * 3 * N;
  for (int i = tid; i < N; i += nt) {
    px[i] = xb[i];
    py[i] = xb[N + i];
    pz[i] = xb[2 * N + i];
    dist[i] = 1e10f;
  }
  if (tid == 0) s_far = start[b];
  __syncthreads();
  for (int r = 0; r < npoint; ++r) {
    const int far = s_far;
    if (tid == 0) out[(size_t)b * npoint + r] = far;
    const float cx = px[far], cy = py[far], cz = pz[far];
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int i = tid; i < N; i += nt) {
      float d = sq3(px[i] - cx, py[i] - cy, pz[i] - cz);
      float cur = dist[i];
      if (d < cur) { cur = d; dist[i] = d; }
      if (cur > bv) { bv = cur; bi = i; }  // ascending i per thread: first maximum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();  // everyone has read s_far
    if ((tid & 31) == 0) { wv[tid >> 5] = bv; wi[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      float v = tid < (nt >> 5) ? wv[tid] : -2.f;
      int ii = tid < (nt >> 5) ? wi[tid] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, ii, o);
        if (ov > v || (ov == v && oi < ii)) { v = ov; ii = oi; }
      }
      if (tid == 0) s_far = ii;
    }
    __syncthreads();
  }
}

// ---- ball query: one warp per query -------------------------------------------------------------
__global__ void __launch_bounds__(256)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ query, int N, int S, float r2, int nsample,
                  int* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= S) return;
  const float* xb = xyz + (size_t)b * 3 * N;
  const float* qb = query + (size_t)b * 3 * S;
  const float qx = qb[warp], qy = qb[S + warp], qz = qb[2 * S + warp];
  const float nq = sq3(qx, qy, qz);
  int* o = out + ((size_t)b * S + warp) * nsample;
  int cnt = 0, first = -1;
  for (int i0 = 0; i0 < N && cnt < nsample; i0 += 32) {
    int i = i0 + lane;
    bool in = false;
    if (i < N) {
      float x = xb[i], y = xb[N + i], z = xb[2 * N + i];
      float d = sqdist_expanded(dot3(qx, qy, qz, x, y, z), nq, sq3(x, y, z));
      in = !(d > r2);
    }
    unsigned m = __ballot_sync(0xffffffffu, in);
    if (first < 0 && m != 0) first = i0 + __ffs(m) - 1;
    int pos = cnt + __popc(m & ((1u << lane) - 1));
    if (in && pos < nsample) o[pos] = i;
    cnt += __popc(m);
  }
  if (cnt > nsample) cnt = nsample;
  if (first < 0) first = 0;
  for (int p = cnt + lane; p < nsample; p += 32) o[p] = first;
}

// ---- nsample nearest points of every query, ascending: bitonic sort of (d, i) in shared memory ---
__global__ void __launch_bounds__(512)
knn_query_kernel(const float* __restrict__ xyz, const float* __restrict__ query, int N, int S, int NP2, int nsample,
                 int* __restrict__ out) {
  extern __shared__ float sm[];
  float* kv = sm;                                  // [NP2]
  int* ki = reinterpret_cast<int*>(sm + NP2);      // [NP2]
  const int s = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const float* xb = xyz + (size_t)b * 3 * N;
  const float* qb = query + (size_t)b * 3 * S;
  const float qx = qb[s], qy = qb[S + s], qz = qb[2 * S + s];
  const float nq = sq3(qx, qy, qz);
  for (int i = tid; i < NP2; i += nt) {
    float d = INFINITY;
    if (i < N) {
      float x = xb[i], y = xb[N + i], z = xb[2 * N + i];
      d = sqdist_expanded(dot3(qx, qy, qz, x, y, z), nq, sq3(x, y, z));
    }
    kv[i] = d;
    ki[i] = i;
  }
  __syncthreads();
  for (int size = 2; size <= NP2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (NP2 >> 1); t += nt) {
        int lo = (t / stride) * stride * 2 + (t % stride);
        int hi = lo + stride;
        bool asc = ((lo & size) == 0);
        float a = kv[lo], c = kv[hi];
        int ai = ki[lo], ci = ki[hi];
        bool gt = (a > c) || (a == c && ai > ci);
        if (gt == asc) { kv[lo] = c; kv[hi] = a; ki[lo] = ci; ki[hi] = ai; }
      }
      __syncthreads();
    }
  }
  int* o = out + ((size_t)b * S + s) * nsample;
  for (int p = tid; p < nsample; p += nt) o[p] = ki[p];
}

// ---- k (<= 8) nearest nodes of every point, ascending --------------------------------------------
__global__ void __launch_bounds__(256)
three_nn_kernel(const float* __restrict__ xyz, const float* __restrict__ nodes, int N, int M, int k,
                int* __restrict__ out) {
  extern __shared__ float sm[];  // nodes x,y,z,|.|^2 : [4][M]
  const int b = blockIdx.y;
  const float* nb = nodes + (size_t)b * 3 * M;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    float x = nb[j], y = nb[M + j], z = nb[2 * M + j];
    sm[j] = x; sm[M + j] = y; sm[2 * M + j] = z; sm[3 * M + j] = sq3(x, y, z);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* xb = xyz + (size_t)b * 3 * N;
  const float x = xb[i], y = xb[N + i], z = xb[2 * N + i];
  const float np = sq3(x, y, z);
  float bd[8];
  int bi[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { bd[u] = INFINITY; bi[u] = 0x7fffffff; }
  for (int j = 0; j < M; ++j) {
    float d = sqdist_expanded(dot3(x, y, z, sm[j], sm[M + j], sm[2 * M + j]), np, sm[3 * M + j]);
    int jj = j;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u < k && (d < bd[u] || (d == bd[u] && jj < bi[u]))) {
        float td = bd[u]; int ti = bi[u];
        bd[u] = d; bi[u] = jj;
        d = td; jj = ti;
      }
    }
  }
  int* o = out + ((size_t)b * N + i) * k;
  for (int u = 0; u < k; ++u) o[u] = bi[u];
}


// ---- nsample nearest points of every query as a SET (order: increasing point index) ----------------
// One warp per query, N <= 32*NPL distances in registers as order-preserving uint keys; the
// nsample-th smallest key is found by a 32-step bitwise search (one warp-wide count per bit), then
// the selected points are compacted with ballots.  ~20x cheaper than sorting the whole row, and all
// adapt_layer_off needs (the group is max-pooled, model_utils.py:121-123).
template <int NPL>
__global__ void __launch_bounds__(256)
knn_query_select_kernel(const float* __restrict__ xyz, const float* __restrict__ query, int N, int S, int nsample,
                        int* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= S) return;
  const float* xb = xyz + (size_t)b * 3 * N;
  const float* qb = query + (size_t)b * 3 * S;
  const float qx = qb[warp], qy = qb[S + warp], qz = qb[2 * S + warp];
  const float nq = sq3(qx, qy, qz);
  unsigned u[NPL];
#pragma unroll
  for (int t = 0; t < NPL; ++t) {
    const int i = lane + 32 * t;
    unsigned key = 0xffffffffu;
    if (i < N) {
      const float x = xb[i], y = xb[N + i], z = xb[2 * N + i];
      const float d = sqdist_expanded(dot3(qx, qy, qz, x, y, z), nq, sq3(x, y, z));
      const unsigned bits = __float_as_uint(d);
      key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
    }
    u[t] = key;
  }
  unsigned T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned cand = T | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < NPL; ++t) cnt += (u[t] < cand) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt < nsample) T = cand;
  }
  int c_lt = 0;
#pragma unroll
  for (int t = 0; t < NPL; ++t) c_lt += (u[t] < T) ? 1 : 0;
  c_lt = __reduce_add_sync(0xffffffffu, c_lt);
  const int need_eq = nsample - c_lt;
  int* o = out + ((size_t)b * S + warp) * nsample;
  const unsigned ltmask = (1u << lane) - 1u;
  int base_lt = 0, seen_eq = 0;
#pragma unroll
  for (int t = 0; t < NPL; ++t) {
    const bool lt = u[t] < T, eq = u[t] == T;
    const unsigned bl = __ballot_sync(0xffffffffu, lt), be = __ballot_sync(0xffffffffu, eq);
    if (lt) o[base_lt + __popc(bl & ltmask)] = lane + 32 * t;
    if (eq) {
      const int r = seen_eq + __popc(be & ltmask);
      if (r < need_eq) o[c_lt + r] = lane + 32 * t;
    }
    base_lt += __popc(bl);
    seen_eq += __popc(be);
  }
}

// ---- fused gather + max over a group (node features), with the arg for the backward ----------------
__global__ void __launch_bounds__(256)
group_max_fwd_kernel(const float* __restrict__ x, const int* __restrict__ idx, int N, int S, int K, int C,
                     float* __restrict__ out, int* __restrict__ arg) {
  const int s = blockIdx.x, b = blockIdx.y;
  const int* ip = idx + ((size_t)b * S + s) * K;
  const float* xb = x + (size_t)b * N * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float best = -INFINITY;
    int bj = 0;
    for (int k = 0; k < K; ++k) {
      const int j = __ldg(ip + k);
      const float v = __ldg(xb + (size_t)j * C + c);
      if (v > best) { best = v; bj = j; }
    }
    out[((size_t)b * S + s) * C + c] = best;
    arg[((size_t)b * S + s) * C + c] = bj;
  }
}
__global__ void __launch_bounds__(256)
group_max_bwd_kernel(const float* __restrict__ g, const int* __restrict__ arg, int N, int S, int C, float* __restrict__ dx) {
  const int s = blockIdx.x, b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t o = ((size_t)b * S + s) * C + c;
    atomicAdd(dx + ((size_t)b * N + arg[o]) * C + c, g[o]);
  }
}

// ---- weighted K-neighbour interpolation of node features back to the points -------------------------
// out[b,n,:] = sum_k w[b,n,k] f[b, idx[b,n,k], :]        (point_utils.py:158-160, upsample_inter)
__global__ void __launch_bounds__(256)
interp_fwd_kernel(const float* __restrict__ f, const int* __restrict__ idx, const float* __restrict__ w, int N, int S,
                  int K, int C, float* __restrict__ out) {
  extern __shared__ float sf[];  // [S][C]
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < S * C; e += blockDim.x) sf[e] = f[(size_t)b * S * C + e];
  __syncthreads();
  const int ppb = blockDim.x / C;  // points per block step (C <= 256 and divides 256)
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  for (int n = blockIdx.x * ppb + pl; n < N; n += gridDim.x * ppb) {
    const size_t r = (size_t)b * N + n;
    float o = 0.f;
    for (int k = 0; k < K; ++k) o = fmaf(__ldg(w + r * K + k), sf[__ldg(idx + r * K + k) * C + c], o);
    out[r * C + c] = o;
  }
}
// df[b,i,:] += w g ;  dw[b,n,k] = <g[b,n,:], f[b,i_k,:]>
__global__ void __launch_bounds__(256)
interp_bwd_kernel(const float* __restrict__ g, const float* __restrict__ f, const int* __restrict__ idx,
                  const float* __restrict__ w, int N, int S, int K, int C, float* __restrict__ df, float* __restrict__ dw) {
  extern __shared__ float sm[];  // f [S][C], acc [S][C]
  float* sf = sm;
  float* sa = sm + (size_t)S * C;
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < S * C; e += blockDim.x) {
    sf[e] = f[(size_t)b * S * C + e];
    sa[e] = 0.f;
  }
  __syncthreads();
  const int ppb = blockDim.x / C;
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  for (int n = blockIdx.x * ppb + pl; n < N; n += gridDim.x * ppb) {
    const size_t r = (size_t)b * N + n;
    const float gv = __ldg(g + r * C + c);
    for (int k = 0; k < K; ++k) {
      const int i = __ldg(idx + r * K + k);
      atomicAdd(&sa[i * C + c], __ldg(w + r * K + k) * gv);
      // dw: reduce gv * f over the C threads of this point (C is a multiple of 32)
      float d = gv * sf[i * C + c];
      d = warp_sum(d);
      if ((threadIdx.x & 31) == 0) atomicAdd(dw + r * K + k, d);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < S * C; e += blockDim.x) atomicAdd(df + (size_t)b * S * C + e, sa[e]);
}

}  // namespace sug

using namespace sug;

extern "C" int sug_fps(const float* xyz, int B, int N, int npoint, const int32_t* start, int32_t* out_idx,
                       sug_stream_t stream) {
  SUG_CHECK_ARG(xyz && start && out_idx, "fps: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && npoint > 0, "fps: bad shape");
  size_t smem = sizeof(float) * 4 * (size_t)N;
  SUG_CHECK_ARG(smem <= 227 * 1024, "fps: N=%d needs %zu B of shared memory", N, smem);
  SUG_TRY(ensure_dyn_smem((const void*)fps_kernel, smem));
  int nt = N >= 1024 ? 1024 : ((N + 31) / 32) * 32;
  ProfScope ps(KC_ADAPT, 0, 0, (cudaStream_t)stream);
  fps_kernel<<<B, nt, smem, (cudaStream_t)stream>>>(xyz, N, npoint, start, out_idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_ball_query(const float* xyz, const float* query, int B, int N, int S, float radius, int nsample,
                              int32_t* out_idx, sug_stream_t stream) {
  SUG_CHECK_ARG(xyz && query && out_idx, "ball_query: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && S > 0 && nsample > 0, "ball_query: bad shape");
  const float r2 = (float)((double)radius * (double)radius);
  dim3 grid(cdiv((long long)S * 32, 256), B);
  ProfScope ps(KC_ADAPT, 0, 0, (cudaStream_t)stream);
  ball_query_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, query, N, S, r2, nsample, out_idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_knn_query_set(const float* xyz, const float* query, int B, int N, int S, int nsample, int32_t* out_idx,
                                 sug_stream_t stream) {
  SUG_CHECK_ARG(xyz && query && out_idx, "knn_query_set: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && S > 0 && nsample > 0 && nsample <= N, "knn_query_set: bad shape");
  if (N > 2048) return sug_knn_query(xyz, query, B, N, S, nsample, out_idx, stream);
  dim3 grid(cdiv((long long)S * 32, 256), B);
  ProfScope ps(KC_ADAPT, 0, 0, (cudaStream_t)stream);
  if (N <= 1024) knn_query_select_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, query, N, S, nsample, out_idx);
  else knn_query_select_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, query, N, S, nsample, out_idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_group_max_fwd(const float* x, const int32_t* idx, int B, int N, int S, int K, int C, float* out,
                                 int32_t* arg, sug_stream_t stream) {
  SUG_CHECK_ARG(x && idx && out && arg && B > 0 && N > 0 && S > 0 && K > 0 && C > 0, "group_max_fwd: bad argument");
  ProfScope ps(KC_ADAPT, 0, 4.0 * B * S * ((double)K * C + 2.0 * C), (cudaStream_t)stream);
  group_max_fwd_kernel<<<dim3(S, B), C < 256 ? ((C + 31) / 32) * 32 : 256, 0, (cudaStream_t)stream>>>(x, idx, N, S, K, C, out, arg);
  SUG_LAUNCH_CHECK();
  return 0;
}
extern "C" int sug_group_max_bwd(const float* g, const int32_t* arg, int B, int N, int S, int C, float* dx,
                                 sug_stream_t stream) {
  SUG_CHECK_ARG(g && arg && dx && B > 0 && N > 0 && S > 0 && C > 0, "group_max_bwd: bad argument");
  ProfScope ps(KC_ADAPT, 0, 12.0 * B * S * C, (cudaStream_t)stream);
  group_max_bwd_kernel<<<dim3(S, B), C < 256 ? ((C + 31) / 32) * 32 : 256, 0, (cudaStream_t)stream>>>(g, arg, N, S, C, dx);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_interp_fwd(const float* f, const int32_t* idx, const float* w, int B, int N, int S, int K, int C,
                              float* out, sug_stream_t stream) {
  SUG_CHECK_ARG(f && idx && w && out && B > 0 && N > 0 && S > 0 && K > 0, "interp_fwd: bad argument");
  SUG_CHECK_ARG(C >= 32 && C <= 256 && 256 % C == 0 && (size_t)S * C * 4 <= 48 * 1024, "interp_fwd: C=%d S=%d unsupported", C, S);
  ProfScope ps(KC_ADAPT, 2.0 * B * N * K * C, 4.0 * B * N * (C + 2.0 * K), (cudaStream_t)stream);
  interp_fwd_kernel<<<dim3(8, B), 256, (size_t)S * C * 4, (cudaStream_t)stream>>>(f, idx, w, N, S, K, C, out);
  SUG_LAUNCH_CHECK();
  return 0;
}
// df [B,S,C] and dw [B,N,K] must be zero on entry (accumulated with atomics).
extern "C" int sug_interp_bwd(const float* g, const float* f, const int32_t* idx, const float* w, int B, int N, int S, int K,
                              int C, float* df, float* dw, sug_stream_t stream) {
  SUG_CHECK_ARG(g && f && idx && w && df && dw && B > 0 && N > 0 && S > 0 && K > 0, "interp_bwd: bad argument");
  SUG_CHECK_ARG(C >= 32 && C <= 256 && 256 % C == 0 && (size_t)S * C * 8 <= 48 * 1024, "interp_bwd: C=%d S=%d unsupported", C, S);
  ProfScope ps(KC_ADAPT, 4.0 * B * N * K * C, 4.0 * B * N * (C + 3.0 * K), (cudaStream_t)stream);
  interp_bwd_kernel<<<dim3(8, B), 256, (size_t)S * C * 8, (cudaStream_t)stream>>>(g, f, idx, w, N, S, K, C, df, dw);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_knn_query(const float* xyz, const float* query, int B, int N, int S, int nsample, int32_t* out_idx,
                             sug_stream_t stream) {
  SUG_CHECK_ARG(xyz && query && out_idx, "knn_query: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && S > 0 && nsample > 0 && nsample <= N, "knn_query: bad shape");
  int np2 = 2;
  while (np2 < N) np2 <<= 1;
  size_t smem = 8 * (size_t)np2;
  SUG_CHECK_ARG(smem <= 227 * 1024, "knn_query: N=%d too large", N);
  SUG_TRY(ensure_dyn_smem((const void*)knn_query_kernel, smem));
  int nt = np2 / 2 < 512 ? (np2 / 2 < 32 ? 32 : np2 / 2) : 512;
  ProfScope ps(KC_ADAPT, 0, 0, (cudaStream_t)stream);
  knn_query_kernel<<<dim3(S, B), nt, smem, (cudaStream_t)stream>>>(xyz, query, N, S, np2, nsample, out_idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_three_nn(const float* xyz, const float* nodes, int B, int N, int M, int k, int32_t* out_idx,
                            sug_stream_t stream) {
  SUG_CHECK_ARG(xyz && nodes && out_idx, "three_nn: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && M > 0 && k > 0 && k <= 8 && k <= M, "three_nn: bad shape");
  size_t smem = sizeof(float) * 4 * (size_t)M;
  SUG_CHECK_ARG(smem <= 48 * 1024, "three_nn: M=%d too large", M);
  ProfScope ps(KC_ADAPT, 0, 0, (cudaStream_t)stream);
  three_nn_kernel<<<dim3(cdiv(N, 256), B), 256, smem, (cudaStream_t)stream>>>(xyz, nodes, N, M, k, out_idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

// ---- node offsets and interpolation weights of the adapt layer, fused -----------------------------------
// (model_utils.py:107-117 and point_utils.py:134-160; the reference spells these out as ~25 tensor ops
// with three advanced-index gathers whose backward is a sort-based index_put each.)
namespace sug {

// node_offset[b,s,:] = mean_j tanh(h[g_j] - h[f]) * (loc[g_j] - loc[f]),  g = group[b,s,:], f = fidx[b,s]
// xyz is the reference's [B,3,N] layout; h is [B,N,3].
__global__ void node_offset_fwd_kernel(const float* __restrict__ h, const float* __restrict__ xyz,
                                       const int* __restrict__ fidx, const int* __restrict__ gidx, int N, int S, int G,
                                       float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= S) return;
  const float* hb = h + (size_t)b * N * 3;
  const float* xb = xyz + (size_t)b * 3 * N;
  const int f = __ldg(fidx + (size_t)b * S + warp);
  const float hf[3] = {__ldg(hb + f * 3), __ldg(hb + f * 3 + 1), __ldg(hb + f * 3 + 2)};
  const float lf[3] = {__ldg(xb + f), __ldg(xb + N + f), __ldg(xb + 2 * N + f)};
  float acc[3] = {0.f, 0.f, 0.f};
  for (int j = lane; j < G; j += 32) {
    const int g = __ldg(gidx + ((size_t)b * S + warp) * G + j);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += tanhf(__ldg(hb + g * 3 + c) - hf[c]) * (__ldg(xb + c * N + g) - lf[c]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) acc[c] = warp_sum(acc[c]);
  if (lane == 0) {
    const float inv = 1.f / (float)G;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((size_t)b * S + warp) * 3 + c] = acc[c] * inv;
  }
}

// dh (zeroed) += scatter of d node_offset through tanh'; float atomics (summation order is not fixed)
__global__ void node_offset_bwd_kernel(const float* __restrict__ go, const float* __restrict__ h,
                                       const float* __restrict__ xyz, const int* __restrict__ fidx,
                                       const int* __restrict__ gidx, int N, int S, int G, float* __restrict__ dh) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= S) return;
  const float* hb = h + (size_t)b * N * 3;
  const float* xb = xyz + (size_t)b * 3 * N;
  float* db = dh + (size_t)b * N * 3;
  const int f = __ldg(fidx + (size_t)b * S + warp);
  const float inv = 1.f / (float)G;
  float hf[3], lf[3], g3[3], accf[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    hf[c] = __ldg(hb + f * 3 + c);
    lf[c] = __ldg(xb + c * N + f);
    g3[c] = __ldg(go + ((size_t)b * S + warp) * 3 + c) * inv;
  }
  for (int j = lane; j < G; j += 32) {
    const int g = __ldg(gidx + ((size_t)b * S + warp) * G + j);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = tanhf(__ldg(hb + g * 3 + c) - hf[c]);
      const float d = g3[c] * (__ldg(xb + c * N + g) - lf[c]) * (1.f - t * t);
      atomicAdd(db + g * 3 + c, d);
      accf[c] += d;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) accf[c] = warp_sum(accf[c]);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(db + f * 3 + c, -accf[c]);
  }
}

// w[b,n,t] = (1/d_t) / sum_u (1/d_u),  d_t = max(|x_n|^2 + |y_t|^2 - 2 x_n.y_t, 1e-10),  y_t = nodes[b, idx[b,n,t]]
// nodes is [B,S,3]; K <= 8.
__global__ void interp_weight_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ nodes,
                                         const int* __restrict__ idx, int N, int S, int K, float* __restrict__ w) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (n >= N) return;
  const float* xb = xyz + (size_t)b * 3 * N;
  const float x[3] = {__ldg(xb + n), __ldg(xb + N + n), __ldg(xb + 2 * N + n)};
  const float xx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  float r[8], R = 0.f;
  for (int t = 0; t < K; ++t) {
    const float* y = nodes + ((size_t)b * S + __ldg(idx + ((size_t)b * N + n) * K + t)) * 3;
    const float y0 = __ldg(y), y1 = __ldg(y + 1), y2 = __ldg(y + 2);
    const float dot = x[0] * y0 + x[1] * y1 + x[2] * y2;
    float d = -2.f * dot + xx + (y0 * y0 + y1 * y1 + y2 * y2);
    d = d < 1e-10f ? 1e-10f : d;
    r[t] = 1.f / d;
    R += r[t];
  }
  for (int t = 0; t < K; ++t) w[((size_t)b * N + n) * K + t] = r[t] / R;
}

// d nodes (zeroed) += scatter; float atomics
__global__ void interp_weight_bwd_kernel(const float* __restrict__ gw, const float* __restrict__ xyz,
                                         const float* __restrict__ nodes, const int* __restrict__ idx, int N, int S, int K,
                                         float* __restrict__ dnodes) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (n >= N) return;
  const float* xb = xyz + (size_t)b * 3 * N;
  const float x[3] = {__ldg(xb + n), __ldg(xb + N + n), __ldg(xb + 2 * N + n)};
  const float xx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  float r[8], y[8][3], R = 0.f, gdotw = 0.f;
  bool clamped[8];
  int id[8];
  for (int t = 0; t < K; ++t) {
    id[t] = __ldg(idx + ((size_t)b * N + n) * K + t);
    const float* yp = nodes + ((size_t)b * S + id[t]) * 3;
    y[t][0] = __ldg(yp); y[t][1] = __ldg(yp + 1); y[t][2] = __ldg(yp + 2);
    const float dot = x[0] * y[t][0] + x[1] * y[t][1] + x[2] * y[t][2];
    float d = -2.f * dot + xx + (y[t][0] * y[t][0] + y[t][1] * y[t][1] + y[t][2] * y[t][2]);
    clamped[t] = d < 1e-10f;
    d = clamped[t] ? 1e-10f : d;
    r[t] = 1.f / d;
    R += r[t];
  }
  for (int t = 0; t < K; ++t) gdotw += __ldg(gw + ((size_t)b * N + n) * K + t) * (r[t] / R);
  for (int t = 0; t < K; ++t) {
    if (clamped[t]) continue;  // the reference's masked assignment cuts the gradient
    const float dr = (__ldg(gw + ((size_t)b * N + n) * K + t) - gdotw) / R;  // dL/dr_t
    const float dd = -dr * r[t] * r[t];                                      // dL/dd_t
    float* o = dnodes + ((size_t)b * S + id[t]) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(o + c, dd * 2.f * (y[t][c] - x[c]));
  }
}

}  // namespace sug

extern "C" int sug_node_offset_fwd(const float* h, const float* xyz, const int32_t* fidx, const int32_t* gidx, int B,
                                   int N, int S, int G, float* out, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(h && xyz && fidx && gidx && out && B > 0 && N > 0 && S > 0 && G > 0, "node_offset_fwd: bad argument");
  ProfScope ps(KC_ADAPT, 12.0 * B * S * G, 4.0 * B * (6.0 * N + S * (G + 4.0)), (cudaStream_t)stream);
  node_offset_fwd_kernel<<<dim3(cdiv((long long)S * 32, 256), B), 256, 0, (cudaStream_t)stream>>>(h, xyz, fidx, gidx, N, S, G, out);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_node_offset_bwd(const float* gout, const float* h, const float* xyz, const int32_t* fidx,
                                   const int32_t* gidx, int B, int N, int S, int G, float* dh, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(gout && h && xyz && fidx && gidx && dh && B > 0 && N > 0 && S > 0 && G > 0, "node_offset_bwd: bad argument");
  ProfScope ps(KC_ADAPT, 20.0 * B * S * G, 4.0 * B * (9.0 * N + S * (G + 4.0)), (cudaStream_t)stream);
  node_offset_bwd_kernel<<<dim3(cdiv((long long)S * 32, 256), B), 256, 0, (cudaStream_t)stream>>>(gout, h, xyz, fidx, gidx, N, S, G, dh);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_interp_weight_fwd(const float* xyz, const float* nodes, const int32_t* idx, int B, int N, int S, int K,
                                     float* w, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(xyz && nodes && idx && w && B > 0 && N > 0 && S > 0 && K > 0 && K <= 8, "interp_weight_fwd: bad argument");
  ProfScope ps(KC_ADAPT, 20.0 * B * N * K, 4.0 * B * (3.0 * N + 3.0 * S + 2.0 * N * K), (cudaStream_t)stream);
  interp_weight_fwd_kernel<<<dim3(cdiv(N, 256), B), 256, 0, (cudaStream_t)stream>>>(xyz, nodes, idx, N, S, K, w);
  SUG_LAUNCH_CHECK();
  return 0;
}

extern "C" int sug_interp_weight_bwd(const float* gw, const float* xyz, const float* nodes, const int32_t* idx, int B,
                                     int N, int S, int K, float* dnodes, sug_stream_t stream) {
  using namespace sug;
  SUG_CHECK_ARG(gw && xyz && nodes && idx && dnodes && B > 0 && N > 0 && S > 0 && K > 0 && K <= 8, "interp_weight_bwd: bad argument");
  ProfScope ps(KC_ADAPT, 40.0 * B * N * K, 4.0 * B * (3.0 * N + 6.0 * S + 2.0 * N * K), (cudaStream_t)stream);
  interp_weight_bwd_kernel<<<dim3(cdiv(N, 256), B), 256, 0, (cudaStream_t)stream>>>(gw, xyz, nodes, idx, N, S, K, dnodes);
  SUG_LAUNCH_CHECK();
  return 0;
}
