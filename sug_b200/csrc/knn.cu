// Fused pairwise-distance + top-k (reference: model/model_utils.py:178-185, knn()).
//
// One CTA owns 128 query points of one cloud and streams the cloud's candidates in tiles of 64;
// the N x N matrix never exists.  The ranking key reproduces the reference's fp32 expression
//     D_ij = (-|x_i|^2 - (-2 x_i.x_j)) - |x_j|^2
// with the dot product accumulated by sequential fp32 FMAs.  Selection: every thread owns one
// query row; candidates that beat the row's current k-th best are appended to a small pending
// queue (predicated stores, no divergence) and the queues of a warp are drained together, which
// keeps the divergent "replace the minimum and rescan" step busy on most lanes.
//
// This CUDA-core kernel is the production path for xyz inputs (C = 3).  Wider feature kNN uses
// the same selection code behind the tcgen05 distance tiles (knn_tc.cu).
#include <type_traits>

#include "knn_select.cuh"

namespace sug {

constexpr int KTN = 64;          // candidates per tile
constexpr int KCC = 32;          // channels per staged candidate chunk
constexpr int KQLD = KTM + 4;    // padded leading dims (16B aligned rows, conflict-free transposed stores)
constexpr int KCLD = KTN + 4;

// K > 0: compile-time k with the register-resident sorted list; K == 0: any k, heap in shared memory.
template <int K>
__global__ void __launch_bounds__(KTM)
knn_simt_kernel(const float* __restrict__ x, int C, int N, int k, long long sb, long long sn, long long sc,
                int* __restrict__ idx_out) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                          // [C][KQLD]   query tile, channel-major
  float* Cs = Qs + (size_t)C * KQLD;         // [KCC][KCLD] candidate chunk
  float* xxc = Cs + KCC * KCLD;              // [KTN]       |x_j|^2 of the current tile
  typename std::conditional<(K > 0), TopKReg<(K > 0 ? K : 1)>, TopK>::type tk;
  if constexpr (K > 0) tk.bind(xxc + KTN);
  else tk.bind(xxc + KTN, k);

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * KTM;
  const float* xb = x + (long long)b * sb;
  const bool point_major = (sc == 1);

  // ---- query tile -> shared (channel-major), query norms in registers --------------------------
  if (point_major) {
    for (int e = tid; e < C * KTM; e += KTM) {
      int c = e % C, r = e / C;
      int gr = row0 + r;
      Qs[c * KQLD + r] = gr < N ? __ldg(xb + (long long)gr * sn + c) : 0.f;
    }
  } else {
    for (int e = tid; e < C * KTM; e += KTM) {
      int r = e % KTM, c = e / KTM;
      int gr = row0 + r;
      Qs[c * KQLD + r] = gr < N ? __ldg(xb + (long long)gr * sn + (long long)c * sc) : 0.f;
    }
  }
  if constexpr (K > 0) tk.init();
  else tk.init(tid);
  __syncthreads();
  float xxq = 0.f;
  for (int c = 0; c < C; ++c) {
    float q = Qs[c * KQLD + tid];
    xxq = fmaf(q, q, xxq);
  }

  for (int col0 = 0; col0 < N; col0 += KTN) {
    float acc[KTN];
#pragma unroll
    for (int j = 0; j < KTN; ++j) acc[j] = 0.f;
    float xx_part = 0.f;  // threads < KTN: |x_j|^2 of candidate col0 + tid

    for (int c0 = 0; c0 < C; c0 += KCC) {
      const int cc = min(KCC, C - c0);
      __syncthreads();  // previous chunk (and xxc) fully consumed
      if (point_major) {
        for (int e = tid; e < cc * KTN; e += KTM) {
          int c = e % cc, j = e / cc;
          int gj = col0 + j;
          Cs[c * KCLD + j] = gj < N ? __ldg(xb + (long long)gj * sn + (c0 + c)) : 0.f;
        }
      } else {
        for (int e = tid; e < cc * KTN; e += KTM) {
          int j = e % KTN, c = e / KTN;
          int gj = col0 + j;
          Cs[c * KCLD + j] = gj < N ? __ldg(xb + (long long)gj * sn + (long long)(c0 + c) * sc) : 0.f;
        }
      }
      __syncthreads();
      if (tid < KTN) {
        for (int c = 0; c < cc; ++c) {
          float v = Cs[c * KCLD + tid];
          xx_part = fmaf(v, v, xx_part);
        }
      }
      for (int c = 0; c < cc; ++c) {
        const float q = Qs[(c0 + c) * KQLD + tid];
        const float4* cp = reinterpret_cast<const float4*>(Cs + c * KCLD);
#pragma unroll
        for (int j4 = 0; j4 < KTN / 4; ++j4) {
          float4 v = cp[j4];
          acc[4 * j4 + 0] = fmaf(q, v.x, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(q, v.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(q, v.z, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(q, v.w, acc[4 * j4 + 3]);
        }
      }
    }
    if (tid < KTN) xxc[tid] = xx_part;
    __syncthreads();

    // ---- selection over this tile's 64 keys, 32 at a time -------------------------------------
#pragma unroll
    for (int h = 0; h < KTN / 32; ++h) {
      float keys[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) keys[q] = fmaf(2.f, acc[h * 32 + q], -xxq) - xxc[h * 32 + q];
      const int base = col0 + h * 32;
      const int nvalid = N - base;
      const uint32_t valid = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
      tk.consider32(tid, keys, valid, base);
    }
  }
  int* ranked = reinterpret_cast<int*>(xxc + KTN);  // [k][KTM] ranked indices for the coalesced write
  if constexpr (K > 0) {
    __syncthreads();  // every thread is done with the stash before it is reused
#pragma unroll
    for (int s = 0; s < K; ++s) ranked[s * KTM + tid] = tk.id[s];
  } else {
    tk.sort_desc(tid);
    ranked = tk.hi;
  }
  __syncthreads();

  // ---- coalesced write of the tile's [rows][k] block --------------------------------------------
  const int rows = min(KTM, N - row0);
  int* out = idx_out + ((long long)b * N + row0) * k;
  for (int e = tid; e < rows * k; e += KTM) {
    int r = e / k, s = e % k;
    out[e] = ranked[s * KTM + r];
  }
}

static size_t knn_simt_smem(int C, int k) {
  const size_t sel = (k == 20 || k == 40) ? (size_t)(k > 32 ? k : 32) * KTM : TopK::smem_floats(k);
  return sizeof(float) * ((size_t)C * KQLD + KCC * KCLD + KTN + sel);
}

template <int K>
static int knn_simt_launch(const float* x, int C, int N, int k, long long sb, long long sn, long long sc, int* idx,
                           dim3 grid, size_t smem, cudaStream_t stream) {
  SUG_TRY(ensure_dyn_smem((const void*)knn_simt_kernel<K>, smem));
  knn_simt_kernel<K><<<grid, KTM, smem, stream>>>(x, C, N, k, sb, sn, sc, idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

int knn_simt(const float* x, int B, int C, int N, int k, long long sb, long long sn, long long sc, int* idx,
             cudaStream_t stream) {
  size_t smem = knn_simt_smem(C, k);
  SUG_CHECK_ARG(smem <= 227 * 1024, "knn: C=%d k=%d needs %zu B of shared memory (> 227 KB)", C, k, smem);
  dim3 grid(cdiv(N, KTM), B);
  ProfScope ps(KC_KNN, 2.0 * B * (double)N * N * C, 4.0 * B * (double)N * (C + k), stream);
  if (k == 20) return knn_simt_launch<20>(x, C, N, k, sb, sn, sc, idx, grid, smem, stream);
  if (k == 40) return knn_simt_launch<40>(x, C, N, k, sb, sn, sc, idx, grid, smem, stream);
  return knn_simt_launch<0>(x, C, N, k, sb, sn, sc, idx, grid, smem, stream);
}

// ---- transposed graph -------------------------------------------------------------------------
// One CTA per cloud: counting sort of the N*k edges by destination.  The order inside a destination's
// list is the arrival order of the atomics (unspecified): sorting the lists would be quadratic in the
// in-degree, and feature-space kNN graphs have hubs with in-degrees in the thousands.
__global__ void __launch_bounds__(1024)
knn_reverse_kernel(const int* __restrict__ idx, int N, int k, int* __restrict__ rev_ptr, int* __restrict__ rev_edge) {
  extern __shared__ int sm[];
  int* deg = sm;          // [N]   in-degree, later the fill cursor
  int* ptr = sm + N;      // [N+1] exclusive scan
  __shared__ int wsum[32];
  __shared__ int carry;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int* ib = idx + (long long)b * N * k;
  int* eb = rev_edge + (long long)b * N * k;
  int* pb = rev_ptr + (long long)b * (N + 1);
  for (int j = tid; j < N; j += nt) deg[j] = 0;
  if (tid == 0) carry = 0;
  __syncthreads();
  const int E = N * k;
  for (int e = tid; e < E; e += nt) atomicAdd(&deg[ib[e]], 1);
  __syncthreads();
  // block-wide exclusive scan in chunks of blockDim
  for (int base = 0; base < N; base += nt) {
    int j = base + tid;
    int v = j < N ? deg[j] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      int w = tid < (nt >> 5) ? wsum[tid] : 0;
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (tid >= o) wi += t;
      }
      wsum[tid] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    int excl = carry + wsum[tid >> 5] + incl - v;
    if (j < N) ptr[j] = excl;
    __syncthreads();
    if (tid == nt - 1) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) ptr[N] = carry;
  __syncthreads();
  for (int j = tid; j <= N; j += nt) pb[j] = ptr[j];
  for (int j = tid; j < N; j += nt) deg[j] = 0;
  __syncthreads();
  for (int e = tid; e < E; e += nt) {
    int j = ib[e];
    int pos = atomicAdd(&deg[j], 1);
    int i = e / k, s = e - i * k;
    eb[ptr[j] + pos] = (i << 8) | s;
  }
}

}  // namespace sug

namespace sug {
bool knn_tc_supported(int C, int k, int N, long long sn, long long sc, const float* x);
size_t knn_tc_ws_bytes(int B, int C, int N);
int knn_tc(const float* x, int B, int C, int N, int k, long long ld, int* idx, void* ws, size_t ws_bytes,
           cudaStream_t stream);
}  // namespace sug

extern "C" size_t sug_knn_ws_bytes(int B, int C, int N, int k) {
  (void)k;
  return sug::knn_tc_ws_bytes(B, C, N);
}

extern "C" int sug_knn_f32(const float* x, int B, int C, int N, int k, int64_t sb, int64_t sn, int64_t sc,
                           int32_t* idx, void* ws, size_t ws_bytes, sug_stream_t stream) {
  SUG_CHECK_ARG(x && idx, "knn: null pointer");
  SUG_CHECK_ARG(B > 0 && C > 0 && N > 0, "knn: bad shape B=%d C=%d N=%d", B, C, N);
  SUG_CHECK_ARG(k > 0 && k <= N && k <= 128, "knn: k=%d must satisfy 1 <= k <= min(N=%d, 128)", k, N);
  // feature inputs: tcgen05 distance tiles; xyz (C = 3) and odd layouts: CUDA cores
  if (sb == (int64_t)N * sn && sug::knn_tc_supported(C, k, N, sn, sc, x))
    return sug::knn_tc(x, B, C, N, k, sn, idx, ws, ws_bytes, (cudaStream_t)stream);
  return sug::knn_simt(x, B, C, N, k, sb, sn, sc, idx, (cudaStream_t)stream);
}

extern "C" int sug_knn_reverse(const int32_t* idx, int B, int N, int k, int32_t* rev_ptr, int32_t* rev_edge,
                               sug_stream_t stream) {
  SUG_CHECK_ARG(idx && rev_ptr && rev_edge, "knn_reverse: null pointer");
  SUG_CHECK_ARG(B > 0 && N > 0 && k > 0 && k <= 255, "knn_reverse: bad shape B=%d N=%d k=%d", B, N, k);
  SUG_CHECK_ARG(N < (1 << 23), "knn_reverse: N=%d too large for the packed edge format", N);
  size_t smem = sizeof(int) * (2 * (size_t)N + 1);
  SUG_CHECK_ARG(smem <= 227 * 1024, "knn_reverse: N=%d needs %zu B of shared memory", N, smem);
  SUG_TRY(sug::ensure_dyn_smem((const void*)sug::knn_reverse_kernel, smem));
  sug::ProfScope ps(sug::KC_KNN_REV, 0, 4.0 * B * ((double)N * k * 2 + N + 1), (cudaStream_t)stream);
  sug::knn_reverse_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(idx, N, k, rev_ptr, rev_edge);
  SUG_LAUNCH_CHECK();
  return 0;
}
