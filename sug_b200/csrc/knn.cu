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
#include <cfloat>
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

// ---- xyz clouds (C = 3): two-sweep selection on CUDA cores ---------------------------------------------
// One thread owns one query point; the cloud's candidates are staged once per tile as (x, y, z, |x|^2) float4
// and read as warp-wide broadcasts.  The distance costs 5 FMA-pipe operations, so -- like the tcgen05 kernel
// (knn_tc.cu) -- the candidates are swept twice instead of maintaining a sorted list (whose ~100-instruction
// inserts diverge across the lanes of a warp):
//   sweep 0: 64 running slot maxima per thread (candidate j -> slot j mod 64): one FMNMX per
//            candidate; the k-th largest slot maximum T is a lower bound of the row's k-th best key;
//   sweep 1: keys are recomputed (bit-identical: same operations) and those >= T (~k + 10 per row) appended
//            to a per-row list in shared memory with predicated stores -- no data-dependent control flow;
//   final:   rank by counting, ties to the lower index; coalesced write of the [rows][k] block.
// Key = (-|x_i|^2 + 2 x_i.x_j) - |x_j|^2 with sequential fp32 FMAs, the reference's order (model_utils.py:179-181).

template <int K>
struct XyzPlan {
  static constexpr int NSLOT = 64;  // 64 slot maxima: ~k + 4 survivors expected for k = 20 (32 slots: k + 11, frequent prunes)
  static constexpr int CAP = K <= 20 ? 56 : 80;  // k + one 16-candidate block + slack (see knn_prune)
  // candidates per staged tile (a multiple of 64): 8 KB for k = 20, i.e. 50 KB per CTA with the lists = 4 CTAs
  // per SM; the tile area is reused for CAP * 128 rank bytes at the end
  static constexpr int XTILE = K <= 20 ? 512 : 640;
  static constexpr size_t tile = 0;                                  // float4 [XTILE]
  static constexpr size_t vals = tile + (size_t)XTILE * 16;          // [CAP][128] f32
  static constexpr size_t ids = vals + (size_t)CAP * 128 * 4;        // [CAP][128] u16
  static constexpr size_t total = ids + (size_t)CAP * 128 * 2;
  static_assert((size_t)K * 128 * 4 <= (size_t)CAP * 128 * 4, "ranked block reuses the value lists");
};

template <int K>
__global__ void __launch_bounds__(128, 4)
knn_xyz_kernel(const float* __restrict__ x, int N, long long sb, long long sn, long long sc, int* __restrict__ idx_out) {
  using L = XyzPlan<K>;
  constexpr int CAP = L::CAP;
  constexpr int NSLOT = L::NSLOT;
  constexpr int XTILE = L::XTILE;
  extern __shared__ __align__(16) uint8_t xsm[];
  float4* tile = reinterpret_cast<float4*>(xsm + L::tile);
  float* mv = reinterpret_cast<float*>(xsm + L::vals);
  unsigned short* mi = reinterpret_cast<unsigned short*>(xsm + L::ids);
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * 128;
  const int row = row0 + tid;
  const float* xb = x + (long long)b * sb;
  float q0 = 0.f, q1 = 0.f, q2 = 0.f;
  if (row < N) {
    q0 = __ldg(xb + (long long)row * sn);
    q1 = __ldg(xb + (long long)row * sn + sc);
    q2 = __ldg(xb + (long long)row * sn + 2 * sc);
  }
  const float xxq = fmaf(q2, q2, fmaf(q1, q1, q0 * q0));
  float gm[NSLOT];
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) gm[q] = -INFINITY;
  float T = -FLT_MAX;
  int cnt = 0;

  auto key_of = [&](const float4 c) {
    const float dot = fmaf(q2, c.z, fmaf(q1, c.y, q0 * c.x));
    return fmaf(2.f, dot, -xxq) - c.w;
  };

#pragma unroll 1
  for (int sweep = 0; sweep < 2; ++sweep) {
#pragma unroll 1
    for (int t0 = 0; t0 < N; t0 += XTILE) {
      __syncthreads();  // the previous tile is fully consumed
      for (int e = tid; e < XTILE; e += 128) {
        const int j = t0 + e;
        float4 c = make_float4(0.f, 0.f, 0.f, INFINITY);  // out of range: key = -inf
        if (j < N) {
          c.x = __ldg(xb + (long long)j * sn);
          c.y = __ldg(xb + (long long)j * sn + sc);
          c.z = __ldg(xb + (long long)j * sn + 2 * sc);
          c.w = fmaf(c.z, c.z, fmaf(c.y, c.y, c.x * c.x));
        }
        tile[e] = c;
      }
      __syncthreads();
      const int nblk = (min(XTILE, N - t0) + NSLOT - 1) / NSLOT;  // whole NSLOT-blocks (padding reads as -inf)
      if (sweep == 0) {
#pragma unroll 1
        for (int bk = 0; bk < nblk; ++bk) {
#pragma unroll
          for (int q = 0; q < NSLOT; ++q) gm[q] = fmaxf(gm[q], key_of(tile[bk * NSLOT + q]));
        }
      } else {
        const int n16 = (min(XTILE, N - t0) + 15) >> 4;
#pragma unroll 1
        for (int hb = 0; hb < n16; ++hb) {
          if (cnt > CAP - 16) knn_prune<K, CAP>(mv, mi, tid, cnt, T);
          // predicated append: ~2 % of the candidates pass, so about half of the warp-wide store instructions have
          // no active lane and cost no shared-memory wavefront
          const uint32_t wv0 = (uint32_t)__cvta_generic_to_shared(mv + tid), wi0 = (uint32_t)__cvta_generic_to_shared(mi + tid);
          uint32_t wv = wv0 + cnt * 512, wi = wi0 + cnt * 256;
          const int base = t0 + hb * 16;
          float v16[16];  // all keys first (independent loads and FMA chains), then the serial tail updates
#pragma unroll
          for (int q = 0; q < 16; ++q) v16[q] = key_of(tile[hb * 16 + q]);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float v = v16[q];
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ge.f32 p, %2, %3;\n"
                "@p st.shared.f32 [%0], %2;\n"
                "@p st.shared.u16 [%1], %4;\n"
                "@p add.s32 %0, %0, 512;\n"
                "@p add.s32 %1, %1, 256;\n"
                "}\n"
                : "+r"(wv), "+r"(wi)
                : "f"(v), "f"(T), "h"((unsigned short)(base + q))
                : "memory");
          }
          cnt = (int)((wv - wv0) >> 9);
        }
      }
    }
    if (sweep == 0) {
      // T = K-th largest slot maximum (every slot maximum is a distinct real candidate)
      if constexpr (NSLOT == 32) {
        oe_sort<0, 32>(gm);
        T = fmaxf(gm[K - 1], -FLT_MAX);
      } else {
        float (&ga)[32] = *reinterpret_cast<float (*)[32]>(&gm[0]);
        float (&gb)[32] = *reinterpret_cast<float (*)[32]>(&gm[32]);
        oe_sort<0, 32>(ga);
        oe_sort<0, 32>(gb);
        float kth = -INFINITY;  // K-th largest of the union of two sorted lists: max_i min(a[i-1], b[K-i-1])
#pragma unroll
        for (int i = 0; i <= K; ++i) {
          const int ia = i - 1, ib = K - i - 1;
          if (ia >= 32 || ib >= 32) continue;
          const float av = ia < 0 ? INFINITY : ga[ia];
          const float bv = ib < 0 ? INFINITY : gb[ib];
          kth = fmaxf(kth, fminf(av, bv));
        }
        T = fmaxf(kth, -FLT_MAX);
      }
    }
  }

  // ---- rank by counting: (value desc, position asc); positions are in candidate order -------------------
  __syncthreads();  // all sweeps done: the candidate tile is free
  // phase A: the rank of every list entry, as a byte, into the (now free) tile area
  unsigned char* rkb = reinterpret_cast<unsigned char*>(xsm + L::tile);
  static_assert((size_t)CAP * 128 <= (size_t)XTILE * 16, "rank bytes reuse the candidate tile");
#pragma unroll 1
  for (int i0 = 0; i0 < cnt; i0 += 4) {
    float vi[4];
    int rk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vi[u] = (i0 + u < cnt) ? mv[(i0 + u) * 128 + tid] : INFINITY;
      rk[u] = 0;
    }
    int j = 0;
#pragma unroll 4
    for (; j < i0; ++j) {  // earlier positions win ties
      const float vj = mv[j * 128 + tid];
#pragma unroll
      for (int u = 0; u < 4; ++u) rk[u] += vj >= vi[u] ? 1 : 0;
    }
    for (; j < i0 + 4 && j < cnt; ++j) {
      const float vj = mv[j * 128 + tid];
#pragma unroll
      for (int u = 0; u < 4; ++u) rk[u] += (vj > vi[u] || (vj == vi[u] && j < i0 + u)) ? 1 : 0;
    }
#pragma unroll 4
    for (; j < cnt; ++j) {
      const float vj = mv[j * 128 + tid];
#pragma unroll
      for (int u = 0; u < 4; ++u) rk[u] += vj > vi[u] ? 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < cnt) rkb[(i0 + u) * 128 + tid] = (unsigned char)min(rk[u], 255);
  }
  __syncthreads();  // every thread is done reading the value lists
  int* rout = reinterpret_cast<int*>(xsm + L::vals);  // [K][128] ranked indices
  for (int i = 0; i < cnt; ++i) {
    const int rk = rkb[i * 128 + tid];
    if (rk < K) rout[rk * 128 + tid] = (int)mi[i * 128 + tid];
  }
  __syncthreads();
  const int rows = min(128, N - row0);
  int* out = idx_out + ((long long)b * N + row0) * K;
  for (int e = tid; e < rows * K; e += 128) {
    const int r = e / K, s = e - r * K;
    out[e] = rout[s * 128 + r];
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

template <int K>
static int knn_xyz_launch(const float* x, int B, int N, long long sb, long long sn, long long sc, int* idx,
                          cudaStream_t stream) {
  const size_t smem = XyzPlan<K>::total;
  SUG_TRY(ensure_dyn_smem((const void*)knn_xyz_kernel<K>, smem));
  ProfScope ps(KC_KNN, 2.0 * B * (double)N * N * 3, 4.0 * B * (double)N * (3 + K), stream);
  knn_xyz_kernel<K><<<dim3(cdiv(N, 128), B), 128, smem, stream>>>(x, N, sb, sn, sc, idx);
  SUG_LAUNCH_CHECK();
  return 0;
}

int knn_simt(const float* x, int B, int C, int N, int k, long long sb, long long sn, long long sc, int* idx,
             cudaStream_t stream) {
  if (C == 3 && N < 65536 && k == 20) return knn_xyz_launch<20>(x, B, N, sb, sn, sc, idx, stream);
  // (k = 40 stays on the sorted-list kernel below: 64 slot maxima leave ~k + 22 survivors per row, which
  //  overflows the lists too often; measured 920 us against 320 us)
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
