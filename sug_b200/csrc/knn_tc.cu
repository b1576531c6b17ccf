// Fused pairwise distance + top-k on the tensor cores (reference: model/model_utils.py:178-185).
//
// For feature inputs (C >= 16) the -2 X X^T term is a dense contraction: each persistent CTA owns
// blocks of 128 query points of one cloud and streams that cloud's candidates in tiles of 128.
// Every (query block, candidate tile) pair is one 128 x 128 x C product issued as 3xTF32
// tcgen05.mma (fp32-accurate hi/lo split, see gemm_tc.cu) from TMA-loaded SWIZZLE_128B stages into
// a double-buffered TMEM accumulator.  The epilogue warps own one query row per thread: they read
// the dot products with tcgen05.ld, form the reference's key
//        D_ij = (-|x_i|^2 + 2 x_i.x_j) - |x_j|^2
// and run the same pending-queue top-k selection as the CUDA-core kernel while the next tile's MMAs
// are in flight.  The N x N matrix exists only tile by tile in TMEM.
#include <type_traits>

#include "knn_select.cuh"
#include "tc_common.cuh"

namespace sug {

using namespace tc;

constexpr int QBN = 128;                   // candidates per tile (UMMA N)
constexpr int QTILE_BYTES = 128 * 32 * 4;  // one [128 x 32] fp32 operand block
constexpr int QSTAGE_BYTES = 4 * QTILE_BYTES;  // A hi/lo + B hi/lo
constexpr int QSTAGES = 1;  // two CTAs share an SM (8 selection warps); TMEM is double buffered
constexpr int QTHREADS = 320;

__global__ void row_sqnorm_kernel(const float* __restrict__ x, long long ld, long long P, int C, float* __restrict__ xx) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= P) return;
  const float* xr = x + row * ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = __ldg(xr + c);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) xx[row] = s;
}

struct KnnTcArgs {
  const float* xx;  // [B*N] squared norms
  int* idx;         // [B,N,k]
  int B, N, C, k;
  int mtiles_per_cloud, ntiles;
};

// K > 0: compile-time k (register-resident sorted list); K == 0: any k (heap in shared memory).
template <int K>
__global__ void __launch_bounds__(QTHREADS, (K == 40 ? 1 : 2))
knn_tc_kernel(const __grid_constant__ CUtensorMap tmX, KnnTcArgs p) {
  constexpr int S = QSTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * QSTAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* ready = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tfull = bars + 3 * S;
  uint64_t* tempty = bars + 3 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);
  float* xxs = reinterpret_cast<float*>(smem + S * QSTAGE_BYTES + 256);  // [2][128] candidate norms
  float* sel = xxs + 2 * QBN;                                             // top-k state

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_m = p.B * p.mtiles_per_cloud;
  const int kbs = (p.C + 31) / 32;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x) {
        const int b = mt / p.mtiles_per_cloud, r0 = (mt % p.mtiles_per_cloud) * 128;
        const int grow = b * p.N + r0;
        for (int nt = 0; nt < p.ntiles; ++nt) {
          const int gcol = b * p.N + nt * QBN;
          for (int kb = 0; kb < kbs; ++kb, ++it) {
            const int s = it % S;
            mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
            uint8_t* sp = smem + (size_t)s * QSTAGE_BYTES;
            mbar_arrive_expect_tx(&full[s], 2 * QTILE_BYTES);
            tma_load_2d(sp, &tmX, &full[s], kb * 32, grow);
            tma_load_2d(sp + 2 * QTILE_BYTES, &tmX, &full[s], kb * 32, gcol);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, QBN, 0, 0);
      uint32_t it = 0, tile_it = 0;
      for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x) {
        for (int nt = 0; nt < p.ntiles; ++nt, ++tile_it) {
          const int ab = tile_it & 1;
          mbar_wait(&tempty[ab], ((tile_it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + ab * QBN;
          for (int kb = 0; kb < kbs; ++kb, ++it) {
            const int s = it % S;
            mbar_wait(&ready[s], (it / S) & 1);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + (size_t)s * QSTAGE_BYTES);
            const uint32_t a_lo = a_hi + QTILE_BYTES, b_hi = a_hi + 2 * QTILE_BYTES, b_lo = a_hi + 3 * QTILE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t dah = smem_desc_kmajor(a_hi + j * 32), dal = smem_desc_kmajor(a_lo + j * 32);
              const uint64_t dbh = smem_desc_kmajor(b_hi + j * 32), dbl = smem_desc_kmajor(b_lo + j * 32);
              mma_tf32(tacc, dal, dbh, idesc, (kb > 0 || j > 0) ? 1u : 0u);
              mma_tf32(tacc, dah, dbl, idesc, 1u);
              mma_tf32(tacc, dah, dbh, idesc, 1u);
            }
            mma_commit(&empty[s]);
          }
          mma_commit(&tfull[ab]);
        }
      }
    }
  } else if (warp < 6) {
    // ===================================== hi/lo split ========================================
    const int tix = threadIdx.x - 64;
    uint32_t it = 0;
    for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x) {
      for (int nt = 0; nt < p.ntiles; ++nt) {
        for (int kb = 0; kb < kbs; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&full[s], (it / S) & 1);
          uint8_t* sp = smem + (size_t)s * QSTAGE_BYTES;
          const float4* a_hi = reinterpret_cast<const float4*>(sp);
          float4* a_lo = reinterpret_cast<float4*>(sp + QTILE_BYTES);
          const float4* b_hi = reinterpret_cast<const float4*>(sp + 2 * QTILE_BYTES);
          float4* b_lo = reinterpret_cast<float4*>(sp + 3 * QTILE_BYTES);
#pragma unroll
          for (int i = 0; i < QTILE_BYTES / 16 / 128; ++i) {
            float4 v = a_hi[tix + i * 128];
            a_lo[tix + i * 128] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
            float4 w = b_hi[tix + i * 128];
            b_lo[tix + i * 128] = make_float4(tf32_residual(w.x), tf32_residual(w.y), tf32_residual(w.z), tf32_residual(w.w));
          }
          fence_proxy_async_smem();
          mbar_arrive(&ready[s]);
        }
      }
    }
  } else {
    // ===================================== selection epilogue =================================
    const int lg = warp & 3;
    const int et = threadIdx.x - 192;  // 0..127, used for cooperative loads
    const int r = lg * 32 + lane;      // query row inside the block == TMEM lane
    typename std::conditional<(K > 0), TopKReg<(K > 0 ? K : 1)>, TopK>::type tk;
    if constexpr (K > 0) tk.bind(sel);
    else tk.bind(sel, p.k);
    uint32_t tile_it = 0;
    for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x) {
      const int b = mt / p.mtiles_per_cloud, r0 = (mt % p.mtiles_per_cloud) * 128;
      const int row = r0 + r;
      const long long cbase = (long long)b * p.N;
      const float xxq = row < p.N ? __ldg(p.xx + cbase + row) : 0.f;
      if constexpr (K > 0) tk.init();
      else tk.init(r);
      for (int nt = 0; nt < p.ntiles; ++nt, ++tile_it) {
        const int ab = tile_it & 1;
        {
          const int cj = nt * QBN + et;
          xxs[ab * QBN + et] = cj < p.N ? __ldg(p.xx + cbase + cj) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tfull[ab], (tile_it >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < QBN / 32; ++c) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + ab * QBN + c * 32, v);
          tmem_ld_wait();
          const float4* xj = reinterpret_cast<const float4*>(xxs + ab * QBN + c * 32);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 n4 = xj[q4];
            v[4 * q4 + 0] = fmaf(2.f, v[4 * q4 + 0], -xxq) - n4.x;
            v[4 * q4 + 1] = fmaf(2.f, v[4 * q4 + 1], -xxq) - n4.y;
            v[4 * q4 + 2] = fmaf(2.f, v[4 * q4 + 2], -xxq) - n4.z;
            v[4 * q4 + 3] = fmaf(2.f, v[4 * q4 + 3], -xxq) - n4.w;
          }
          const int base = nt * QBN + c * 32;
          const int nvalid = p.N - base;
          const uint32_t valid = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
          tk.consider32(r, v, valid, base);
        }
        tc_fence_before();
        mbar_arrive(&tempty[ab]);
      }
      if constexpr (K > 0) {
        if (row < p.N) {
          int* o = p.idx + (cbase + row) * K;
#pragma unroll
          for (int s = 0; s < K; ++s) o[s] = tk.id[s];
        }
      } else {
        tk.sort_desc(r);
        if (row < p.N) {
          int* o = p.idx + (cbase + row) * p.k;
          for (int s = 0; s < p.k; ++s) o[s] = tk.hi[s * KTM + r];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static size_t knn_tc_smem(int k) {
  const size_t sel = (k == 20 || k == 40) ? 32 * KTM : TopK::smem_floats(k);
  return (size_t)QSTAGES * QSTAGE_BYTES + 1024 + 256 + sizeof(float) * (2 * QBN + sel);
}

template <int K>
static int knn_tc_launch(const CUtensorMap& tmX, const KnnTcArgs& a, int grid, size_t smem, cudaStream_t stream) {
  static size_t configured = 0;
  if (smem > configured) {
    SUG_CUDA(cudaFuncSetAttribute(knn_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  knn_tc_kernel<K><<<grid, QTHREADS, smem, stream>>>(tmX, a);
  SUG_LAUNCH_CHECK();
  return 0;
}

bool knn_tc_supported(int C, int k, long long sn, long long sc, const float* x) {
  return C >= 16 && sc == 1 && sn % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && knn_tc_smem(k) <= 227 * 1024;
}

size_t knn_tc_ws_bytes(int B, int N) { return align_up(sizeof(float) * (size_t)B * N, 256) + 256; }

// x point-major: rows b*N + n with stride ld (sb == N*ld).
int knn_tc(const float* x, int B, int C, int N, int k, long long ld, int* idx, void* ws, size_t ws_bytes,
           cudaStream_t stream) {
  Workspace W(ws, ws_bytes);
  float* xx = W.take<float>((size_t)B * N);
  if (!W.ok()) { set_error("knn: workspace too small (%zu B, need %zu)", ws_bytes, knn_tc_ws_bytes(B, N)); return SUG_E_WORKSPACE; }
  const long long P = (long long)B * N;
  {
    ProfScope ps(KC_MISC, 2.0 * P * C, 4.0 * P * (C + 1), stream);
    row_sqnorm_kernel<<<cdiv(P * 32, 256), 256, 0, stream>>>(x, ld, P, C, xx);
  }
  SUG_LAUNCH_CHECK();
  CUtensorMap tmX;
  SUG_TRY(make_tmap_2d(&tmX, x, (uint64_t)C, (uint64_t)P, (uint64_t)ld, 128));
  KnnTcArgs a;
  a.xx = xx; a.idx = idx; a.B = B; a.N = N; a.C = C; a.k = k;
  a.mtiles_per_cloud = cdiv(N, 128);
  a.ntiles = cdiv(N, QBN);
  const size_t smem = knn_tc_smem(k);
  const int grid = min((k == 40 ? 1 : 2) * num_sms(), B * a.mtiles_per_cloud);
  ProfScope ps(KC_KNN_TC, 2.0 * B * (double)N * N * C, 4.0 * B * (double)N * (C + k), stream);
  if (k == 20) return knn_tc_launch<20>(tmX, a, grid, smem, stream);
  if (k == 40) return knn_tc_launch<40>(tmX, a, grid, smem, stream);
  return knn_tc_launch<0>(tmX, a, grid, smem, stream);
}

}  // namespace sug
