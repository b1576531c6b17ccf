// Fused pairwise distance + top-k on the tensor cores (reference: model/model_utils.py:178-185).
//
// For feature inputs (C >= 16) the -2 X X^T term is a dense contraction.  A small pre-pass writes the
// squared norms and the tf32 residuals lo(X) = X - hi(X) once per call; the main kernel is persistent:
// each CTA owns blocks of 128 query points of one cloud and streams that cloud's candidates in tiles
// of 128.  Every (query block, candidate tile, 32-channel slab) is one TMA stage {A_hi, A_lo, B_hi,
// B_lo} (SWIZZLE_128B) consumed by 3 x 4 tcgen05.mma.kind::tf32 (fp32-accurate hi/lo split, see
// gemm_tc.cu) into a double-buffered TMEM accumulator.  Selection is the expensive part (a sorted
// register list per query row, ~100 instructions per accepted candidate, latency bound), so BOTH
// remaining warp groups do it: group g takes the tiles that land in TMEM buffer g, keeps its own
// top-k per row straight out of tcgen05.ld, and the two lists are merged at the end of the query
// block.  Two CTAs share an SM (16 selection warps).  The N x N matrix exists only tile by tile in
// TMEM.  Key = (-|x_i|^2 + 2 x_i.x_j) - |x_j|^2, the reference's operation order.
#include <type_traits>

#include "knn_select.cuh"
#include "tc_common.cuh"

namespace sug {

using namespace tc;

constexpr int QBN = 128;                       // candidates per tile (UMMA N)
constexpr int QTILE_BYTES = 128 * 32 * 4;      // one [128 x 32] fp32 operand block
constexpr int QSTAGE_BYTES = 4 * QTILE_BYTES;  // A hi/lo + B hi/lo
constexpr int QTHREADS = 320;

// xx[row] = |x_row|^2 ; lo[row][c] = x - (x with the low 13 mantissa bits cleared)
__global__ void knn_prep_kernel(const float* __restrict__ x, long long ld, long long P, int C, float* __restrict__ xx,
                                float* __restrict__ lo) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= P) return;
  const float* xr = x + row * ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = __ldg(xr + c);
    s = fmaf(v, v, s);
    lo[row * C + c] = tf32_residual(v);
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

// K > 0: compile-time k (register-resident sorted lists, two selection groups);
// K == 0: any k (heap in shared memory, one selection group).
template <int K>
__global__ void __launch_bounds__(QTHREADS, (K == 20 ? 2 : 1))
knn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmL, KnnTcArgs p) {
  constexpr int NG = K > 0 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QSTAGE_BYTES);
  uint64_t* full = bars;        // TMA -> MMA
  uint64_t* empty = bars + 1;   // MMA -> TMA
  uint64_t* tfull = bars + 2;   // [2] MMA -> selection
  uint64_t* tempty = bars + 4;  // [2] selection -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* xxs = reinterpret_cast<float*>(smem + QSTAGE_BYTES + 256);  // [2 groups][128] candidate norms
  float* sel = xxs + 2 * QBN;  // group g: stash [32][128] at sel + g*32*128 (K>0) | heap+stash (K==0)
  // merge buffer: K == 20 aliases group 1's stash (20 value rows + 10 rows of packed 16-bit ids);
  // K == 40 has its own region behind the two stashes
  float* mrg = (K == 20) ? (sel + 32 * KTM) : (sel + 2 * 32 * KTM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_m = p.B * p.mtiles_per_cloud;
  const int kbs = (p.C + 31) / 32;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmL);
    mbar_init(full, 1);
    mbar_init(empty, 1);
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
            mbar_wait(empty, (it & 1) ^ 1);
            mbar_arrive_expect_tx(full, QSTAGE_BYTES);
            tma_load_2d(smem, &tmX, full, kb * 32, grow);
            tma_load_2d(smem + QTILE_BYTES, &tmL, full, kb * 32, grow);
            tma_load_2d(smem + 2 * QTILE_BYTES, &tmX, full, kb * 32, gcol);
            tma_load_2d(smem + 3 * QTILE_BYTES, &tmL, full, kb * 32, gcol);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, QBN, 0, 0);
      uint32_t it = 0, tile_it = 0;
      const uint32_t a_hi = smem_u32(smem);
      const uint32_t a_lo = a_hi + QTILE_BYTES, b_hi = a_hi + 2 * QTILE_BYTES, b_lo = a_hi + 3 * QTILE_BYTES;
      for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x) {
        for (int nt = 0; nt < p.ntiles; ++nt, ++tile_it) {
          const int ab = tile_it & 1;
          mbar_wait(&tempty[ab], ((tile_it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + ab * QBN;
          for (int kb = 0; kb < kbs; ++kb, ++it) {
            mbar_wait(full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t dah = smem_desc_kmajor(a_hi + j * 32), dal = smem_desc_kmajor(a_lo + j * 32);
              const uint64_t dbh = smem_desc_kmajor(b_hi + j * 32), dbl = smem_desc_kmajor(b_lo + j * 32);
              mma_tf32(tacc, dal, dbh, idesc, (kb > 0 || j > 0) ? 1u : 0u);
              mma_tf32(tacc, dah, dbl, idesc, 1u);
              mma_tf32(tacc, dah, dbh, idesc, 1u);
            }
            mma_commit(empty);
          }
          mma_commit(&tfull[ab]);
        }
      }
    }
  } else {
    // ===================================== selection ===========================================
    const int grp = warp >= 6 ? 0 : 1;  // warps 6-9: group 0, warps 2-5: group 1
    if (grp < NG) {
      const int lg = warp & 3;            // TMEM lane quarter this warp may access
      const int r = lg * 32 + lane;       // query row inside the block == TMEM lane
      const int gt = (warp - (grp == 0 ? 6 : 2)) * 32 + lane;  // 0..127 inside the group
      float* gxx = xxs + grp * QBN;
      typename std::conditional<(K > 0), TopKReg<(K > 0 ? K : 1)>, TopK>::type tk;
      if constexpr (K > 0) tk.bind(sel + grp * 32 * KTM);
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
          if (NG == 2 && ab != grp) continue;
          {
            const int cj = nt * QBN + gt;
            gxx[gt] = cj < p.N ? __ldg(p.xx + cbase + cj) : 0.f;
          }
          if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          mbar_wait(&tfull[ab], (tile_it >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < QBN / 32; ++c) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + ab * QBN + c * 32, v);
            tmem_ld_wait();
            const float4* xj = reinterpret_cast<const float4*>(gxx + c * 32);
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
          // the group's norms buffer is rewritten two tiles later, after the group barrier above
          if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        if constexpr (K > 0) {
          // ---- merge the two groups' lists (same rows, disjoint candidate tiles) ----
          unsigned short* mid = reinterpret_cast<unsigned short*>(mrg + K * KTM);  // [K][128] 16-bit ids
          if (grp == 1) {
#pragma unroll
            for (int s = 0; s < K; ++s) {
              mrg[s * KTM + r] = tk.v[s];
              mid[s * KTM + r] = (unsigned short)tk.id[s];
            }
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          if (grp == 0) {
#pragma unroll 1
            for (int s = 0; s < K; ++s) {
              const float key = mrg[s * KTM + r];
              if (!(key > tk.thr)) break;  // group 1's list is sorted: nothing further can enter
              tk.insert(key, (int)mid[s * KTM + r]);
            }
            if (row < p.N) {
              int* o = p.idx + (cbase + row) * K;
#pragma unroll
              for (int s = 0; s < K; ++s) o[s] = tk.id[s];
            }
          }
          asm volatile("bar.sync 4, 256;" ::: "memory");
        } else {
          tk.sort_desc(r);
          if (row < p.N) {
            int* o = p.idx + (cbase + row) * p.k;
            for (int s = 0; s < p.k; ++s) o[s] = tk.hi[s * KTM + r];
          }
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
  size_t sel;
  if (k == 20) sel = 2 * 32 * KTM;                       // two stashes (the merge aliases the second)
  else if (k == 40) sel = 2 * 32 * KTM + 40 * KTM * 2;   // + merge values and 16-bit ids (rounded up)
  else sel = TopK::smem_floats(k);
  return (size_t)QSTAGE_BYTES + 1024 + 256 + sizeof(float) * (2 * QBN + sel);
}

template <int K>
static int knn_tc_launch(const CUtensorMap& tmX, const CUtensorMap& tmL, const KnnTcArgs& a, int grid, size_t smem,
                         cudaStream_t stream) {
  static size_t configured = 0;
  if (smem > configured) {
    SUG_CUDA(cudaFuncSetAttribute(knn_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  knn_tc_kernel<K><<<grid, QTHREADS, smem, stream>>>(tmX, tmL, a);
  SUG_LAUNCH_CHECK();
  return 0;
}

bool knn_tc_supported(int C, int k, int N, long long sn, long long sc, const float* x) {
  return C >= 16 && C % 4 == 0 && N < 65536 && sc == 1 && sn % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
         knn_tc_smem(k) <= 227 * 1024;
}

size_t knn_tc_ws_bytes(int B, int C, int N) {
  return align_up(sizeof(float) * (size_t)B * N, 256) + align_up(sizeof(float) * (size_t)B * N * C, 256) + 512;
}

// x point-major: rows b*N + n with stride ld (sb == N*ld).
int knn_tc(const float* x, int B, int C, int N, int k, long long ld, int* idx, void* ws, size_t ws_bytes,
           cudaStream_t stream) {
  Workspace W(ws, ws_bytes);
  float* xx = W.take<float>((size_t)B * N);
  float* lo = W.take<float>((size_t)B * N * C);
  if (!W.ok()) { set_error("knn: workspace too small (%zu B, need %zu)", ws_bytes, knn_tc_ws_bytes(B, C, N)); return SUG_E_WORKSPACE; }
  const long long P = (long long)B * N;
  {
    ProfScope ps(KC_MISC, 2.0 * P * C, 4.0 * P * (2.0 * C + 1), stream);
    knn_prep_kernel<<<cdiv(P * 32, 256), 256, 0, stream>>>(x, ld, P, C, xx, lo);
  }
  SUG_LAUNCH_CHECK();
  CUtensorMap tmX, tmL;
  SUG_TRY(make_tmap_2d(&tmX, x, (uint64_t)C, (uint64_t)P, (uint64_t)ld, 128));
  SUG_TRY(make_tmap_2d(&tmL, lo, (uint64_t)C, (uint64_t)P, (uint64_t)C, 128));
  KnnTcArgs a;
  a.xx = xx; a.idx = idx; a.B = B; a.N = N; a.C = C; a.k = k;
  a.mtiles_per_cloud = cdiv(N, 128);
  a.ntiles = cdiv(N, QBN);
  const size_t smem = knn_tc_smem(k);
  const int grid = min((k == 20 ? 2 : 1) * num_sms(), B * a.mtiles_per_cloud);
  ProfScope ps(KC_KNN_TC, 2.0 * B * (double)N * N * C, 4.0 * B * (double)N * (C + k), stream);
  if (k == 20) return knn_tc_launch<20>(tmX, tmL, a, grid, smem, stream);
  if (k == 40) return knn_tc_launch<40>(tmX, tmL, a, grid, smem, stream);
  return knn_tc_launch<0>(tmX, tmL, a, grid, smem, stream);
}

}  // namespace sug
