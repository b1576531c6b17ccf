// fp32-accurate tensor-core GEMM for sm_100a:  C[M,N] (+)= A * B^T (+ bias),  3xTF32 split.
//
//   x = hi + lo,  hi = top 19 bits of the fp32 container (what kind::tf32 reads), lo = x - hi (exact)
//   A B^T ~= lo(A) hi(B)^T + hi(A) lo(B)^T + hi(A) hi(B)^T        (error ~2^-21 |a||b| per product)
//
// Persistent, warp-specialised CTA (one per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the raw fp32 A / B k-blocks (128 B rows,
//               SWIZZLE_128B) into a multi-stage ring; out-of-bounds rows / k read as zero
//   warps 2-5   split: read each landed stage, write the `lo` tiles next to the raw ones
//   warp 1      one elected thread issues 3 tcgen05.mma.kind::tf32 per 8-wide k-step into a TMEM
//               accumulator (128 lanes x BN columns, double buffered), tcgen05.commit frees the stage
//   warps 6-9   epilogue: tcgen05.ld the finished accumulator, add bias, store / atomically add to C
// Both operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the latter
// serves the weight-gradient reductions dY^T X without materialising transposes.
#include "tc_common.cuh"

namespace sug {

using namespace tc;

constexpr int TBM = 128;  // UMMA M (TMEM lanes)
constexpr int TBK = 32;   // fp32 per k-block = one 128 B swizzle row
constexpr int TC_THREADS = 320;

enum { EPI_STORE = 0, EPI_ATOMIC = 1 };

// A_TS: the A operand is staged in TMEM (hi and lo written by the split warps with tcgen05.st; an
// MN-major tile is transposed on the way: thread m gathers its 32 k-values from the swizzled tile), so
// shared memory only holds the raw A tile (TMA destination) and B hi / lo.  A 128x128x8
// tf32 MMA with both operands in shared memory reads 8 KB per 64 cycles = the whole 128 B/clk of an SM;
// with three products per k-step the SS form is shared-memory-bandwidth bound, the TS form is not.
template <int BN, bool A_TS>
struct TcCfg {
  static constexpr int A_BYTES = TBM * TBK * 4;  // 16 KB
  static constexpr int B_BYTES = BN * TBK * 4;
  static constexpr int A_SMEM = A_TS ? A_BYTES : 2 * A_BYTES;
  static constexpr int STAGE_BYTES = A_SMEM + 2 * B_BYTES;
  static constexpr int STAGES = A_TS ? ((2 * BN + 4 * 64 <= 512 && 4 * STAGE_BYTES <= 192 * 1024) ? 4 : 3)
                                     : (192 * 1024) / STAGE_BYTES;
  static constexpr int EPI_BYTES = 4 * 2 * 4096;  // per epilogue warp: two 32x32 fp32 staging tiles
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int NEED_COLS = ACC_COLS + (A_TS ? STAGES * 64 : 0);
  static constexpr int TMEM_COLS = NEED_COLS <= 32 ? 32 : NEED_COLS <= 64 ? 64 : NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
  static_assert(NEED_COLS <= 512, "TMEM overflow");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*column partials*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared-memory plan");
};

struct TcArgs {
  float* C;
  const float* bias;
  long long ldc;
  int M, N, K;
  int tiles_m, tiles_n, ksplits, kb_per_split;
  int epi;
  int tma_store;  // epilogue through shared memory + cp.async.bulk.tensor (needs an aligned C)
  double* colsums;  // optional [2*N]: per-column sum and sum of squares of C (BatchNorm statistics), accumulated
                    // by the epilogue from the staged tiles -- the separate pass over C (col_stats) disappears
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, TcArgs p) {
  constexpr bool A_TS = true;  // A always goes through TMEM (MN-major tiles are transposed on the way)
  using Cfg = TcCfg<BN, A_TS>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment as an OFFSET from the __shared__ symbol: the pointer stays in the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_smem = smem + S * Cfg::STAGE_BYTES;  // 1024 B aligned (stage sizes are multiples of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* full = bars;            // [S]  TMA -> split
  uint64_t* ready = bars + S;       // [S]  split -> MMA
  uint64_t* empty = bars + 2 * S;   // [S]  MMA -> TMA
  uint64_t* tfull = bars + 3 * S;   // [2]  MMA -> epilogue
  uint64_t* tempty = bars + 3 * S + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);
  float* colpart = reinterpret_cast<float*>(epi_smem + Cfg::EPI_BYTES + 256);  // [4 warps][2][32] column partials

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n * p.ksplits;
  const int kb_total = (p.K + TBK - 1) / TBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
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
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + (size_t)s * Cfg::STAGE_BYTES; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // (the whole warp runs the loop; one elected lane issues -- see elect_one_sync)
    {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % p.tiles_n, mt = (t / p.tiles_n) % p.tiles_m, ks = t / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
          if (elect_one_sync()) {
            uint8_t* sp = stage_ptr(s);
            mbar_arrive_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
            if (!A_MN) {
              tma_load_2d(sp, &tmA, &full[s], kb * TBK, mt * TBM);
            } else {
#pragma unroll
              for (int mb = 0; mb < TBM / 32; ++mb)
                tma_load_2d(sp + mb * 4096, &tmA, &full[s], mt * TBM + mb * 32, kb * TBK);
            }
            uint8_t* bp = sp + Cfg::A_SMEM;
            if (!B_MN) {
              tma_load_2d(bp, &tmB, &full[s], kb * TBK, nt * BN);
            } else {
#pragma unroll
              for (int nb = 0; nb < BN / 32; ++nb)
                tma_load_2d(bp + nb * 4096, &tmB, &full[s], nt * BN + nb * 32, kb * TBK);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    {
      constexpr uint32_t idesc = idesc_tf32(TBM, BN, (A_MN && !A_TS) ? 1 : 0, B_MN ? 1 : 0);  // TMEM A is K-major
      uint32_t it = 0, tile_it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tile_it) {
        const int ks = t / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        const int ab = tile_it & 1;
        mbar_wait(&tempty[ab], ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + ab * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&ready[s], (it / S) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stage_ptr(s));
          const uint32_t a_lo = a_hi + Cfg::A_BYTES;
          const uint32_t b_hi = a_hi + Cfg::A_SMEM;
          const uint32_t b_lo = b_hi + Cfg::B_BYTES;
          const uint32_t ta_hi = tmem_base + Cfg::ACC_COLS + s * 64, ta_lo = ta_hi + 32;  // A_TS only
          // the descriptors of the four 8-wide k-steps differ only in the start-address field
          // (K-major: +32 B = +2; MN-major: +1024 B = +64)
          const uint64_t dbh0 = B_MN ? smem_desc_mnmajor(b_hi, 4096) : smem_desc_kmajor(b_hi);
          const uint64_t dbl0 = B_MN ? smem_desc_mnmajor(b_lo, 4096) : smem_desc_kmajor(b_lo);
          const uint64_t dah0 = smem_desc_mnmajor(a_hi, 4096), dal0 = smem_desc_mnmajor(a_lo, 4096);  // SS form only
          constexpr uint64_t bstep = B_MN ? 64 : 2;
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < TBK / 8; ++j) {
              const uint64_t dbh = dbh0 + bstep * j, dbl = dbl0 + bstep * j;
              const uint32_t acc0 = (kb > kb0 || j > 0) ? 1u : 0u;
              if (A_TS) {
                mma_tf32_ts(tacc, ta_lo + j * 8, dbh, idesc, acc0);
                mma_tf32_ts(tacc, ta_hi + j * 8, dbl, idesc, 1u);
                mma_tf32_ts(tacc, ta_hi + j * 8, dbh, idesc, 1u);
              } else {
                const uint64_t dah = dah0 + 64 * j, dal = dal0 + 64 * j;
                mma_tf32(tacc, dal, dbh, idesc, acc0);
                mma_tf32(tacc, dah, dbl, idesc, 1u);
                mma_tf32(tacc, dah, dbh, idesc, 1u);
              }
            }
            mma_commit(&empty[s]);
            if (kb == kb1 - 1) mma_commit(&tfull[ab]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ===================================== hi/lo split ========================================
    const int tix = threadIdx.x - 64;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int ks = t / (p.tiles_n * p.tiles_m);
      const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % S;
        mbar_wait(&full[s], (it / S) & 1);
        uint8_t* sp = stage_ptr(s);
        const float4* a_hi = reinterpret_cast<const float4*>(sp);
        const float4* b_hi = reinterpret_cast<const float4*>(sp + Cfg::A_SMEM);
        float4* b_lo = reinterpret_cast<float4*>(sp + Cfg::A_SMEM + Cfg::B_BYTES);
        if (A_TS) {
          // this thread owns A row r (== TMEM lane): un-swizzle its 32 k-values, write hi and lo to TMEM
          const int r = (warp & 3) * 32 + lane;
          float hi[32], lo[32];
          if (!A_MN) {  // K-major tile: the row is 128 contiguous (SWIZZLE_128B) bytes
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(sp + r * 128 + ((c ^ (r & 7)) << 4));
              hi[4 * c] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
            }
          } else {      // MN-major tile: four [32 k][32 m] blocks, 32 B chunks XOR-ed with (k % 4); lanes read
                        // consecutive m of one k-row, i.e. conflict-free
            const uint8_t* blk = sp + (r >> 5) * 4096;
            const int ml = r & 31;
#pragma unroll
            for (int kk = 0; kk < 32; ++kk)
              hi[kk] = *reinterpret_cast<const float*>(blk + kk * 128 + ((((ml >> 3) ^ (kk & 3)) << 5) | ((ml & 7) << 2)));
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) lo[q] = tf32_residual(hi[q]);
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + Cfg::ACC_COLS + s * 64;
          tmem_st32(ta, hi);
          tmem_st32(ta + 32, lo);
          tmem_st_wait();
        } else {
          float4* a_lo = reinterpret_cast<float4*>(sp + Cfg::A_BYTES);
#pragma unroll
          for (int i = 0; i < Cfg::A_BYTES / 16 / 128; ++i) {
            float4 v = a_hi[tix + i * 128];
            a_lo[tix + i * 128] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
          }
        }
#pragma unroll
        for (int i = 0; i < Cfg::B_BYTES / 16 / 128; ++i) {
          float4 v = b_hi[tix + i * 128];
          b_lo[tix + i * 128] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&ready[s]);
      }
    }
  } else {
    // ===================================== epilogue ===========================================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    uint32_t tile_it = 0;
    uint8_t* stg = epi_smem + (warp - 6) * 8192;  // this warp's two 32x32 staging tiles
    uint32_t chunk_it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tile_it) {
      const int nt = t % p.tiles_n, mt = (t / p.tiles_n) % p.tiles_m;
      // split-K: the partial sums are reduced into a zeroed C, the first k-range brings the bias along
      const bool add_bias = p.bias != nullptr && t < p.tiles_n * p.tiles_m;
      const int ab = tile_it & 1;
      mbar_wait(&tfull[ab], (tile_it >> 1) & 1);
      tc_fence_after();
      const int row = mt * TBM + lg * 32 + lane;
      if (p.tma_store) {
        // TMEM -> registers -> 128B-swizzled staging tile -> bulk tensor store (or reduce-add for
        // split-K): every global write is a full, coalesced 128 B row segment, tails are clipped by TMA
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c, ++chunk_it) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + ab * BN + c * 32, v);
          tmem_ld_wait();
          const int col0 = nt * BN + c * 32;
          if (add_bias) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (col0 + q < p.N) v[q] += __ldg(p.bias + col0 + q);
          }
          uint8_t* buf = stg + (chunk_it & 1) * 4096;
          if (lane == 0) tma_store_wait_read<1>();  // the store issued two chunks ago has read this buffer
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && col0 < p.N && mt * TBM + lg * 32 < p.M) {
            if (p.epi == EPI_ATOMIC) tma_reduce_add_2d(&tmC, buf, col0, mt * TBM + lg * 32);
            else tma_store_2d(&tmC, buf, col0, mt * TBM + lg * 32);
          }
          if (lane == 0) tma_store_commit();
          if (p.colsums != nullptr) {
            // lane c sums column col0 + c over this warp's (valid) rows of the staged tile: the 16 B chunk of
            // column quad q sits at (q ^ (row & 7)), so the 32 lanes of a row read 32 different banks.  The four
            // epilogue warps (the four 32-row groups of the tile) combine through 1 KB of shared memory and warp 0
            // issues the 64 fp64 atomics of the chunk: M/128 instead of M/32 atomics per column.
            const int nvalid = min(32, p.M - (mt * TBM + lg * 32));
            const int qq = lane >> 2, e4 = (lane & 3) << 2;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < nvalid; ++r) {
              const float val = *reinterpret_cast<const float*>(buf + r * 128 + ((qq ^ (r & 7)) << 4) + e4);
              s1 += val;
              s2 = fmaf(val, val, s2);
            }
            colpart[(warp - 6) * 64 + lane] = s1;
            colpart[(warp - 6) * 64 + 32 + lane] = s2;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 6) {
              const int col = col0 + lane;
              const float t1 = colpart[lane] + colpart[64 + lane] + colpart[128 + lane] + colpart[192 + lane];
              const float t2 = colpart[32 + lane] + colpart[96 + lane] + colpart[160 + lane] + colpart[224 + lane];
              if (col < p.N) {
                atomicAdd(p.colsums + col, (double)t1);
                atomicAdd(p.colsums + p.N + col, (double)t2);
              }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
        }
      } else {
        float* crow = p.C + (long long)row * p.ldc;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + ab * BN + c * 32, v);
          tmem_ld_wait();
          const int col0 = nt * BN + c * 32;
          if (row < p.M && col0 < p.N) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              if (col0 + q < p.N) {
                float o = v[q];
                if (add_bias) o += __ldg(p.bias + col0 + q);
                if (p.epi == EPI_ATOMIC) atomicAdd(crow + col0 + q, o);
                else crow[col0 + q] = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[ab]);
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const float* base, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows,
                 bool swizzle32b, uint32_t box_inner, int l2promo) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SUG_E_UNSUPPORTED; }
  SUG_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 4 == 0,
                "TMA operand needs a 16 B aligned base and a row stride that is a multiple of 4 floats (ld=%llu)",
                (unsigned long long)ld);
  // The encoder is a driver-API call: it needs the primary context bound to THIS thread, which a
  // fresh autograd worker thread does not have until its first runtime call.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    SUG_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstr[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_inner != 32 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                   : (swizzle32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                : (l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SUG_E_BADARG; }
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcArgs& args, int grid,
                     cudaStream_t stream) {
  using Cfg = TcCfg<BN, true>;
  SUG_TRY(ensure_dyn_smem((const void*)gemm_tc_kernel<BN, A_MN, B_MN>, Cfg::SMEM_BYTES));
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmC, args);
  SUG_LAUNCH_CHECK();
  return 0;
}

// a: K-major -> [M, K] row-major with stride lda; MN-major -> [K, M] row-major with stride lda (same for b / N).
int gemm_tc_f32(const float* a, int64_t lda, int a_mn, const float* b, int64_t ldb, int b_mn, const float* bias, float* c,
                int64_t ldc, int M, int N, int K, cudaStream_t stream) {
  return gemm_tc_stats_f32(a, lda, a_mn, b, ldb, b_mn, bias, c, ldc, M, N, K, nullptr, nullptr, stream);
}

// Same, optionally accumulating the column sums / sums of squares of C into colsums[2*N] (fp64, zeroed by the
// caller).  *fused tells whether the epilogue did it (it cannot with split-K or an unaligned C).
int gemm_tc_stats_f32(const float* a, int64_t lda, int a_mn, const float* b, int64_t ldb, int b_mn, const float* bias,
                      float* c, int64_t ldc, int M, int N, int K, double* colsums, bool* fused, cudaStream_t stream) {
  SUG_CHECK_ARG(M > 0 && N > 0 && K > 0 && a && b && c, "gemm_tc: bad problem M=%d N=%d K=%d", M, N, K);
  if (fused != nullptr) *fused = false;
  const int BN = (N <= 64) ? 64 : 128;
  CUtensorMap tmA, tmB, tmC;
  const bool c_tma = (reinterpret_cast<uintptr_t>(c) & 15) == 0 && ldc % 4 == 0;
  if (c_tma) SUG_TRY(make_tmap_2d(&tmC, c, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32));
  else tmC = CUtensorMap();
  if (!a_mn) SUG_TRY(make_tmap_2d(&tmA, a, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TBM));
  else SUG_TRY(make_tmap_2d(&tmA, a, (uint64_t)M, (uint64_t)K, (uint64_t)lda, TBK, true));
  if (!b_mn) SUG_TRY(make_tmap_2d(&tmB, b, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)BN));
  else SUG_TRY(make_tmap_2d(&tmB, b, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, TBK, true));
  TcArgs args;
  args.C = c; args.bias = bias; args.ldc = ldc; args.M = M; args.N = N; args.K = K;
  args.tiles_m = cdiv(M, TBM);
  args.tiles_n = cdiv(N, BN);
  const int kb_total = cdiv(K, TBK);
  const long long tiles = (long long)args.tiles_m * args.tiles_n;
  int splits = 1;
  if (tiles < num_sms() && kb_total >= 16) {
    splits = (int)min((long long)(kb_total / 8), (num_sms() + tiles - 1) / tiles);
    if (splits < 1) splits = 1;
  }
  args.kb_per_split = cdiv(kb_total, splits);
  args.ksplits = cdiv(kb_total, args.kb_per_split);
  args.epi = args.ksplits > 1 ? EPI_ATOMIC : EPI_STORE;
  args.tma_store = c_tma ? 1 : 0;
  args.colsums = (colsums != nullptr && args.epi == EPI_STORE && c_tma) ? colsums : nullptr;
  if (fused != nullptr) *fused = args.colsums != nullptr;
  if (args.epi == EPI_ATOMIC) {
    SUG_CUDA(cudaMemset2DAsync(c, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream));
  }
  const int grid = (int)min((long long)num_sms(), tiles * args.ksplits);
  ProfScope ps(KC_GEMM_TC, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
#define SUG_TC(BN_)                                                                     \
  do {                                                                                  \
    if (!a_mn && !b_mn) return launch_tc<BN_, false, false>(tmA, tmB, tmC, args, grid, stream); \
    if (!a_mn && b_mn) return launch_tc<BN_, false, true>(tmA, tmB, tmC, args, grid, stream);   \
    if (a_mn && !b_mn) return launch_tc<BN_, true, false>(tmA, tmB, tmC, args, grid, stream);   \
    return launch_tc<BN_, true, true>(tmA, tmB, tmC, args, grid, stream);                       \
  } while (0)
  if (BN == 64) SUG_TC(64);
  SUG_TC(128);
#undef SUG_TC
}

}  // namespace sug

extern "C" int sug_gemm_tc_f32(const float* a, int64_t lda, int a_mn_major, const float* b, int64_t ldb, int b_mn_major,
                               const float* bias, float* c, int64_t ldc, int M, int N, int K, sug_stream_t stream) {
  return sug::gemm_tc_f32(a, lda, a_mn_major, b, ldb, b_mn_major, bias, c, ldc, M, N, K, (cudaStream_t)stream);
}
