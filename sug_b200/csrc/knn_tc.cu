// Fused pairwise distance + top-k on the tensor cores (reference: model/model_utils.py:178-185).
//
// For feature inputs (16 <= C <= 128, k = 20 or 40) the -2 X X^T term is a dense contraction.  A persistent
// CTA per SM owns blocks of 128 query points of one cloud.  The block's query slabs are loaded ONCE
// (TMA), split into tf32 hi / lo parts and parked in TMEM (tcgen05.st) for the whole block; the
// cloud's candidates then stream through a 3-stage TMA ring in tiles of 128 (only the raw candidate
// slab crosses L2 -> shared memory; its lo part is produced next to it), and every (tile, 32-channel
// slab) is consumed by 3 x 4 tcgen05.mma.kind::tf32 with the A operand read from TMEM (fp32-accurate
// hi/lo split, see gemm_tc.cu) into a double-buffered TMEM accumulator.  The N x N matrix exists only
// tile by tile in TMEM.
//
// Selection is what bounds this kernel (ALU pipe: a streaming sorted-list insert costs ~100 min/max/
// select per accepted candidate and ~5 k log-many candidates are accepted per row), so the candidates
// are swept TWICE -- the tensor pipe is nearly idle, recomputing the tile is free:
//   sweep 0: every thread keeps 32 running maxima (candidate j -> slot j mod 32), sorts them with a
//            32-input odd-even merge network and the k-th largest over the row's 64 slot maxima (two
//            threads per row, see below) is a lower bound T <= (k-th largest key);
//   sweep 1: candidates with key >= T (about k + a few per row) are appended to a per-row list in
//            shared memory with predicated stores -- no data-dependent loop, no divergence;
//   final:   rank of every listed candidate by counting, out[rank] = id for rank < k (sorted output).
// Measured and dropped (round 2, B = 64, N = 1024, C = 64, k = 20, all bit-exact): the threshold sweep over only half
// of the tiles (the k-th best of a subset is still a valid bound) -- the ~2k + 12 survivors per row make the
// rank-by-counting and the list prunes quadratically more expensive: 372 us against 168 us; candidate lo tiles
// precomputed per call and loaded by TMA (a two-party TMA -> MMA ring without the split hop): 172 us, no gain.
// More of the same series (all bit-exact, timed back to back with this version on one box, 181 us with the Python call
// around it): a third selection group (576 threads, 96 registers, 16 slot maxima per thread) -- only the sweeps shrink
// per thread, the per-block phases (sort, threshold exchange, rank) do not, and with as many groups as accumulators the
// MMA of a group's next tile waits for that group: 213 us; 8-byte {key, id} list entries (one STS.64 per candidate,
// 9 -> 7 instructions) with or without a predicated tail advance: 180 / 190 us; rank by compare + predicated add
// (3 -> 2 instructions per pair): 179 us; 32-candidate chunks without register spills: 189 us.  A clock64 breakdown
// of this kernel (profiles/r02_notes.md) shows why none of them pays: the selection warps spend their time in
// dependency stalls (0.2 instructions per cycle each, two of them per scheduler), not in issue slots.
// Two selection groups (4 warps each) alternate over the accumulator buffers, so a query row is served
// by two threads (one per group) which exchange T, list sizes and lists through shared memory.
// Key = (-|x_i|^2 + 2 x_i.x_j) - |x_j|^2, the reference's operation order.
#include <cfloat>
#include <type_traits>

#include "knn_select.cuh"
#include "tc_common.cuh"

namespace sug {

using namespace tc;

constexpr int QBN = 128;                       // candidates per tile (UMMA N)
constexpr int QTILE_BYTES = 128 * 32 * 4;      // one [128 x 32] fp32 operand block
constexpr int QSTAGE_BYTES = 2 * QTILE_BYTES;  // B hi + B lo
constexpr int QMAXKB = 4;                      // C <= 128
constexpr int QTHREADS = 448;                  // TMA, MMA, 4 split warps, 2 x 4 selection warps

// xx[row] = |x_row|^2
__global__ void knn_prep_kernel(const float* __restrict__ x, long long ld, long long P, int C, float* __restrict__ xx) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= P) return;
  const float* xr = x + row * ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = __ldg(xr + c);
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

// Shared-memory plan of knn_tc_kernel<K> (offsets from the 1 KB aligned base).
template <int K>
struct KnnSmem {
  static constexpr int CAP = K <= 20 ? 56 : 80;  // list capacity per (row, group): k + 16 (one half chunk) + slack
  static constexpr int S = K <= 20 ? 4 : 3;      // candidate ring stages
  static constexpr size_t ring = (size_t)S * QSTAGE_BYTES;
  static constexpr size_t bars = ring;                       // 256 B
  static constexpr size_t xxs = bars + 256;                  // [2 buffers][2 groups][128] f32 candidate norms
  static constexpr size_t tbuf = xxs + 4 * QBN * 4;          // [128] f32 thresholds
  static constexpr size_t cnts = tbuf + 128 * 4;             // [2][128] i32 list sizes
  static constexpr size_t vals = cnts + 2 * 128 * 4;         // [2][CAP][128] f32
  static constexpr size_t ids = vals + 2 * (size_t)CAP * 128 * 4;  // [2][CAP][128] u16
  static constexpr size_t total = ids + 2 * (size_t)CAP * 128 * 2 + 1024;
};

template <int K>
__global__ void __launch_bounds__(QTHREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmX, KnnTcArgs p) {
  using L = KnnSmem<K>;
  constexpr int CAP = L::CAP;
  constexpr int S = L::S;
  static_assert(CAP <= 96, "knn_prune keeps three 32-bit masks");
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment as an OFFSET from the __shared__ symbol: the pointer stays in the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_stage = smem;  // the query slabs borrow the (drained) candidate ring at the start of a block
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* full = bars;             // [S] TMA -> split
  uint64_t* ready = bars + S;        // [S] split -> MMA
  uint64_t* empty = bars + 2 * S;    // [S] MMA -> TMA
  uint64_t* tfull = bars + 3 * S;    // [3] MMA -> selection
  uint64_t* tempty = bars + 3 * S + 3;  // [3] selection -> MMA
  uint64_t* a_full = bars + 3 * S + 6;  // TMA(A) -> split
  uint64_t* a_ready = bars + 3 * S + 7; // split -> MMA, TMA   (A hi/lo are in TMEM, staging is free)
  uint64_t* a_free = bars + 3 * S + 8;  // MMA -> TMA          (all MMAs of the query block retired)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 9);
  float* xxs = reinterpret_cast<float*>(smem + L::xxs);
  float* tbuf = reinterpret_cast<float*>(smem + L::tbuf);
  int* cnts = reinterpret_cast<int*>(smem + L::cnts);
  float* vals = reinterpret_cast<float*>(smem + L::vals);
  unsigned short* ids = reinterpret_cast<unsigned short*>(smem + L::ids);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_m = p.B * p.mtiles_per_cloud;
  const int kbs = (p.C + 31) / 32;
  // TMEM: nacc accumulators of 128 columns, then 64 columns (hi|lo) per query slab.  A group is bound to the
  // tiles of one parity; with only two accumulators the MMA of its next tile cannot start before the group
  // has drained the current one (cycle = select + MMA).  Three accumulators (C <= 64: 3*128 + 2*64 = 512
  // columns) let the MMA run one tile further ahead, so a group finds its next tile ready.
  const uint32_t nacc = kbs <= 2 ? 3u : 2u;
  const uint32_t A_COL = nacc * QBN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 128);
    }
    mbar_init(a_full, 1);
    mbar_init(a_ready, 128);
    mbar_init(a_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // (the whole warp runs the loop; one elected lane issues -- see elect_one_sync)
    {
      uint32_t it = 0, mi = 0;
      for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x, ++mi) {
        const int b = mt / p.mtiles_per_cloud, r0 = (mt % p.mtiles_per_cloud) * 128;
        mbar_wait(a_free, (mi & 1) ^ 1);  // previous query block fully consumed: the ring is drained
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(a_full, kbs * QTILE_BYTES);
          for (int kb = 0; kb < kbs; ++kb) tma_load_2d(a_stage + kb * QTILE_BYTES, &tmX, a_full, kb * 32, b * p.N + r0);
        }
        __syncwarp();
        mbar_wait(a_ready, mi & 1);  // queries are in TMEM: the ring is free for candidates
        for (int sweep = 0; sweep < 2; ++sweep) {
          for (int nt = 0; nt < p.ntiles; ++nt) {
            const int gcol = b * p.N + nt * QBN;
            for (int kb = 0; kb < kbs; ++kb, ++it) {
              const int s = it % S;
              mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
              if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full[s], QTILE_BYTES);
                tma_load_2d(smem + (size_t)s * QSTAGE_BYTES, &tmX, &full[s], kb * 32, gcol);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    {
      constexpr uint32_t idesc = idesc_tf32(128, QBN, 0, 0);
      uint32_t it = 0, tile_it = 0, mi = 0;
      for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x, ++mi) {
        mbar_wait(a_ready, mi & 1);
        tc_fence_after();
        for (int t = 0; t < 2 * p.ntiles; ++t, ++tile_it) {
          const uint32_t ab = tile_it % nacc;
          mbar_wait(&tempty[ab], ((tile_it / nacc) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + ab * QBN;
          for (int kb = 0; kb < kbs; ++kb, ++it) {
            const int s = it % S;
            mbar_wait(&ready[s], (it / S) & 1);
            tc_fence_after();
            const uint32_t b_hi = smem_u32(smem + (size_t)s * QSTAGE_BYTES);
            // descriptors of the 8-wide k-steps differ only in the start address field (+32 B = +2)
            const uint64_t dbh0 = smem_desc_kmajor(b_hi), dbl0 = smem_desc_kmajor(b_hi + QTILE_BYTES);
            const uint32_t ta_hi = tmem_base + A_COL + kb * 64, ta_lo = ta_hi + 32;
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t dbh = dbh0 + 2 * j, dbl = dbl0 + 2 * j;
                mma_tf32_ts(tacc, ta_lo + j * 8, dbh, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                mma_tf32_ts(tacc, ta_hi + j * 8, dbl, idesc, 1u);
                mma_tf32_ts(tacc, ta_hi + j * 8, dbh, idesc, 1u);
              }
              mma_commit(&empty[s]);
              if (kb == kbs - 1) mma_commit(&tfull[ab]);
            }
            __syncwarp();
          }
        }
        if (elect_one_sync()) mma_commit(a_free);
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ========================= split: queries -> TMEM (hi, lo), candidates -> lo tile =================
    const int tix = threadIdx.x - 64;
    const int r = (warp & 3) * 32 + lane;  // query row == TMEM lane this thread may write
    uint32_t it = 0, mi = 0;
    for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x, ++mi) {
      mbar_wait(a_full, mi & 1);
      for (int kb = 0; kb < kbs; ++kb) {
        const uint8_t* sp = a_stage + kb * QTILE_BYTES;
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(sp + r * 128 + ((c ^ (r & 7)) << 4));
          hi[4 * c] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) lo[q] = tf32_residual(hi[q]);
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + A_COL + kb * 64;
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(a_ready);
      for (int t = 0; t < 2 * p.ntiles; ++t) {
        for (int kb = 0; kb < kbs; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&full[s], (it / S) & 1);
          const float4* b_hi = reinterpret_cast<const float4*>(smem + (size_t)s * QSTAGE_BYTES);
          float4* b_lo = reinterpret_cast<float4*>(smem + (size_t)s * QSTAGE_BYTES + QTILE_BYTES);
#pragma unroll
          for (int i = 0; i < QTILE_BYTES / 16 / 128; ++i) {
            const float4 w = b_hi[tix + i * 128];
            b_lo[tix + i * 128] = make_float4(tf32_residual(w.x), tf32_residual(w.y), tf32_residual(w.z), tf32_residual(w.w));
          }
          fence_proxy_async_smem();
          mbar_arrive(&ready[s]);
        }
      }
    }
  } else {
    // ===================================== selection ===========================================
    const int grp = warp >= 10 ? 1 : 0;  // warps 6-9: group 0, warps 10-13: group 1
    const int lg = warp & 3;             // TMEM lane quarter this warp may access
    const int r = lg * 32 + lane;        // query row inside the block == TMEM lane
    const int gt = (warp - (grp == 0 ? 6 : 10)) * 32 + lane;  // 0..127 inside the group
    float* gxx = xxs + grp * QBN;
    float* mv = vals + (size_t)grp * CAP * 128;            // this thread's list: mv[e * 128 + r]
    unsigned short* mi = ids + (size_t)grp * CAP * 128;
    const float* ov = vals + (size_t)(1 - grp) * CAP * 128;  // the row's other list
    float* xch = vals + (size_t)CAP * 128;                  // group 1's sorted maxima (its list is empty then)
    uint32_t tile_it = 0;  // accumulator tiles issued before the current query block
    for (int mt = blockIdx.x; mt < total_m; mt += gridDim.x, tile_it += 2 * p.ntiles) {
      const int b = mt / p.mtiles_per_cloud, r0 = (mt % p.mtiles_per_cloud) * 128;
      const int row = r0 + r;
      const long long cbase = (long long)b * p.N;
      const float xxq = row < p.N ? __ldg(p.xx + cbase + row) : 0.f;
      float T = -FLT_MAX;
      int cnt = 0;
      float gm[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) gm[q] = -INFINITY;

      // tiles t = 0 .. 2*ntiles-1 (sweep 0 then sweep 1); mine are those that land in accumulator `grp`
      auto load_norm = [&](int t) -> float {
        const int cj = (t < p.ntiles ? t : t - p.ntiles) * QBN + gt;
        return (t < 2 * p.ntiles && cj < p.N) ? __ldg(p.xx + cbase + cj) : 0.f;
      };
      int t = (int)((tile_it ^ (uint32_t)grp) & 1u);
      float nx = load_norm(t);
      int nbuf = 0;
      bool have_T = false;
#pragma unroll 1
      for (;;) {
        if (!have_T && t >= p.ntiles) {
          // ---- threshold: k-th largest of the row's 64 slot maxima (each is a real candidate) ----
          have_T = true;
          oe_sort<0, 32>(gm);
          constexpr int NB = K < 32 ? K : 32;
          if (grp == 1) {
#pragma unroll
            for (int s = 0; s < NB; ++s) xch[s * 128 + r] = gm[s];
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          if (grp == 0) {
            float kth = -INFINITY;  // max over i of min(a[i-1], b[K-i-1]): i values from my maxima, K-i from the other's
#pragma unroll
            for (int i = 0; i <= K; ++i) {
              const int ia = i - 1, ib = K - i - 1;
              if (ia >= 32 || ib >= NB) continue;
              const float av = ia < 0 ? INFINITY : gm[ia];
              const float bv = ib < 0 ? INFINITY : xch[ib * 128 + r];
              kth = fmaxf(kth, fminf(av, bv));
            }
            tbuf[r] = kth;
          }
          asm volatile("bar.sync 4, 256;" ::: "memory");
          T = fmaxf(tbuf[r], -FLT_MAX);
        }
        if (t >= 2 * p.ntiles) break;
        const bool sweep1 = t >= p.ntiles;
        const int nt = sweep1 ? t - p.ntiles : t;
        const uint32_t my_it = tile_it + (uint32_t)t;
        const uint32_t ab = my_it % nacc;
        float* gx = gxx + nbuf * (2 * QBN);  // double-buffered norms: one group barrier per tile
        gx[gt] = nx;
        nx = load_norm(t + 2);  // in flight while this tile is processed
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        mbar_wait(&tfull[ab], (my_it / nacc) & 1);
        tc_fence_after();
        const int nh = min(QBN / 16, (p.N - nt * QBN + 15) >> 4);  // 16-candidate half chunks in this tile
        const uint32_t tsrc = tmem_base + ((uint32_t)(lg * 32) << 16) + ab * QBN;

        // one half chunk: candidates base .. base+15; H = 0/1 -> running-maximum slots 0-15 / 16-31
        auto process = [&](float (&v)[16], int h, auto Hc) {
          constexpr int H = decltype(Hc)::value;
          const int base = nt * QBN + h * 16;
          const int nvalid = p.N - base;
          const float4* xj = reinterpret_cast<const float4*>(gx + h * 16);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 n4 = xj[q4];
            v[4 * q4 + 0] = fmaf(2.f, v[4 * q4 + 0], -xxq) - n4.x;
            v[4 * q4 + 1] = fmaf(2.f, v[4 * q4 + 1], -xxq) - n4.y;
            v[4 * q4 + 2] = fmaf(2.f, v[4 * q4 + 2], -xxq) - n4.z;
            v[4 * q4 + 3] = fmaf(2.f, v[4 * q4 + 3], -xxq) - n4.w;
          }
          if (nvalid < 16) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = q < nvalid ? v[q] : -INFINITY;
          }
          if (!sweep1) {
#pragma unroll
            for (int q = 0; q < 16; ++q) gm[H * 16 + q] = fmaxf(gm[H * 16 + q], v[q]);
          } else {
            if (cnt > CAP - 16) knn_prune<K, CAP>(mv, mi, r, cnt, T);
            // branch-free append: every candidate is stored at the list tail, the tail only advances past
            // accepted ones (mask = -1) -- no predicated in-place pointer updates behind the stores
            const uint32_t wv0 = smem_u32(mv + r), wi0 = smem_u32(mi + r);
            uint32_t wv = wv0 + cnt * 512, wi = wi0 + cnt * 256;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              asm volatile(
                  "{\n"
                  ".reg .s32 m;\n"
                  "st.shared.f32 [%0], %2;\n"
                  "st.shared.u16 [%1], %4;\n"
                  "set.ge.s32.f32 m, %2, %3;\n"
                  "mad.lo.s32 %0, m, -512, %0;\n"
                  "mad.lo.s32 %1, m, -256, %1;\n"
                  "}\n"
                  : "+r"(wv), "+r"(wi)
                  : "f"(v[q]), "f"(T), "h"((unsigned short)(base + q))
                  : "memory");
            }
            cnt = (int)((wv - wv0) >> 9);
          }
        };

        // TMEM reads are software-pipelined: the next half chunk is in flight while one is consumed
        float va[16], vb[16];
        tmem_ld16(tsrc, va);
#pragma unroll 1
        for (int h = 0; h < nh; h += 2) {
          tmem_ld_wait();
          if (h + 1 < nh) tmem_ld16(tsrc + (h + 1) * 16, vb);
          process(va, h, std::integral_constant<int, 0>{});
          if (h + 1 < nh) {
            tmem_ld_wait();
            if (h + 2 < nh) tmem_ld16(tsrc + (h + 2) * 16, va);
            process(vb, h + 1, std::integral_constant<int, 1>{});
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty[ab]);
        nbuf ^= 1;
        t += 2;
      }

      // ---- final: rank by counting over the row's two lists; (value desc, group, position) order ----
      cnts[grp * 128 + r] = cnt;
      asm volatile("bar.sync 3, 256;" ::: "memory");
      {
        const int m_own = cnt, m_oth = cnts[(1 - grp) * 128 + r];
        int* o = p.idx + (cbase + row) * K;
#pragma unroll 1
        for (int i0 = 0; i0 < m_own; i0 += 4) {
          float vi[4];
          int rk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            vi[u] = (i0 + u < m_own) ? mv[(i0 + u) * 128 + r] : INFINITY;
            rk[u] = 0;
          }
          int j = 0;
#pragma unroll 4
          for (; j < i0; ++j) {  // earlier positions of my list win ties
            const float vj = mv[j * 128 + r];
#pragma unroll
            for (int u = 0; u < 4; ++u) rk[u] += vj >= vi[u] ? 1 : 0;
          }
          for (; j < i0 + 4 && j < m_own; ++j) {
            const float vj = mv[j * 128 + r];
#pragma unroll
            for (int u = 0; u < 4; ++u) rk[u] += (vj > vi[u] || (vj == vi[u] && j < i0 + u)) ? 1 : 0;
          }
#pragma unroll 4
          for (; j < m_own; ++j) {
            const float vj = mv[j * 128 + r];
#pragma unroll
            for (int u = 0; u < 4; ++u) rk[u] += vj > vi[u] ? 1 : 0;
          }
          if (grp == 1) {  // group 0's entries win ties
#pragma unroll 4
            for (j = 0; j < m_oth; ++j) {
              const float vj = ov[j * 128 + r];
#pragma unroll
              for (int u = 0; u < 4; ++u) rk[u] += vj >= vi[u] ? 1 : 0;
            }
          } else {
#pragma unroll 4
            for (j = 0; j < m_oth; ++j) {
              const float vj = ov[j * 128 + r];
#pragma unroll
              for (int u = 0; u < 4; ++u) rk[u] += vj > vi[u] ? 1 : 0;
            }
          }
          if (row < p.N) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (i0 + u < m_own && rk[u] < K) o[rk[u]] = (int)mi[(i0 + u) * 128 + r];
          }
        }
      }
      asm volatile("bar.sync 4, 256;" ::: "memory");  // lists are reused by the next query block
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int K>
static int knn_tc_launch(const CUtensorMap& tmX, const KnnTcArgs& a, int grid, cudaStream_t stream) {
  const size_t smem = KnnSmem<K>::total;
  SUG_TRY(ensure_dyn_smem((const void*)knn_tc_kernel<K>, smem));
  knn_tc_kernel<K><<<grid, QTHREADS, smem, stream>>>(tmX, a);
  SUG_LAUNCH_CHECK();
  return 0;
}

bool knn_tc_supported(int C, int k, int N, long long sn, long long sc, const float* x) {
  static_assert(KnnSmem<20>::total <= 227 * 1024 && KnnSmem<40>::total <= 227 * 1024, "shared-memory plan");
  return (k == 20 || k == 40) && C >= 16 && C <= 32 * QMAXKB && C % 4 == 0 && N < 65536 && sc == 1 && sn % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

size_t knn_tc_ws_bytes(int B, int C, int N) {
  (void)C;
  return align_up(sizeof(float) * (size_t)B * N, 256) + 512;
}

// x point-major: rows b*N + n with stride ld (sb == N*ld).
int knn_tc(const float* x, int B, int C, int N, int k, long long ld, int* idx, void* ws, size_t ws_bytes,
           cudaStream_t stream) {
  Workspace W(ws, ws_bytes);
  float* xx = W.take<float>((size_t)B * N);
  if (!W.ok()) { set_error("knn: workspace too small (%zu B, need %zu)", ws_bytes, knn_tc_ws_bytes(B, C, N)); return SUG_E_WORKSPACE; }
  const long long P = (long long)B * N;
  {
    ProfScope ps(KC_MISC, 2.0 * P * C, 4.0 * P * (C + 1), stream);
    knn_prep_kernel<<<cdiv(P * 32, 256), 256, 0, stream>>>(x, ld, P, C, xx);
  }
  SUG_LAUNCH_CHECK();
  CUtensorMap tmX;
  SUG_TRY(make_tmap_2d(&tmX, x, (uint64_t)C, (uint64_t)P, (uint64_t)ld, 128));
  KnnTcArgs a;
  a.xx = xx; a.idx = idx; a.B = B; a.N = N; a.C = C; a.k = k;
  a.mtiles_per_cloud = cdiv(N, 128);
  a.ntiles = cdiv(N, QBN);
  const int grid = min(num_sms(), B * a.mtiles_per_cloud);
  // algorithmic work (one distance matrix); the kernel computes it twice (two sweeps), which is its own business
  ProfScope ps(KC_KNN_TC, 2.0 * B * (double)N * N * C, 4.0 * B * (double)N * (C + k), stream);
  if (k == 20) return knn_tc_launch<20>(tmX, a, grid, stream);
  return knn_tc_launch<40>(tmX, a, grid, stream);
}

}  // namespace sug
