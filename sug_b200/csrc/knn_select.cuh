// Per-row top-k selection shared by the CUDA-core (knn.cu) and tcgen05 (knn_tc.cu) kNN kernels.
//
// One thread owns one query row.  The k best keys seen so far live in a binary MIN-heap in shared
// memory (column-major [slot][row] so a warp's accesses are conflict-free); the root is the row's
// current k-th best and doubles as the admission threshold.  Candidates arrive in register chunks
// of 32: a branch-free pass builds a bitmask of the keys that beat the threshold and stashes the
// chunk to shared memory; only set bits are then visited (a replace-root + sift-down, <= 4 levels
// for k = 20), so the cost per candidate is ~5 instructions plus ~60 per accepted candidate, and an
// accepted candidate tightens the threshold immediately.
#pragma once
#include "common.cuh"

namespace sug {

constexpr int KTM = 128;  // query rows per selection group (one thread each)

struct TopK {
  float* hv;     // [k][KTM]  heap values
  int* hi;       // [k][KTM]  heap payload (candidate index)
  float* stash;  // [32][KTM] the chunk being examined
  int k;
  float thr;     // == hv[0]: k-th best so far

  static __host__ __device__ size_t smem_floats(int k) { return 2 * (size_t)k * KTM + 32 * KTM; }

  __device__ __forceinline__ void bind(float* base, int k_) {
    k = k_;
    hv = base;
    hi = reinterpret_cast<int*>(base + (size_t)k_ * KTM);
    stash = base + 2 * (size_t)k_ * KTM;
  }
  __device__ __forceinline__ void init(int tid) {
    for (int s = 0; s < k; ++s) {
      hv[s * KTM + tid] = -INFINITY;
      hi[s * KTM + tid] = 0;
    }
    thr = -INFINITY;
  }
  // key > thr: replace the root and restore the heap
  __device__ __forceinline__ void insert(int tid, float key, int j) {
    int pos = 0;
    while (true) {
      const int l = 2 * pos + 1;
      if (l >= k) break;
      const int r = l + 1;
      const float vl = hv[l * KTM + tid];
      const float vr = r < k ? hv[r * KTM + tid] : INFINITY;
      const int c = vr < vl ? r : l;
      const float vc = fminf(vl, vr);
      if (!(vc < key)) break;
      hv[pos * KTM + tid] = vc;
      hi[pos * KTM + tid] = hi[c * KTM + tid];
      pos = c;
    }
    hv[pos * KTM + tid] = key;
    hi[pos * KTM + tid] = j;
    thr = hv[tid];
  }
  // keys[0..31] are candidates base .. base+31; `valid` masks out-of-range columns.
  __device__ __forceinline__ void consider32(int tid, const float (&keys)[32], uint32_t valid, int base) {
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      if (keys[q] > thr) mask |= 1u << q;
      stash[q * KTM + tid] = keys[q];
    }
    mask &= valid;
    while (mask) {
      const int q = __ffs(mask) - 1;
      mask &= mask - 1;
      const float key = stash[q * KTM + tid];
      if (key > thr) insert(tid, key, base + q);
    }
  }
  // Sort the k survivors, best (largest key) first; ties -> smaller index first.
  __device__ __forceinline__ void sort_desc(int tid) {
    for (int s = 0; s < k - 1; ++s) {
      float bv = hv[s * KTM + tid];
      int bi = hi[s * KTM + tid];
      int bp = s;
      for (int t = s + 1; t < k; ++t) {
        float v = hv[t * KTM + tid];
        int i = hi[t * KTM + tid];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; bp = t; }
      }
      if (bp != s) {
        hv[bp * KTM + tid] = hv[s * KTM + tid];
        hi[bp * KTM + tid] = hi[s * KTM + tid];
        hv[s * KTM + tid] = bv;
        hi[s * KTM + tid] = bi;
      }
    }
  }
};

// Register-resident variant for the common k (20, 40): the list is kept SORTED (best first) in
// registers with fully static indexing, so an accepted candidate costs ~5 ALU instructions per slot
// (max/min on the keys, one compare + two selects on the payload) with no shared-memory round trips
// on the dependent chain; the k-th best is v[K-1] and the final list needs no sort.
template <int K>
struct TopKReg {
  float v[K];
  int id[K];
  float* stash;  // [32][KTM]
  float thr;

  static __host__ __device__ size_t smem_floats() { return 32 * KTM; }

  __device__ __forceinline__ void bind(float* base) { stash = base; }
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < K; ++s) { v[s] = -INFINITY; id[s] = 0; }
    thr = -INFINITY;
  }
  __device__ __forceinline__ void insert(float key, int j) {
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const bool gt = key > v[s];  // strict: an equal key stays behind the earlier (smaller) index
      const float hi_ = fmaxf(v[s], key), lo_ = fminf(v[s], key);
      const int ni = gt ? j : id[s];
      j = gt ? id[s] : j;
      v[s] = hi_;
      key = lo_;
      id[s] = ni;
    }
    thr = v[K - 1];
  }
  __device__ __forceinline__ void consider32(int tid, const float (&keys)[32], uint32_t valid, int base) {
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      if (keys[q] > thr) mask |= 1u << q;
      stash[q * KTM + tid] = keys[q];
    }
    mask &= valid;
    while (mask) {
      const int q = __ffs(mask) - 1;
      mask &= mask - 1;
      const float key = stash[q * KTM + tid];
      if (key > thr) insert(key, base + q);
    }
  }
};

// ---- 32-input odd-even merge sort (Batcher), descending, on registers: 191 compare-exchanges ----
__device__ __forceinline__ void cex(float& a, float& b) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  a = hi;
  b = lo;
}
template <int LO, int N, int R>
__device__ __forceinline__ void oe_merge(float (&a)[32]) {
  constexpr int M = R * 2;
  if constexpr (M < N) {
    oe_merge<LO, N, M>(a);
    oe_merge<LO + R, N, M>(a);
#pragma unroll
    for (int i = LO + R; i + R < LO + N; i += M) cex(a[i], a[i + R]);
  } else {
    cex(a[LO], a[LO + R]);
  }
}
template <int LO, int N>
__device__ __forceinline__ void oe_sort(float (&a)[32]) {
  if constexpr (N > 1) {
    oe_sort<LO, N / 2>(a);
    oe_sort<LO + N / 2, N / 2>(a);
    oe_merge<LO, N, 1>(a);
  }
}

// Rare path: a (row, group) list is close to full.  Keep its k best (value desc, position asc),
// compact, and raise the group's threshold to the k-th kept value.
template <int K, int CAP>
__device__ __noinline__ void knn_prune(float* mv, unsigned short* mi, int r, int& cnt, float& T) {
  uint32_t keep0 = 0, keep1 = 0, keep2 = 0;
  float newT = INFINITY;
  for (int i = 0; i < cnt; ++i) {
    const float vi = mv[i * 128 + r];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) {
      const float vj = mv[j * 128 + r];
      rank += (vj > vi || (vj == vi && j < i)) ? 1 : 0;
    }
    if (rank < K) {
      if (i < 32) keep0 |= 1u << i;
      else if (i < 64) keep1 |= 1u << (i - 32);
      else keep2 |= 1u << (i - 64);
      newT = fminf(newT, vi);
    }
  }
  int w = 0;
  for (int i = 0; i < cnt; ++i) {
    const uint32_t m = i < 32 ? keep0 : (i < 64 ? keep1 : keep2);
    if ((m >> (i & 31)) & 1u) {
      mv[w * 128 + r] = mv[i * 128 + r];
      mi[w * 128 + r] = mi[i * 128 + r];
      ++w;
    }
  }
  cnt = w;
  T = fmaxf(T, newT);
}

}  // namespace sug
