// Per-row top-k selection state shared by the CUDA-core (knn.cu) and tcgen05 (knn_tc.cu) kNN kernels.
#pragma once
#include "common.cuh"

namespace sug {

constexpr int KTM = 128;         // query rows per CTA (one thread each)
constexpr int KTN = 64;          // candidates per tile
constexpr int KCC = 32;          // channels per staged candidate chunk
constexpr int KQLD = KTM + 4;    // padded leading dims (16B aligned rows, conflict-free transposed stores)
constexpr int KCLD = KTN + 4;
constexpr int KPEND = 16;        // pending-queue capacity per row

struct TopK {
  float* topv;  // [k][KTM]
  int* topi;    // [k][KTM]
  float* pv;    // [KPEND][KTM]
  int* pi;      // [KPEND][KTM]
  int k;
  float thr;
  int minpos;
  int cnt;

  __device__ __forceinline__ void init(int tid) {
    for (int s = 0; s < k; ++s) {
      topv[s * KTM + tid] = -INFINITY;
      topi[s * KTM + tid] = 0;
    }
    thr = -INFINITY;
    minpos = 0;
    cnt = 0;
  }
  // Predicated append; the caller guarantees cnt < KPEND on entry.
  __device__ __forceinline__ void offer(int tid, float key, int j) {
    if (key > thr) {
      pv[cnt * KTM + tid] = key;
      pi[cnt * KTM + tid] = j;
      ++cnt;
    }
  }
  __device__ __forceinline__ void drain(int tid) {
    for (int p = 0; p < cnt; ++p) {
      float v = pv[p * KTM + tid];
      if (v > thr) {
        topv[minpos * KTM + tid] = v;
        topi[minpos * KTM + tid] = pi[p * KTM + tid];
        float mn = INFINITY;
        int mp = 0;
        for (int s = 0; s < k; ++s) {
          float t = topv[s * KTM + tid];
          if (t < mn) { mn = t; mp = s; }
        }
        thr = mn;
        minpos = mp;
      }
    }
    cnt = 0;
  }
  // Sort the k survivors, best (largest key) first; ties -> smaller index first.
  __device__ __forceinline__ void sort_desc(int tid) {
    for (int s = 0; s < k - 1; ++s) {
      float bv = topv[s * KTM + tid];
      int bi = topi[s * KTM + tid];
      int bp = s;
      for (int t = s + 1; t < k; ++t) {
        float v = topv[t * KTM + tid];
        int i = topi[t * KTM + tid];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; bp = t; }
      }
      if (bp != s) {
        topv[bp * KTM + tid] = topv[s * KTM + tid];
        topi[bp * KTM + tid] = topi[s * KTM + tid];
        topv[s * KTM + tid] = bv;
        topi[s * KTM + tid] = bi;
      }
    }
  }
};


}  // namespace sug
