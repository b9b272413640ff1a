// numpy's summation orders on device, so per-region means match the reference's `.mean()` bit for bit.
//
// numpy adds a contiguous 1-D array with pairwise_sum (numpy/core/src/umath/loops_utils.h.src): n < 8 sequential from 0;
// n <= 128: eight strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail in order;
// otherwise split at n2 = n/2 rounded down to a multiple of 8 and add the two halves.  LiDAL.py:97-98 / ReDAL.py:78 take
// `.mean()` of fancy-indexed (hence contiguous) float32 / float64 vectors, so this is the order they see.
#pragma once
#include "common.cuh"

namespace lb {

// numpy float32 pairwise sum over n <= 32 lane-resident values (lane j holds a[j]); result valid on lane 0.
__device__ __forceinline__ float np_pairwise_sum32(float a, int n, int lane) {
  const unsigned full = 0xffffffffu;
  if (n < 8) {
    float res = 0.f;
    for (int j = 0; j < n; ++j) res = __fadd_rn(res, __shfl_sync(full, a, j));
    return res;
  }
  float r = a;                                    // lanes 0..7 hold r[j]
  const int body = n - (n % 8);
  for (int i = 8; i < body; i += 8) {
    float t = __shfl_sync(full, a, (lane & 7) + i);
    r = __fadd_rn(r, t);
  }
  float s1 = __fadd_rn(r, __shfl_down_sync(full, r, 1));      // valid on even lanes: r[j] + r[j+1]
  float s2 = __fadd_rn(s1, __shfl_down_sync(full, s1, 2));    // valid on lanes 0,4
  float res = __fadd_rn(s2, __shfl_down_sync(full, s2, 4));   // valid on lane 0
  for (int i = body; i < n; ++i) res = __fadd_rn(res, __shfl_sync(full, a, i));
  return res;
}

constexpr int NP_BLOCK_THREADS = 256;
constexpr int NP_MAX_LEAVES = 2048;       // leaves hold 65..128 elements: regions of up to ~131k points
struct NpLeafLayout {                     // the recursion's leaves, left to right
  int off[NP_MAX_LEAVES];
  int len[NP_MAX_LEAVES];
  unsigned char depth[NP_MAX_LEAVES];
  int n_leaves;
};
template <typename T> __device__ __forceinline__ T np_add(T a, T b);
template <> __device__ __forceinline__ float np_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double np_add<double>(double a, double b) { return __dadd_rn(a, b); }

// Thread 0 lays out the leaves of numpy's recursion over n elements; ends with a block barrier.
__device__ __forceinline__ void np_build_leaves(int n, NpLeafLayout& L) {
  if (threadIdx.x == 0) {
    int so[40], sl[40], sd[40], sp = 0, nl = 0;
    so[0] = 0; sl[0] = n; sd[0] = 0; sp = 1;
    while (sp > 0) {
      --sp;
      const int o = so[sp], l = sl[sp], d = sd[sp];
      if (l <= 128 || nl + sp + 2 >= NP_MAX_LEAVES) {      // capacity guard (regions > ~131k points): the remaining ranges
        L.off[nl] = o; L.len[nl] = l; L.depth[nl] = (unsigned char)d;   // become long leaves -- still deterministic
        ++nl;
      } else {
        int n2 = l / 2;
        n2 -= n2 % 8;
        so[sp] = o + n2; sl[sp] = l - n2; sd[sp] = d + 1; ++sp;      // right half is summed second
        so[sp] = o; sl[sp] = n2; sd[sp] = d + 1; ++sp;
      }
    }
    L.n_leaves = nl;
  }
  __syncthreads();
}

// Sum of vals[pts[0..n)] in numpy's pairwise order over a prepared leaf layout, by one block of NP_BLOCK_THREADS threads;
// the result is valid on thread 0.  Groups of 8 lanes sum one leaf each (numpy's eight strided accumulators), thread 0 folds
// the leaf sums back up the tree: equal-depth neighbours combine, exactly the recursion's additions.  `sums` is block-shared
// scratch of NP_MAX_LEAVES elements.  Ends with a block barrier (scratch reusable).
template <typename T>
__device__ T np_sum_leaves(const T* __restrict__ vals, const int* __restrict__ pts, const NpLeafLayout& L, T* sums) {
  const int nl = L.n_leaves;
  const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  constexpr int G = NP_BLOCK_THREADS / 8;
  for (int leaf = grp; leaf < (nl + G - 1) / G * G; leaf += G) {
    const bool live = leaf < nl;
    const int o = live ? L.off[leaf] : 0, l = live ? L.len[leaf] : 0;
    T res = (T)0;
    if (l >= 8) {                                         // taken by whole 8-lane groups
      T r = vals[__ldg(&pts[o + j])];
      const int body = l - (l % 8);
      for (int i = 8; i < body; i += 8) r = np_add<T>(r, vals[__ldg(&pts[o + i + j])]);
      const T s1 = np_add<T>(r, __shfl_down_sync(gmask, r, 1, 8));
      const T s2 = np_add<T>(s1, __shfl_down_sync(gmask, s1, 2, 8));
      res = np_add<T>(s2, __shfl_down_sync(gmask, s2, 4, 8));
      if (j == 0)
        for (int i = body; i < l; ++i) res = np_add<T>(res, vals[__ldg(&pts[o + i])]);
    } else if (j == 0) {                                  // n < 8: sequential from zero
      for (int i = 0; i < l; ++i) res = np_add<T>(res, vals[__ldg(&pts[o + i])]);
    }
    if (live && j == 0) sums[leaf] = res;
  }
  __syncthreads();
  T total = (T)0;
  if (threadIdx.x == 0) {
    T sv[40];
    int sd[40], sp = 0;
    for (int leaf = 0; leaf < nl; ++leaf) {
      sv[sp] = sums[leaf]; sd[sp] = L.depth[leaf]; ++sp;
      while (sp >= 2 && sd[sp - 1] == sd[sp - 2]) {
        sv[sp - 2] = np_add<T>(sv[sp - 2], sv[sp - 1]);
        sd[sp - 2] -= 1;
        --sp;
      }
    }
    total = sp > 0 ? sv[0] : (T)0;
  }
  __syncthreads();
  return total;
}

}  // namespace lb
