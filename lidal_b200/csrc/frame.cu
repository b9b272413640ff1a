// Per-frame kernels either side of the network on the LiDAL scoring chain (SURVEY.md section 8f rows F1-F4 and the
// out_feat branch of row A8).  Everything here is HBM-bound integer / fp32 / fp64 work; the arithmetic follows the
// reference's numpy expressions operation by operation (no FMA contraction, numpy's summation orders) so results are
// reproducible to the last bit wherever libm agrees.
//
//   lb_register_points      dataset/prepare_kdtree_sk.py:76-80   raw scan + 4x4 pose -> registered float64 xyz  (F2)
//   lb_tta_views            dataset/sk_dataset.py:143-169        all TTA views of one scan in two launches       (F1)
//   lb_tta_feat_mean        score/prob_inference.py:103-105,116-118  out_feat gathered per point, mean over views (A8)
//   lb_segment_entropy      score/frame_level/segment_entropy.py:41-48                                           (F3)
//   lb_redal_point_scores   score/sv_level/ReDAL.py:63-67                                                        (F4)
//   lb_region_mean_f32 / lb_region_feat_mean   ReDAL.py:74-79 (and the f32 mean of LiDAL.py:98)
#include "common.cuh"
#include "np_sum.cuh"
#include <type_traits>

namespace lb {

static inline int grid_for(int64_t n, int block, int per_sm = 16) {
  int64_t b = (n + block - 1) / block, cap = (int64_t)sm_count() * per_sm;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------- F2: pose registration
// hcoords = np.hstack((coords, ones)) is float32; np.expand_dims(hcoords, 2) * pose.T promotes every product to float64;
// np.sum(axis=1) over the 4 products of a column adds them in k order starting from the first (no zero initial value).
struct Pose { double m[16]; };
__global__ void register_points_kernel(const float* __restrict__ raw, int64_t ld, int64_t n, const __grid_constant__ Pose P,
                                       double* __restrict__ xyz) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = (double)__ldg(&raw[i * ld]), y = (double)__ldg(&raw[i * ld + 1]), z = (double)__ldg(&raw[i * ld + 2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      // (pose.T)[k][j] = pose[j][k]
      double s = __dmul_rn(x, P.m[4 * j]);
      s = __dadd_rn(s, __dmul_rn(y, P.m[4 * j + 1]));
      s = __dadd_rn(s, __dmul_rn(z, P.m[4 * j + 2]));
      s = __dadd_rn(s, P.m[4 * j + 3]);                 // 1.0f * pose[j][3] is exact
      xyz[3 * i + j] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------- F1: batched TTA views
constexpr int MAX_VIEWS = 16;
struct ViewParams {
  double m[MAX_VIEWS][9];
  double r1[MAX_VIEWS][3];
  double r2[MAX_VIEWS][3];
};
__device__ __forceinline__ unsigned long long ord_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
  return __longlong_as_double((long long)b);
}
// view v, point i: coords_p = raw[:, :3] @ trans_m (float64, k order, no contraction); feats = (coords_p as f32, intensity);
// coords_p *= scale; per-view min / max through ordered-integer atomics (exact, order independent).
__global__ void __launch_bounds__(256)
tta_views_transform_kernel(const float4* __restrict__ raw, int64_t n, int reps, const __grid_constant__ ViewParams V, double scale,
                           double* __restrict__ cp /*[reps,n,3]*/, float4* __restrict__ feats /*[reps,n]*/,
                           unsigned long long* __restrict__ mm /*[reps][3] min | [reps][3] max*/) {
  __shared__ unsigned long long s_mm[6][8];
  const int v = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long lo[3] = {~0ULL, ~0ULL, ~0ULL}, hi[3] = {0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 r = __ldg(&raw[i]);
    const double x = r.x, y = r.y, z = r.z;
    double c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c[j] = __dadd_rn(__dadd_rn(__dmul_rn(x, V.m[v][j]), __dmul_rn(y, V.m[v][3 + j])), __dmul_rn(z, V.m[v][6 + j]));
    feats[(int64_t)v * n + i] = make_float4((float)c[0], (float)c[1], (float)c[2], r.w);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double s = __dmul_rn(c[j], scale);
      cp[((int64_t)v * n + i) * 3 + j] = s;
      const unsigned long long k = ord_key(s);
      lo[j] = k < lo[j] ? k : lo[j];
      hi[j] = k > hi[j] ? k : hi[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo[j], d), b = __shfl_xor_sync(0xffffffffu, hi[j], d);
      lo[j] = a < lo[j] ? a : lo[j];
      hi[j] = b > hi[j] ? b : hi[j];
    }
    if (lane == 0) { s_mm[j][w] = lo[j]; s_mm[3 + j][w] = hi[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    unsigned long long a = s_mm[threadIdx.x][0];
    for (int i = 1; i < 8; ++i) {
      const unsigned long long b = s_mm[threadIdx.x][i];
      a = threadIdx.x < 3 ? (b < a ? b : a) : (b > a ? b : a);
    }
    if (threadIdx.x < 3) atomicMin(&mm[v * 3 + threadIdx.x], a);
    else atomicMax(&mm[(reps + v) * 3 + threadIdx.x - 3], a);
  }
}
// dataset/sk_dataset.py:154-169: offset = -cmin + clip(fs - cmax + cmin - 0.001, 0, None) * r1 + clip(fs - cmax + cmin + 0.001, None, 0) * r2
// (float64, evaluated left to right), coords += offset, astype(int); key = view | x | y | z so that one ascending sort of all
// views reproduces np.unique(axis=0) per view followed by the collate concatenation.
__global__ void __launch_bounds__(256)
tta_views_quantize_kernel(const double* __restrict__ cp, int64_t n, int reps, const __grid_constant__ ViewParams V, double fs,
                          int coord_bits, const unsigned long long* __restrict__ mm, int4* __restrict__ coords,
                          int64_t* __restrict__ keys, int* __restrict__ err) {
  const int v = blockIdx.y;
  double off[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double cmin = ord_val(__ldg(&mm[v * 3 + j])), cmax = ord_val(__ldg(&mm[(reps + v) * 3 + j]));
    const double t = __dadd_rn(__dsub_rn(fs, cmax), cmin);
    double a = __dsub_rn(t, 0.001);
    a = a < 0.0 ? 0.0 : a;
    double b = __dadd_rn(t, 0.001);
    b = b > 0.0 ? 0.0 : b;
    off[j] = __dadd_rn(__dadd_rn(-cmin, __dmul_rn(a, V.r1[v][j])), __dmul_rn(b, V.r2[v][j]));
  }
  const long long lim = 1LL << coord_bits;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = (int64_t)v * n + i;
    const double fx = __dadd_rn(cp[3 * g], off[0]), fy = __dadd_rn(cp[3 * g + 1], off[1]), fz = __dadd_rn(cp[3 * g + 2], off[2]);
    const long long x = (long long)fx, y = (long long)fy, z = (long long)fz;      // astype(int): truncation toward zero
    // dataset/sk_dataset.py:160-161 asserts min >= 0 and max < full_scale on the float coordinates
    if (!(fx >= 0) || !(fy >= 0) || !(fz >= 0) || !(fx < fs) || !(fy < fs) || !(fz < fs) || x >= lim || y >= lim || z >= lim) atomicOr(err, 1);
    coords[g] = make_int4((int)x, (int)y, (int)z, v);
    keys[g] = (int64_t)(((unsigned long long)v << (3 * coord_bits)) | ((unsigned long long)x << (2 * coord_bits)) |
                        ((unsigned long long)y << coord_bits) | (unsigned long long)z);
  }
}

// ------------------------------------------------------------------------------------------- A8: out_feat mean over views
// out_feat_p_b = out_feat_v_b[inverse]; reshape(reps, -1, C); np.mean(axis=0): float32, views added in order, then / reps.
template <typename T>
__global__ void __launch_bounds__(256)
tta_feat_mean_kernel(const T* __restrict__ feat, int64_t ld, int64_t n_vox, const int64_t* __restrict__ inv, int reps,
                     int64_t n_pts, int c, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float fr = (float)reps;
  for (int64_t p = warp; p < n_pts; p += nwarps) {
    int64_t my_row = lane < reps ? __ldg(&inv[(int64_t)lane * n_pts + p]) : -1;
    for (int c0 = lane; c0 < c; c0 += 32) {
      float acc = 0.f;
      for (int v = 0; v < reps; ++v) {
        const int64_t row = __shfl_sync(0xffffffffu, my_row, v);
        const float x = (row >= 0 && row < n_vox) ? (float)feat[row * ld + c0] : 0.f;
        acc = v == 0 ? x : __fadd_rn(acc, x);
      }
      out[p * c + c0] = __fdiv_rn(acc, fr);
    }
  }
}

// ------------------------------------------------------------------------------------------- F3: segment entropy
// per region: q_c = count_c / len (float64); sv_sege += -q_c * log2(q_c + 1e-12) in class order;
// frame_sege += sv_sege * len / N in region order.
__global__ void __launch_bounds__(256)
segment_entropy_kernel(const int64_t* __restrict__ pred, int64_t n, int n_cls, const int* __restrict__ ptr,
                       const int* __restrict__ pts, double* __restrict__ terms) {
  __shared__ int hist[64];
  const int r = blockIdx.x;
  const int b = ptr[r], e = ptr[r + 1];
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    const int64_t c = __ldg(&pred[__ldg(&pts[i])]);
    if (c >= 0 && c < n_cls) atomicAdd(&hist[(int)c], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double len = (double)(e - b);
    double sege = 0.0;
    for (int c = 0; c < n_cls; ++c) {
      const double q = __ddiv_rn((double)hist[c], len);
      sege = __dadd_rn(sege, __dmul_rn(-q, log2(__dadd_rn(q, 1e-12))));
    }
    terms[r] = __ddiv_rn(__dmul_rn(sege, len), (double)n);
  }
}
__global__ void sum_in_order_kernel(const double* __restrict__ terms, int m, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < m; ++i) t = __dadd_rn(t, terms[i]);
    out[0] = t;
  }
}

// ------------------------------------------------------------------------------------------- F4: ReDAL point scores
// uncertain = np.mean(-prob * np.log2(prob + 1e-12), axis=1)  (float32 throughout; pairwise sum over classes)
// point_score = alpha * uncertain + gamma * curvature
__global__ void __launch_bounds__(256)
redal_point_kernel(const float* __restrict__ prob, int64_t n, int n_cls, const float* __restrict__ curv, float alpha, float gamma,
                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < n; p += nwarps) {
    float t = 0.f;
    if (lane < n_cls) {
      const float x = __ldg(&prob[p * n_cls + lane]);
      t = __fmul_rn(-x, log2f(__fadd_rn(x, 1e-12f)));
    }
    const float s = np_pairwise_sum32(t, n_cls, lane);
    if (lane == 0) {
      const float u = __fdiv_rn(s, (float)n_cls);
      const float c = curv ? __ldg(&curv[p]) : 0.f;
      out[p] = __fadd_rn(__fmul_rn(alpha, u), __fmul_rn(gamma, c));
    }
  }
}

// vals[pts[.]].mean() in numpy's float32 arithmetic (pairwise tree of np_sum.cuh), one block per region
__global__ void __launch_bounds__(NP_BLOCK_THREADS)
region_mean_f32_kernel(const float* __restrict__ vals, const int* __restrict__ ptr, const int* __restrict__ pts, float* __restrict__ out) {
  __shared__ NpLeafLayout L;
  __shared__ float sums[NP_MAX_LEAVES];
  const int r = blockIdx.x, b = ptr[r], e = ptr[r + 1];
  np_build_leaves(e - b, L);
  const float s = np_sum_leaves<float>(vals, pts + b, L, sums);
  if (threadIdx.x == 0) out[r] = __fdiv_rn(s, (float)(e - b));
}
// feat[pts[.], :].mean(0): float32, rows added in order per column (numpy reduces axis 0 of a C-contiguous matrix row by row)
__global__ void __launch_bounds__(128)
region_feat_mean_kernel(const float* __restrict__ feat, int64_t ld, int c, const int* __restrict__ ptr, const int* __restrict__ pts,
                        float* __restrict__ out) {
  const int r = blockIdx.y, b = ptr[r], e = ptr[r + 1];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= c) return;
  float acc = 0.f;
  int i = b;
  if (i < e) { acc = __ldg(&feat[(int64_t)__ldg(&pts[i]) * ld + col]); ++i; }
  for (; i + 4 <= e; i += 4) {                          // four independent loads in flight, adds stay in row order
    const float x0 = __ldg(&feat[(int64_t)__ldg(&pts[i]) * ld + col]), x1 = __ldg(&feat[(int64_t)__ldg(&pts[i + 1]) * ld + col]);
    const float x2 = __ldg(&feat[(int64_t)__ldg(&pts[i + 2]) * ld + col]), x3 = __ldg(&feat[(int64_t)__ldg(&pts[i + 3]) * ld + col]);
    acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, x0), x1), x2), x3);
  }
  for (; i < e; ++i) acc = __fadd_rn(acc, __ldg(&feat[(int64_t)__ldg(&pts[i]) * ld + col]));
  out[(int64_t)r * c + col] = __fdiv_rn(acc, (float)(e - b));
}

// ------------------------------------------------------------------------------------------- segmented voxelize (engine)
// F.spvoxelize (network/utils.py:56) without atomics: the points of every voxel are listed contiguously (segment_order: a
// counting sort of the point -> voxel index, built once per level from the coordinates alone), then one thread group per
// voxel adds its points' rows in fp32 and writes the mean in the 16-bit activation type -- no zero fill, no fp32 round trip.
__global__ void segment_scatter_kernel(const int* __restrict__ idx, int64_t n, int64_t m, const uint32_t* __restrict__ seg_ptr,
                                       unsigned* __restrict__ cursor, int* __restrict__ order) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = __ldg(&idx[i]);
    if (v < 0 || v >= m) continue;
    order[seg_ptr[v] + atomicAdd(&cursor[v], 1u)] = (int)i;
  }
}
template <typename T>
__device__ __forceinline__ void acc16(float (&a)[8], const uint4 w) {
  const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) {
      a[2 * j] += __uint_as_float(wd[j] << 16);
      a[2 * j + 1] += __uint_as_float(wd[j] & 0xffff0000u);
    } else {
      a[2 * j] += __half2float(__ushort_as_half((unsigned short)(wd[j] & 0xffffu)));
      a[2 * j + 1] += __half2float(__ushort_as_half((unsigned short)(wd[j] >> 16)));
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
voxelize_segments_kernel(const T* __restrict__ feats, int64_t ld_f, const int* __restrict__ order, const uint32_t* __restrict__ seg_ptr,
                         int64_t m, int c, T* __restrict__ out, int64_t ld_o) {
  const int cpr = c >> 3;                                   // 16-byte chunks per row
  const int64_t total = m * cpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / cpr;
    const int ch = (int)(t - v * cpr);
    const uint32_t b = seg_ptr[v], e = seg_ptr[v + 1];
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t i = b;
    for (; i + 2 <= e; i += 2) {                            // two independent gathers in flight
      const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(&feats[(int64_t)__ldg(&order[i]) * ld_f + ch * 8]));
      const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(&feats[(int64_t)__ldg(&order[i + 1]) * ld_f + ch * 8]));
      acc16<T>(a, w0);
      acc16<T>(a, w1);
    }
    if (i < e) acc16<T>(a, __ldg(reinterpret_cast<const uint4*>(&feats[(int64_t)__ldg(&order[i]) * ld_f + ch * 8])));
    const float r = e > b ? __frcp_rn((float)(e - b)) : 0.f;
    uint4 o;
    if (std::is_same<T, __nv_bfloat16>::value) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a[0] * r, a[1] * r), p1 = __floats2bfloat162_rn(a[2] * r, a[3] * r);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(a[4] * r, a[5] * r), p3 = __floats2bfloat162_rn(a[6] * r, a[7] * r);
      o = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                     *reinterpret_cast<uint32_t*>(&p3));
    } else {
      __half2 p0 = __floats2half2_rn(a[0] * r, a[1] * r), p1 = __floats2half2_rn(a[2] * r, a[3] * r);
      __half2 p2 = __floats2half2_rn(a[4] * r, a[5] * r), p3 = __floats2half2_rn(a[6] * r, a[7] * r);
      o = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                     *reinterpret_cast<uint32_t*>(&p3));
    }
    *reinterpret_cast<uint4*>(&out[v * ld_o + ch * 8]) = o;
  }
}

}  // namespace lb
using namespace lb;

extern "C" int lb_register_points(const float* raw, int64_t ld_raw, int64_t n, const double* pose, double* xyz, void* stream) {
  LB_CHECK_ARG(n >= 0 && ld_raw >= 3 && pose, "bad arguments");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(raw && xyz, "null pointer");
  Pose P;
  for (int i = 0; i < 16; ++i) P.m[i] = pose[i];
  register_points_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(raw, ld_raw, n, P, xyz); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_tta_views_ws_bytes(int64_t n, int reps) {
  if (n < 1) n = 1;
  return (((size_t)reps * n * 24 + 255) & ~(size_t)255) + 256 + (size_t)reps * 6 * 8;
}
extern "C" int lb_tta_views(const float* raw, int64_t n, int reps, const double* trans_m, const double* r1, const double* r2,
                            double scale, double full_scale, int coord_bits, float* feats_p, int32_t* coords_p, int64_t* keys,
                            int32_t* err_flag, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && reps >= 1 && reps <= MAX_VIEWS && trans_m && r1 && r2, "bad arguments (1 <= reps <= 16)");
  LB_CHECK_ARG(coord_bits > 0 && coord_bits <= 19 && err_flag && ws, "bad arguments");
  if (ws_bytes < lb_tta_views_ws_bytes(n, reps)) { set_error("lb_tta_views: workspace too small"); return LB_ECAP; }
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(raw && feats_p && coords_p && keys && ((((uintptr_t)raw) | ((uintptr_t)feats_p) | ((uintptr_t)coords_p)) & 15) == 0,
               "null or unaligned pointer");
  cudaStream_t st = as_stream(stream);
  ViewParams V;
  for (int v = 0; v < reps; ++v) {
    for (int i = 0; i < 9; ++i) V.m[v][i] = trans_m[v * 9 + i];
    for (int i = 0; i < 3; ++i) { V.r1[v][i] = r1[v * 3 + i]; V.r2[v][i] = r2[v * 3 + i]; }
  }
  double* cp = (double*)ws;
  unsigned long long* mm = (unsigned long long*)((char*)ws + (((size_t)reps * n * 24 + 255) & ~(size_t)255));
  LB_CUDA(cudaMemsetAsync(mm, 0xFF, (size_t)reps * 3 * 8, st));
  LB_CUDA(cudaMemsetAsync(mm + reps * 3, 0, (size_t)reps * 3 * 8, st));
  int gx = grid_for(n, 256, 4);
  tta_views_transform_kernel<<<dim3(gx, reps), 256, 0, st>>>((const float4*)raw, n, reps, V, scale, cp, (float4*)feats_p, mm); LB_LAUNCHED(1);
  tta_views_quantize_kernel<<<dim3(gx, reps), 256, 0, st>>>(cp, n, reps, V, full_scale, coord_bits, mm, (int4*)coords_p, keys, err_flag); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_tta_feat_mean(const void* feat, int feat_dtype, int64_t ld, int64_t n_vox, const int64_t* inverse, int reps,
                                int64_t n_pts, int c, float* out, void* stream) {
  LB_CHECK_ARG(reps >= 1 && reps <= 32 && n_pts >= 0 && n_vox >= 0 && c > 0 && ld >= c, "bad sizes (reps <= 32)");
  if (n_pts == 0) return LB_OK;
  LB_CHECK_ARG(feat && inverse && out, "null pointer");
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(n_pts, 8, 16);
  if (feat_dtype == LB_DT_F32) { tta_feat_mean_kernel<float><<<g, 256, 0, st>>>((const float*)feat, ld, n_vox, inverse, reps, n_pts, c, out); LB_LAUNCHED(1); }
  else if (feat_dtype == LB_DT_BF16) { tta_feat_mean_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)feat, ld, n_vox, inverse, reps, n_pts, c, out); LB_LAUNCHED(1); }
  else if (feat_dtype == LB_DT_F16) { tta_feat_mean_kernel<__half><<<g, 256, 0, st>>>((const __half*)feat, ld, n_vox, inverse, reps, n_pts, c, out); LB_LAUNCHED(1); }
  else { set_error("lb_tta_feat_mean: bad dtype"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_segment_entropy(const int64_t* pred, int64_t n, int n_cls, const int32_t* region_ptr, const int32_t* region_pts,
                                  int n_regions, double* out, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && n_cls >= 1 && n_cls <= 64 && n_regions >= 0 && out && ws, "bad arguments (n_cls <= 64)");
  if (ws_bytes < (size_t)(n_regions > 0 ? n_regions : 1) * 8) { set_error("lb_segment_entropy: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (n_regions > 0) {
    LB_CHECK_ARG(pred && region_ptr && region_pts, "null pointer");
    segment_entropy_kernel<<<n_regions, 256, 0, st>>>(pred, n, n_cls, region_ptr, region_pts, (double*)ws); LB_LAUNCHED(1);
  }
  sum_in_order_kernel<<<1, 32, 0, st>>>((const double*)ws, n_regions, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_redal_point_scores(const float* prob, int64_t n, int n_cls, const float* curvature, float alpha, float gamma,
                                     float* point_score, void* stream) {
  LB_CHECK_ARG(n >= 0 && n_cls >= 1 && n_cls <= 32, "n_cls must be in [1,32]");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(prob && point_score, "null pointer");
  redal_point_kernel<<<grid_for(n, 8, 16), 256, 0, as_stream(stream)>>>(prob, n, n_cls, curvature, alpha, gamma, point_score); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_region_mean_f32(const float* vals, const int32_t* region_ptr, const int32_t* region_pts, int n_regions, float* out,
                                  void* stream) {
  LB_CHECK_ARG(n_regions >= 0, "n_regions < 0");
  if (n_regions == 0) return LB_OK;
  LB_CHECK_ARG(vals && region_ptr && region_pts && out, "null pointer");
  region_mean_f32_kernel<<<n_regions, NP_BLOCK_THREADS, 0, as_stream(stream)>>>(vals, region_ptr, region_pts, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_region_feat_mean(const float* feat, int64_t ld, int c, const int32_t* region_ptr, const int32_t* region_pts,
                                   int n_regions, float* out, void* stream) {
  LB_CHECK_ARG(n_regions >= 0 && c > 0 && ld >= c, "bad sizes");
  if (n_regions == 0) return LB_OK;
  LB_CHECK_ARG(feat && region_ptr && region_pts && out, "null pointer");
  region_feat_mean_kernel<<<dim3((c + 127) / 128, n_regions), 128, 0, as_stream(stream)>>>(feat, ld, c, region_ptr, region_pts, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_segment_order_ws_bytes(int64_t m) {
  if (m < 1) m = 1;
  return (((size_t)m * 4 + 255) & ~(size_t)255) + scan_ws_bytes(m) + 512;
}
extern "C" int lb_segment_order(const int32_t* idx, int64_t n, const int32_t* counts, int64_t m, int32_t* seg_ptr, int32_t* order,
                                void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && seg_ptr && ws, "bad arguments");
  if (ws_bytes < lb_segment_order_ws_bytes(m)) { set_error("lb_segment_order: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (m == 0) { LB_CUDA(cudaMemsetAsync(seg_ptr, 0, 4, st)); return LB_OK; }
  LB_CHECK_ARG(counts && (n == 0 || (idx && order)), "null pointer");
  unsigned* cursor = (unsigned*)ws;
  void* sws = (char*)ws + (((size_t)m * 4 + 255) & ~(size_t)255);
  // seg_ptr[0..m) = exclusive scan of counts, seg_ptr[m] = total
  int rc = exclusive_scan_u32((const uint32_t*)counts, (uint32_t*)seg_ptr, m, (uint32_t*)seg_ptr + m, sws, st);
  if (rc != LB_OK) return rc;
  LB_CUDA(cudaMemsetAsync(cursor, 0, (size_t)m * 4, st));
  if (n > 0) { segment_scatter_kernel<<<grid_for(n, 256), 256, 0, st>>>(idx, n, m, (const uint32_t*)seg_ptr, cursor, order); LB_LAUNCHED(1); }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_voxelize_segments(const void* feats, int dtype, int64_t ld_feats, const int32_t* order, const int32_t* seg_ptr, int64_t m,
                                    int c, void* out, int64_t ld_out, void* stream) {
  LB_CHECK_ARG(m >= 0 && c > 0 && c % 8 == 0 && ld_feats % 8 == 0 && ld_out % 8 == 0 && ld_feats >= c && ld_out >= c, "bad sizes (c, strides % 8)");
  if (m == 0) return LB_OK;
  LB_CHECK_ARG(feats && order && seg_ptr && out && ((((uintptr_t)feats) | ((uintptr_t)out)) & 15) == 0, "null or unaligned pointer");
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(m * (c / 8), 256, 32);
  if (dtype == LB_DT_BF16) { voxelize_segments_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)feats, ld_feats, order, (const uint32_t*)seg_ptr, m, c, (__nv_bfloat16*)out, ld_out); LB_LAUNCHED(1); }
  else if (dtype == LB_DT_F16) { voxelize_segments_kernel<__half><<<g, 256, 0, st>>>((const __half*)feats, ld_feats, order, (const uint32_t*)seg_ptr, m, c, (__half*)out, ld_out); LB_LAUNCHED(1); }
  else { set_error("lb_voxelize_segments: 16-bit features only"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
