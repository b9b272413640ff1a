// SPVCNN point<->voxel kernels, dtype casts, and the prob_inference tail (row A8).
// All of these are HBM-bound: one warp walks one feature row with float4 accesses so every global
// transaction is a full 128-byte line; indices are read once per row, not once per channel.
#include "common.cuh"

namespace lb {

static inline int rows_grid(int64_t rows, int warps_per_block) {
  int64_t b = (rows + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)sm_count() * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------- count
__global__ void count_kernel(const int* __restrict__ idx, int64_t n, int* __restrict__ counts, int64_t m) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int v = __ldg(&idx[i]);
    if (v >= 0 && v < m) atomicAdd(&counts[v], 1);
  }
}

// ---------------------------------------------------------------------------------------- voxelize
// out[idx[i]] += feats[i] / counts[idx[i]]   (torchsparse voxelize_forward: fp32 atomics)
__global__ void voxelize_fwd_kernel(const float* __restrict__ feats, const int* __restrict__ idx,
                                    const int* __restrict__ counts, int64_t n, int64_t m, int c,
                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    int v = __ldg(&idx[i]);
    if (v < 0 || v >= m) continue;
    int cnt = __ldg(&counts[v]);
    if (cnt == 0) continue;
    float fc = (float)cnt;
    for (int j = lane; j < c; j += 32) atomicAdd(&out[(int64_t)v * c + j], feats[i * c + j] / fc);
  }
}
// grad_feats[i] = grad_out[idx[i]] / counts[idx[i]]
__global__ void voxelize_bwd_kernel(const float* __restrict__ gout, const int* __restrict__ idx,
                                    const int* __restrict__ counts, int64_t n, int64_t m, int c,
                                    float* __restrict__ gf) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    int v = __ldg(&idx[i]);
    bool ok = v >= 0 && v < m && __ldg(&counts[v]) != 0;
    float fc = ok ? (float)__ldg(&counts[v]) : 1.f;
    for (int j = lane; j < c; j += 32) gf[i * c + j] = ok ? gout[(int64_t)v * c + j] / fc : 0.f;
  }
}

// ---------------------------------------------------------------------------------------- devoxelize
// out[i] = sum_k w[i,k] * feats[idx[i,k]]  (corner order 0..7, skipping -1) -- deterministic, no atomics
template <int VEC>
__global__ void devoxelize_fwd_kernel(const float* __restrict__ feats, const int* __restrict__ idx,
                                      const float* __restrict__ w, int64_t n, int64_t m, int c,
                                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    int my_idx = -1;
    float my_w = 0.f;
    if (lane < 8) {
      my_idx = __ldg(&idx[i * 8 + lane]);
      my_w = __ldg(&w[i * 8 + lane]);
      if (my_idx >= m) my_idx = -1;
    }
    int rk[8];
    float wk8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {              // broadcast once, outside the (divergent) channel loop
      rk[k] = __shfl_sync(0xffffffffu, my_idx, k);
      wk8[k] = __shfl_sync(0xffffffffu, my_w, k);
    }
    for (int j = lane * VEC; j < c; j += 32 * VEC) {
      float acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = rk[k];
        const float wk = wk8[k];
        if (r >= 0) {
          if (VEC == 4) {
            float4 f = __ldg((const float4*)&feats[(int64_t)r * c + j]);
            acc[0] = __fadd_rn(acc[0], __fmul_rn(wk, f.x)); acc[1] = __fadd_rn(acc[1], __fmul_rn(wk, f.y));
            acc[2] = __fadd_rn(acc[2], __fmul_rn(wk, f.z)); acc[3] = __fadd_rn(acc[3], __fmul_rn(wk, f.w));
          } else {
            acc[0] = __fadd_rn(acc[0], __fmul_rn(wk, __ldg(&feats[(int64_t)r * c + j])));
          }
        }
      }
      if (VEC == 4) *(float4*)&out[i * c + j] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else out[i * c + j] = acc[0];
    }
  }
}
// grad_feats[idx[i,k]] += w[i,k] * grad_out[i]
__global__ void devoxelize_bwd_kernel(const float* __restrict__ gout, const int* __restrict__ idx,
                                      const float* __restrict__ w, int64_t n, int64_t m, int c,
                                      float* __restrict__ gf) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    for (int k = 0; k < 8; ++k) {
      int r = __ldg(&idx[i * 8 + k]);
      if (r < 0 || r >= m) continue;
      float wk = __ldg(&w[i * 8 + k]);
      for (int j = lane; j < c; j += 32) atomicAdd(&gf[(int64_t)r * c + j], wk * gout[i * c + j]);
    }
  }
}

// ---------------------------------------------------------------------------------------- trilinear weights
// F.calc_ti_weights: fp32 arithmetic in the same operation order as the torch expression.
__global__ void ti_weights_kernel(const float* __restrict__ coords, int64_t ld, const int64_t* __restrict__ idx,
                                  int64_t n, float scale, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float p[3], pf[3], pc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p[a] = __ldg(&coords[i * ld + a]);
      pf[a] = (scale != 1.f) ? __fmul_rn(floorf(__fdiv_rn(p[a], scale)), scale) : floorf(p[a]);
      pc[a] = __fadd_rn(pf[a], scale);
    }
    float hx = __fsub_rn(pc[0], p[0]), lx = __fsub_rn(p[0], pf[0]);
    float hy = __fsub_rn(pc[1], p[1]), ly = __fsub_rn(p[1], pf[1]);
    float hz = __fsub_rn(pc[2], p[2]), lz = __fsub_rn(p[2], pf[2]);
    float w[8];
    w[0] = __fmul_rn(__fmul_rn(hx, hy), hz); w[1] = __fmul_rn(__fmul_rn(hx, hy), lz);
    w[2] = __fmul_rn(__fmul_rn(hx, ly), hz); w[3] = __fmul_rn(__fmul_rn(hx, ly), lz);
    w[4] = __fmul_rn(__fmul_rn(lx, hy), hz); w[5] = __fmul_rn(__fmul_rn(lx, hy), lz);
    w[6] = __fmul_rn(__fmul_rn(lx, ly), hz); w[7] = __fmul_rn(__fmul_rn(lx, ly), lz);
    const float s3 = scale * scale * scale;    // python: scale ** 3 on an int stride -> exact in fp32 here
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (scale != 1.f) w[k] = __fdiv_rn(w[k], s3);
      if (__ldg(&idx[(int64_t)k * n + i]) == -1) w[k] = 0.f;
      sum = __fadd_rn(sum, w[k]);
    }
    sum = __fadd_rn(sum, 1e-8f);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[(int64_t)k * n + i] = __fdiv_rn(w[k], sum);
  }
}

// ---------------------------------------------------------------------------------------- casts
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, int64_t ld_s, D* __restrict__ dst, int64_t ld_d, int64_t rows,
                            int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = t / cols, c = t - r * cols;
    dst[r * ld_d + c] = (D)(float)src[r * ld_s + c];
  }
}

// ---------------------------------------------------------------------------------------- TTA tail
// prob[p] = mean_v softmax(logits[inverse[v*n_pts + p]]); pred = argmax.  One warp per point: lane c owns
// class c (n_cls <= 32); the 8 gathered logit rows are 76-byte segments -> 1-2 sectors each.
__global__ void tta_kernel(const float* __restrict__ logits, int64_t n_vox, int n_cls, const int64_t* __restrict__ inv,
                           int reps, int64_t n_pts, float* __restrict__ prob, int64_t* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < n_pts; p += nwarps) {
    float acc = 0.f;
    for (int v = 0; v < reps; ++v) {
      int64_t row = __ldg(&inv[(int64_t)v * n_pts + p]);
      float x = (lane < n_cls && row >= 0 && row < n_vox) ? __ldg(&logits[row * n_cls + lane]) : -INFINITY;
      float mx = x;
#pragma unroll
      for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      float e = lane < n_cls ? expf(x - mx) : 0.f;
      float s = e;
#pragma unroll
      for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      acc = __fadd_rn(acc, __fdiv_rn(e, s));            // np.mean(axis=0): sequential adds over views
    }
    float pm = __fdiv_rn(acc, (float)reps);
    if (lane < n_cls) prob[p * n_cls + lane] = pm;
    // argmax with first-index tie break (np.argmax)
    float bv = lane < n_cls ? pm : -INFINITY;
    int bi = lane;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, d);
      int oi = __shfl_xor_sync(0xffffffffu, bi, d);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) pred[p] = bi;
  }
}

}  // namespace lb
using namespace lb;

extern "C" int lb_count(const int32_t* idx, int64_t n, int32_t* counts, int64_t m, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0, "negative size");
  cudaStream_t st = as_stream(stream);
  if (m > 0) { LB_CHECK_ARG(counts, "null counts"); LB_CUDA(cudaMemsetAsync(counts, 0, (size_t)m * 4, st)); }
  if (n == 0 || m == 0) return LB_OK;
  LB_CHECK_ARG(idx, "null idx");
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  count_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(idx, n, counts, m); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_voxelize_fwd(const float* feats, const int32_t* idx, const int32_t* counts, int64_t n, int64_t m,
                               int c, float* out, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && c > 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (m > 0) { LB_CHECK_ARG(out, "null out"); LB_CUDA(cudaMemsetAsync(out, 0, (size_t)m * c * 4, st)); }
  if (n == 0 || m == 0) return LB_OK;
  LB_CHECK_ARG(feats && idx && counts, "null pointer");
  voxelize_fwd_kernel<<<rows_grid(n, 8), 256, 0, st>>>(feats, idx, counts, n, m, c, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_voxelize_bwd(const float* gout, const int32_t* idx, const int32_t* counts, int64_t n, int64_t m,
                               int c, float* gf, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && c > 0, "bad sizes");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(gout && idx && counts && gf, "null pointer");
  voxelize_bwd_kernel<<<rows_grid(n, 8), 256, 0, as_stream(stream)>>>(gout, idx, counts, n, m, c, gf); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_devoxelize_fwd(const float* feats, const int32_t* idx, const float* w, int64_t n, int64_t m, int c,
                                 float* out, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && c > 0, "bad sizes");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(feats && idx && w && out, "null pointer");
  bool vec = (c % 4 == 0) && (((uintptr_t)feats | (uintptr_t)out) & 15) == 0;
  if (vec) { devoxelize_fwd_kernel<4><<<rows_grid(n, 8), 256, 0, as_stream(stream)>>>(feats, idx, w, n, m, c, out); LB_LAUNCHED(1); }
  else { devoxelize_fwd_kernel<1><<<rows_grid(n, 8), 256, 0, as_stream(stream)>>>(feats, idx, w, n, m, c, out); LB_LAUNCHED(1); }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_devoxelize_bwd(const float* gout, const int32_t* idx, const float* w, int64_t n, int64_t m, int c,
                                 float* gf, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && c > 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (m > 0) { LB_CHECK_ARG(gf, "null grad"); LB_CUDA(cudaMemsetAsync(gf, 0, (size_t)m * c * 4, st)); }
  if (n == 0 || m == 0) return LB_OK;
  LB_CHECK_ARG(gout && idx && w, "null pointer");
  devoxelize_bwd_kernel<<<rows_grid(n, 8), 256, 0, st>>>(gout, idx, w, n, m, c, gf); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_ti_weights(const float* coords, int64_t ld, const int64_t* idx, int64_t n, float scale, float* out,
                             void* stream) {
  LB_CHECK_ARG(n >= 0 && ld >= 3 && scale > 0.f, "bad sizes");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(coords && idx && out, "null pointer");
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  ti_weights_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, as_stream(stream)>>>(coords, ld, idx, n, scale, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

template <typename S>
static int cast_dispatch(const S* src, int64_t ld_s, void* dst, int dd, int64_t ld_d, int64_t rows, int64_t cols,
                         cudaStream_t st) {
  int64_t blocks = (rows * cols + 255) / 256, cap = (int64_t)sm_count() * 32;
  int g = (int)(blocks > cap ? cap : blocks);
  if (dd == LB_DT_F32) { cast_kernel<S, float><<<g, 256, 0, st>>>(src, ld_s, (float*)dst, ld_d, rows, cols); LB_LAUNCHED(1); }
  else if (dd == LB_DT_BF16) { cast_kernel<S, __nv_bfloat16><<<g, 256, 0, st>>>(src, ld_s, (__nv_bfloat16*)dst, ld_d, rows, cols); LB_LAUNCHED(1); }
  else if (dd == LB_DT_F16) { cast_kernel<S, __half><<<g, 256, 0, st>>>(src, ld_s, (__half*)dst, ld_d, rows, cols); LB_LAUNCHED(1); }
  else { set_error("lb_cast: bad dst dtype"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_cast(const void* src, int sd, int64_t ld_s, void* dst, int dd, int64_t ld_d, int64_t rows,
                       int64_t cols, void* stream) {
  LB_CHECK_ARG(rows >= 0 && cols >= 0, "bad sizes");
  if (rows == 0 || cols == 0) return LB_OK;         // empty tensors carry arbitrary strides
  LB_CHECK_ARG(ld_s >= cols && ld_d >= cols, "row stride smaller than the column count");
  LB_CHECK_ARG(src && dst, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (sd == LB_DT_F32) return cast_dispatch((const float*)src, ld_s, dst, dd, ld_d, rows, cols, st);
  if (sd == LB_DT_BF16) return cast_dispatch((const __nv_bfloat16*)src, ld_s, dst, dd, ld_d, rows, cols, st);
  if (sd == LB_DT_F16) return cast_dispatch((const __half*)src, ld_s, dst, dd, ld_d, rows, cols, st);
  set_error("lb_cast: bad src dtype");
  return LB_EINVAL;
}

// ---------------------------------------------------------------------------------------- scaled 16-bit cast (training)
// Gradients of a mean-reduced loss sit far below fp16's normal range; the backward convolutions therefore cast g * s with
// s = the power of two that brings max|g| near `target`, and fold 1 / s into the kernel epilogue (exact: powers of two).
namespace lb {
__global__ void absmax_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols, unsigned* __restrict__ out) {
  float m = 0.f;
  const int64_t total = rows * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cols;
    const float v = fabsf(src[r * ld + (t - r * cols)]);
    m = (v == v && v < INFINITY) ? fmaxf(m, v) : m;                    // NaN / Inf do not steer the scale
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bits
}
template <typename D>
__global__ void cast_scaled_kernel(const float* __restrict__ src, int64_t ld_s, D* __restrict__ dst, int64_t ld_d, int64_t rows,
                                   int64_t cols, const float* __restrict__ amax, float target, float* __restrict__ scale_out,
                                   float* __restrict__ inv_vec) {
  const float a = *amax;
  const float s = (a > 0.f && a < INFINITY) ? exp2f(floorf(log2f(target / a))) : 1.f;
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0 && scale_out) { scale_out[0] = s; scale_out[1] = 1.f / s; }
    if (inv_vec) inv_vec[threadIdx.x] = 1.f / s;                        // blockDim.x == 256 entries
  }
  const int64_t total = rows * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cols, c = t - r * cols;
    dst[r * ld_d + c] = (D)(src[r * ld_s + c] * s);           // round-to-nearest conversion constructors
  }
}
}  // namespace lb
extern "C" int lb_absmax_f32(const float* src, int64_t ld, int64_t rows, int64_t cols, float* amax, void* stream) {
  LB_CHECK_ARG(rows >= 0 && cols >= 0 && amax, "bad arguments");
  cudaStream_t st = as_stream(stream);
  LB_CUDA(cudaMemsetAsync(amax, 0, 4, st));
  if (rows == 0 || cols == 0) return LB_OK;
  LB_CHECK_ARG(src && ld >= cols, "null pointer or row stride smaller than the column count");
  int64_t blocks = (rows * cols + 1023) / 1024, cap = (int64_t)sm_count() * 16;
  absmax_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(src, ld, rows, cols, (unsigned*)amax); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_cast_scaled(const float* src, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst, int64_t rows, int64_t cols,
                              const float* amax, float target, float* scale_out, float* inv_vec, void* stream) {
  LB_CHECK_ARG(rows >= 0 && cols >= 0 && amax && target > 0.f, "bad arguments");
  LB_CHECK_ARG(rows == 0 || cols == 0 || (src && dst && ld_src >= cols && ld_dst >= cols), "null pointer or bad strides");
  cudaStream_t st = as_stream(stream);
  int64_t blocks = (rows * cols + 255) / 256, cap = (int64_t)sm_count() * 32;
  int g = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  if (dst_dtype == LB_DT_F16) { cast_scaled_kernel<__half><<<g, 256, 0, st>>>(src, ld_src, (__half*)dst, ld_dst, rows, cols, amax, target, scale_out, inv_vec); LB_LAUNCHED(1); }
  else if (dst_dtype == LB_DT_BF16) { cast_scaled_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, amax, target, scale_out, inv_vec); LB_LAUNCHED(1); }
  else { set_error("lb_cast_scaled: dst must be a 16-bit type"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_tta_softmax_mean_argmax(const float* logits, int64_t n_vox, int n_cls, const int64_t* inverse,
                                          int reps, int64_t n_pts, float* prob, int64_t* pred, void* stream) {
  LB_CHECK_ARG(n_cls > 0 && n_cls <= 32, "n_cls must be in [1,32]");
  LB_CHECK_ARG(reps > 0 && n_pts >= 0 && n_vox >= 0, "bad sizes");
  if (n_pts == 0) return LB_OK;
  LB_CHECK_ARG(logits && inverse && prob && pred, "null pointer");
  tta_kernel<<<rows_grid(n_pts, 8), 256, 0, as_stream(stream)>>>(logits, n_vox, n_cls, inverse, reps, n_pts, prob, pred); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---------------------------------------------------------------------------------------- fused point queries (engine)
namespace lb {
// network/utils.py:42-48: cell = floor(p / s) * s (int), hash, lookup among the voxel hashes.
__global__ void point_cell_query_kernel(const float* __restrict__ pts, int64_t ld, int64_t n, int s, TableView t,
                                        int* __restrict__ idx) {
  const float fs = (float)s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int cx = (int)floorf(__fdiv_rn(pts[i * ld], fs)) * s, cy = (int)floorf(__fdiv_rn(pts[i * ld + 1], fs)) * s,
        cz = (int)floorf(__fdiv_rn(pts[i * ld + 2], fs)) * s;
    idx[i] = table_find(t, (uint64_t)fnv60(cx, cy, cz, (int)pts[i * ld + ld - 1]));
  }
}
// network/utils.py:69-79: 8 corners (x slowest, z fastest) + F.calc_ti_weights, one thread per point.
__global__ void point_corner_query_kernel(const float* __restrict__ pts, int64_t ld, int64_t n, int s, TableView t,
                                          int* __restrict__ idx /*[n,8]*/, float* __restrict__ w /*[n,8]*/) {
  const float fs = (float)s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float p[3], lo[3], hi[3];
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p[a] = pts[i * ld + a];
      float fl = floorf(__fdiv_rn(p[a], fs));
      c[a] = (int)fl * s;
      float pf = (s != 1) ? __fmul_rn(fl, fs) : floorf(p[a]);
      float pc = __fadd_rn(pf, fs);
      hi[a] = __fsub_rn(pc, p[a]);
      lo[a] = __fsub_rn(p[a], pf);
    }
    const int b = (int)pts[i * ld + ld - 1];
    const float s3 = fs * fs * fs;
    float wk[8], sum = 0.f;
    int ik[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ox = (k >> 2) & 1, oy = (k >> 1) & 1, oz = k & 1;
      ik[k] = table_find(t, (uint64_t)fnv60(c[0] + ox * s, c[1] + oy * s, c[2] + oz * s, b));
      float v = __fmul_rn(__fmul_rn(ox ? lo[0] : hi[0], oy ? lo[1] : hi[1]), oz ? lo[2] : hi[2]);
      if (s != 1) v = __fdiv_rn(v, s3);
      if (ik[k] < 0) v = 0.f;
      wk[k] = v;
      sum = __fadd_rn(sum, v);
    }
    sum = __fadd_rn(sum, 1e-8f);
    int4* ip = (int4*)&idx[i * 8];
    float4* wp = (float4*)&w[i * 8];
    ip[0] = make_int4(ik[0], ik[1], ik[2], ik[3]);
    ip[1] = make_int4(ik[4], ik[5], ik[6], ik[7]);
    wp[0] = make_float4(__fdiv_rn(wk[0], sum), __fdiv_rn(wk[1], sum), __fdiv_rn(wk[2], sum), __fdiv_rn(wk[3], sum));
    wp[1] = make_float4(__fdiv_rn(wk[4], sum), __fdiv_rn(wk[5], sum), __fdiv_rn(wk[6], sum), __fdiv_rn(wk[7], sum));
  }
}

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ld_f<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void st_f<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// 8 lanes x 2 channels... generic row-strided variants: lane j walks channels j, j+32, ... (2- or 4-byte elements)
template <typename TI>
__global__ void voxelize_ex_kernel(const TI* __restrict__ feats, int64_t ld_f_, const int* __restrict__ idx,
                                   const int* __restrict__ counts, int64_t n, int64_t m, int c, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    int v = __ldg(&idx[i]);
    if (v < 0 || v >= m) continue;
    int cnt = __ldg(&counts[v]);
    if (cnt == 0) continue;
    float fc = (float)cnt;
    for (int j = lane; j < c; j += 32) atomicAdd(&out[(int64_t)v * c + j], ld_f<TI>(&feats[i * ld_f_ + j]) / fc);
  }
}
// c % 4 == 0, 16-byte aligned rows: one lane owns 4 channels -> one float4 atomic (sm_90+), or a plain 16-byte store when
// the voxel holds a single point (no other writer exists for that row).
template <typename TI>
__global__ void voxelize_ex_vec4_kernel(const TI* __restrict__ feats, int64_t ld_f_, const int* __restrict__ idx,
                                        const int* __restrict__ counts, int64_t n, int64_t m, int c, float* __restrict__ out) {
  // One thread = 4 channels of VOX_RUN consecutive points.  Points arrive in scan order, so consecutive points mostly fall
  // into the same (coarse) voxel: their contributions are summed in registers and leave as ONE float4 atomic per run
  // instead of one per point (stride-16 voxels hold ~25 points each).
  constexpr int VOX_RUN = 16;
  const int cpr = c >> 2;
  const int64_t chunks = (n + VOX_RUN - 1) / VOX_RUN;
  const int64_t total = chunks * cpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = t / cpr;
    const int j = (int)(t - ch * cpr) * 4;
    const int64_t i0 = ch * VOX_RUN, i1 = (i0 + VOX_RUN < n) ? i0 + VOX_RUN : n;
    int cur = -1, cnt = 0;
    float fc = 1.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = i0; i < i1; ++i) {
      const int v = __ldg(&idx[i]);
      if (v < 0 || v >= m) continue;
      if (v != cur) {
        if (cur >= 0 && cnt > 0) {
          float4* dst = (float4*)&out[(int64_t)cur * c + j];
          if (cnt == 1) *dst = acc;                     // single-point voxel: no other writer exists for this row
          else atomicAdd(dst, acc);
        }
        cur = v;
        cnt = __ldg(&counts[v]);
        fc = (float)cnt;
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (cnt > 0) {
        const TI* f = &feats[i * ld_f_ + j];
        acc.x += ld_f<TI>(f) / fc;
        acc.y += ld_f<TI>(f + 1) / fc;
        acc.z += ld_f<TI>(f + 2) / fc;
        acc.w += ld_f<TI>(f + 3) / fc;
      }
    }
    if (cur >= 0 && cnt > 0) {
      float4* dst = (float4*)&out[(int64_t)cur * c + j];
      if (cnt == 1) *dst = acc;
      else atomicAdd(dst, acc);
    }
  }
}
template <typename T> __device__ __forceinline__ float lo16_to_f32(uint32_t w);
template <typename T> __device__ __forceinline__ float hi16_to_f32(uint32_t w);
// 16-bit features (engine path): same run-length scheme, 8-byte loads of 4 channels, the run is summed first and scaled
// once by 1 / count at the flush (the reference adds f / count per point in atomic order -- neither order is canonical).
template <typename TI>
__global__ void voxelize16_vec4_kernel(const TI* __restrict__ feats, int64_t ld_f_, const int* __restrict__ idx,
                                       const int* __restrict__ counts, int64_t n, int64_t m, int c, float* __restrict__ out) {
  constexpr int VOX_RUN = 16;
  const int cpr = c >> 2;
  const int64_t chunks = (n + VOX_RUN - 1) / VOX_RUN;
  const int64_t total = chunks * cpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = t / cpr;
    const int j = (int)(t - ch * cpr) * 4;
    const int64_t i0 = ch * VOX_RUN, i1 = (i0 + VOX_RUN < n) ? i0 + VOX_RUN : n;
    int cur = -1, cnt = 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&]() {
      if (cur >= 0 && cnt > 0) {
        const float r = __frcp_rn((float)cnt);
        const float4 v = make_float4(acc.x * r, acc.y * r, acc.z * r, acc.w * r);
        float4* dst = (float4*)&out[(int64_t)cur * c + j];
        if (cnt == 1) *dst = v;                        // single-point voxel: no other writer exists for this row
        else atomicAdd(dst, v);
      }
    };
    for (int64_t i = i0; i < i1; ++i) {
      const int v = __ldg(&idx[i]);
      if (v < 0 || v >= m) continue;
      if (v != cur) {
        flush();
        cur = v;
        cnt = __ldg(&counts[v]);
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (cnt > 0) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(&feats[i * ld_f_ + j]));
        acc.x += lo16_to_f32<TI>(w.x); acc.y += hi16_to_f32<TI>(w.x);
        acc.z += lo16_to_f32<TI>(w.y); acc.w += hi16_to_f32<TI>(w.y);
      }
    }
    flush();
  }
}
template <typename TI, typename TO>
__global__ void devoxelize_ex_kernel(const TI* __restrict__ feats, int64_t ld_f_, const int* __restrict__ idx,
                                     const float* __restrict__ w, int64_t n, int64_t m, int c, TO* __restrict__ out,
                                     int64_t ld_o) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    int my_idx = -1;
    float my_w = 0.f;
    if (lane < 8) {
      my_idx = __ldg(&idx[i * 8 + lane]);
      my_w = __ldg(&w[i * 8 + lane]);
      if (my_idx >= m) my_idx = -1;
    }
    int rk[8];
    float wk8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      rk[k] = __shfl_sync(0xffffffffu, my_idx, k);
      wk8[k] = __shfl_sync(0xffffffffu, my_w, k);
    }
    for (int j = lane; j < c; j += 32) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (rk[k] >= 0) acc = __fadd_rn(acc, __fmul_rn(wk8[k], ld_f<TI>(&feats[(int64_t)rk[k] * ld_f_ + j])));
      st_f<TO>(&out[i * ld_o + j], acc);
    }
  }
}
}  // namespace lb

namespace lb {
// 16-bit in / 16-bit out, c % 8 == 0: one thread per (point, 8-channel chunk), 16-byte loads and stores, fp32 math in
// the same corner order.  Corners with weight exactly 0 are not fetched (LiDAL's points sit on voxel corners, so at
// stride 1 seven of the eight weights are 0); the result is identical for finite features.
template <typename T> __device__ __forceinline__ float lo16_to_f32(uint32_t w);
template <typename T> __device__ __forceinline__ float hi16_to_f32(uint32_t w);
template <> __device__ __forceinline__ float lo16_to_f32<__nv_bfloat16>(uint32_t w) { return __uint_as_float(w << 16); }
template <> __device__ __forceinline__ float hi16_to_f32<__nv_bfloat16>(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
template <> __device__ __forceinline__ float lo16_to_f32<__half>(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
template <> __device__ __forceinline__ float hi16_to_f32<__half>(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
template <typename T> __device__ __forceinline__ uint32_t pack16x2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack16x2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack16x2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T>
__global__ void devoxelize16_kernel(const T* __restrict__ feats, int64_t ld_f_, const int* __restrict__ idx,
                                    const float* __restrict__ w, int64_t n, int64_t m, int c, T* __restrict__ out,
                                    int64_t ld_o) {
  const int cpr = c >> 3;
  const int64_t total = n * cpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = t / cpr;
    const int ch = (int)(t - p * cpr);
    const int4 i0 = __ldg((const int4*)&idx[p * 8]), i1 = __ldg((const int4*)&idx[p * 8 + 4]);
    const float4 w0 = __ldg((const float4*)&w[p * 8]), w1 = __ldg((const float4*)&w[p * 8 + 4]);
    const int ik[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {          // all gathers in flight before any is consumed
      const bool use = ik[k] >= 0 && ik[k] < m && wk[k] != 0.f;
      v[k] = use ? __ldg((const uint4*)&feats[(int64_t)ik[k] * ld_f_ + ch * 8]) : make_uint4(0, 0, 0, 0);
    }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ik[k] >= 0 && ik[k] < m && wk[k] != 0.f) {
        const uint32_t wd[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {                 // two 16-bit elements per word; fp32 multiply then add, corner order
          acc[2 * j] = __fadd_rn(acc[2 * j], __fmul_rn(wk[k], lo16_to_f32<T>(wd[j])));
          acc[2 * j + 1] = __fadd_rn(acc[2 * j + 1], __fmul_rn(wk[k], hi16_to_f32<T>(wd[j])));
        }
      }
    }
    uint4 o;
    o.x = pack16x2<T>(acc[0], acc[1]); o.y = pack16x2<T>(acc[2], acc[3]);
    o.z = pack16x2<T>(acc[4], acc[5]); o.w = pack16x2<T>(acc[6], acc[7]);
    *(uint4*)&out[p * ld_o + ch * 8] = o;
  }
}
}  // namespace lb

extern "C" int lb_point_cell_query(const float* pts, int64_t ld, int64_t n, int stride, const void* table,
                                   size_t table_bytes, int32_t* idx, void* stream) {
  LB_CHECK_ARG(n >= 0 && ld >= 4 && stride > 0, "bad sizes");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(pts && table && idx, "null pointer");
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  point_cell_query_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, as_stream(stream)>>>(
      pts, ld, n, stride, table_view(table, table_bytes), idx); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_point_corner_query(const float* pts, int64_t ld, int64_t n, int stride, const void* table,
                                     size_t table_bytes, int32_t* idx, float* w, void* stream) {
  LB_CHECK_ARG(n >= 0 && ld >= 4 && stride > 0, "bad sizes");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(pts && table && idx && w, "null pointer");
  LB_CHECK_ARG((((uintptr_t)idx | (uintptr_t)w) & 15) == 0, "idx / w must be 16-byte aligned");
  int64_t blocks = (n + 127) / 128, cap = (int64_t)sm_count() * 16;
  point_corner_query_kernel<<<(int)(blocks > cap ? cap : blocks), 128, 0, as_stream(stream)>>>(
      pts, ld, n, stride, table_view(table, table_bytes), idx, w); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_voxelize_fwd_ex(const void* feats, int feats_dtype, int64_t ld_f, const int32_t* idx,
                                  const int32_t* counts, int64_t n, int64_t m, int c, float* out, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && c > 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (m > 0) { LB_CHECK_ARG(out, "null out"); LB_CUDA(cudaMemsetAsync(out, 0, (size_t)m * c * 4, st)); }
  if (n == 0 || m == 0) return LB_OK;
  LB_CHECK_ARG(ld_f >= c, "row stride smaller than the channel count");
  LB_CHECK_ARG(feats && idx && counts, "null pointer");
  if (c % 4 == 0 && (((uintptr_t)out) & 15) == 0) {
    const int64_t total = ((n + 15) / 16) * (c / 4), blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 32;
    const int g4 = (int)(blocks > cap ? cap : blocks);
    if (feats_dtype == LB_DT_F32) { voxelize_ex_vec4_kernel<float><<<g4, 256, 0, st>>>((const float*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
    else if (feats_dtype == LB_DT_BF16 && ld_f % 4 == 0 && (((uintptr_t)feats) & 7) == 0) { voxelize16_vec4_kernel<__nv_bfloat16><<<g4, 256, 0, st>>>((const __nv_bfloat16*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
    else if (feats_dtype == LB_DT_F16 && ld_f % 4 == 0 && (((uintptr_t)feats) & 7) == 0) { voxelize16_vec4_kernel<__half><<<g4, 256, 0, st>>>((const __half*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
    else if (feats_dtype == LB_DT_BF16) { voxelize_ex_vec4_kernel<__nv_bfloat16><<<g4, 256, 0, st>>>((const __nv_bfloat16*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
    else if (feats_dtype == LB_DT_F16) { voxelize_ex_vec4_kernel<__half><<<g4, 256, 0, st>>>((const __half*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
    else { set_error("lb_voxelize_fwd_ex: bad dtype"); return LB_EINVAL; }
    LB_LAUNCH_CHECK();
    return LB_OK;
  }
  int g = rows_grid(n, 8);
  if (feats_dtype == LB_DT_F32) { voxelize_ex_kernel<float><<<g, 256, 0, st>>>((const float*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
  else if (feats_dtype == LB_DT_BF16) { voxelize_ex_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
  else if (feats_dtype == LB_DT_F16) { voxelize_ex_kernel<__half><<<g, 256, 0, st>>>((const __half*)feats, ld_f, idx, counts, n, m, c, out); LB_LAUNCHED(1); }
  else { set_error("lb_voxelize_fwd_ex: bad dtype"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename TI>
static int devox_out(const TI* feats, int64_t ld_f, const int32_t* idx, const float* w, int64_t n, int64_t m, int c,
                     void* out, int od, int64_t ld_o, cudaStream_t st) {
  int g = rows_grid(n, 8);
  if (od == LB_DT_F32) { devoxelize_ex_kernel<TI, float><<<g, 256, 0, st>>>(feats, ld_f, idx, w, n, m, c, (float*)out, ld_o); LB_LAUNCHED(1); }
  else if (od == LB_DT_BF16) { devoxelize_ex_kernel<TI, __nv_bfloat16><<<g, 256, 0, st>>>(feats, ld_f, idx, w, n, m, c, (__nv_bfloat16*)out, ld_o); LB_LAUNCHED(1); }
  else if (od == LB_DT_F16) { devoxelize_ex_kernel<TI, __half><<<g, 256, 0, st>>>(feats, ld_f, idx, w, n, m, c, (__half*)out, ld_o); LB_LAUNCHED(1); }
  else { set_error("lb_devoxelize_fwd_ex: bad out dtype"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_devoxelize_fwd_ex(const void* feats, int feats_dtype, int64_t ld_f, const int32_t* idx, const float* w,
                                    int64_t n, int64_t m, int c, void* out, int out_dtype, int64_t ld_o, void* stream) {
  LB_CHECK_ARG(n >= 0 && m >= 0 && c > 0, "bad sizes");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(ld_o >= c && (m == 0 || ld_f >= c), "row stride smaller than the channel count");
  LB_CHECK_ARG(feats && idx && w && out, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (feats_dtype == out_dtype && feats_dtype != LB_DT_F32 && c % 8 == 0 && ld_f % 8 == 0 && ld_o % 8 == 0 &&
      ((((uintptr_t)feats) | ((uintptr_t)out) | ((uintptr_t)idx) | ((uintptr_t)w)) & 15) == 0) {
    const int64_t total = n * (c / 8), blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 32;
    const int g = (int)(blocks > cap ? cap : blocks);
    if (feats_dtype == LB_DT_BF16) { devoxelize16_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)feats, ld_f, idx, w, n, m, c, (__nv_bfloat16*)out, ld_o); LB_LAUNCHED(1); }
    else { devoxelize16_kernel<__half><<<g, 256, 0, st>>>((const __half*)feats, ld_f, idx, w, n, m, c, (__half*)out, ld_o); LB_LAUNCHED(1); }
    LB_LAUNCH_CHECK();
    return LB_OK;
  }
  if (feats_dtype == LB_DT_F32) return devox_out((const float*)feats, ld_f, idx, w, n, m, c, out, out_dtype, ld_o, st);
  if (feats_dtype == LB_DT_BF16) return devox_out((const __nv_bfloat16*)feats, ld_f, idx, w, n, m, c, out, out_dtype, ld_o, st);
  if (feats_dtype == LB_DT_F16) return devox_out((const __half*)feats, ld_f, idx, w, n, m, c, out, out_dtype, ld_o, st);
  set_error("lb_devoxelize_fwd_ex: bad feats dtype");
  return LB_EINVAL;
}

namespace lb {
__global__ void gather_rows16_kernel(const int4* __restrict__ src, const int* __restrict__ idx, int64_t n, int4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(&src[__ldg(&idx[i])]);
}
}  // namespace lb
extern "C" int lb_gather_rows16(const void* src, const int32_t* idx, int64_t n, void* out, void* stream) {
  LB_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(src && idx && out && ((((uintptr_t)src) | ((uintptr_t)out)) & 15) == 0, "null or unaligned pointer");
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  gather_rows16_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, as_stream(stream)>>>((const int4*)src, idx, n, (int4*)out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---------------------------------------------------------------------------------------- F1: score-mode voxelizer
// dataset/sk_dataset.py:143-169 on device.  Step 1: coords_p = raw[:, :3] @ trans_m (float64), feats = (coords_p as f32,
// intensity), coords_p *= scale.  Step 2 (after the caller derived the random shift from min/max): += offset,
// astype(int), pack the (x, y, z) key whose ascending order is np.unique(axis=0)'s lexicographic row order.
namespace lb {
struct Mat3 { double m[9]; double off[3]; };
__global__ void tta_transform_kernel(const float4* __restrict__ raw, int64_t n, const __grid_constant__ Mat3 M, double scale,
                                     double* __restrict__ cp /*[n,3]*/, float4* __restrict__ feats) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 r = __ldg(&raw[i]);
    const double x = r.x, y = r.y, z = r.z;
    double c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)     // row-vector times matrix: sum_k p[k] * m[k][j], accumulated in k order without FMA contraction
      c[j] = __dadd_rn(__dadd_rn(__dmul_rn(x, M.m[j]), __dmul_rn(y, M.m[3 + j])), __dmul_rn(z, M.m[6 + j]));
    feats[i] = make_float4((float)c[0], (float)c[1], (float)c[2], r.w);
    cp[3 * i] = __dmul_rn(c[0], scale);
    cp[3 * i + 1] = __dmul_rn(c[1], scale);
    cp[3 * i + 2] = __dmul_rn(c[2], scale);
  }
}
__global__ void tta_quantize_kernel(const double* __restrict__ cp, int64_t n, const __grid_constant__ Mat3 M, int batch,
                                    int coord_bits, int4* __restrict__ coords, int64_t* __restrict__ keys, int* __restrict__ err) {
  const long long lim = 1LL << coord_bits;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double fx = __dadd_rn(cp[3 * i], M.off[0]), fy = __dadd_rn(cp[3 * i + 1], M.off[1]), fz = __dadd_rn(cp[3 * i + 2], M.off[2]);
    const long long x = (long long)fx, y = (long long)fy, z = (long long)fz;      // astype(int): truncation toward zero
    if (fx < 0 || fy < 0 || fz < 0 || x >= lim || y >= lim || z >= lim) atomicOr(err, 1);   // the reference asserts validity (:160-161)
    coords[i] = make_int4((int)x, (int)y, (int)z, batch);
    keys[i] = (int64_t)(((unsigned long long)x << (2 * coord_bits)) | ((unsigned long long)y << coord_bits) | (unsigned long long)z);
  }
}
}  // namespace lb
extern "C" int lb_tta_transform(const float* raw, int64_t n, const double* trans_m /*[host] 9*/, double scale, double* coords_f64,
                                float* feats, void* stream) {
  LB_CHECK_ARG(n >= 0 && trans_m, "bad arguments");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(raw && coords_f64 && feats && ((((uintptr_t)raw) | ((uintptr_t)feats)) & 15) == 0, "null or unaligned pointer");
  Mat3 M;
  for (int i = 0; i < 9; ++i) M.m[i] = trans_m[i];
  M.off[0] = M.off[1] = M.off[2] = 0;
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  tta_transform_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, as_stream(stream)>>>((const float4*)raw, n, M, scale, coords_f64, (float4*)feats); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_tta_quantize(const double* coords_f64, int64_t n, const double* offset /*[host] 3*/, int batch, int coord_bits,
                               int32_t* coords, int64_t* keys, int32_t* err_flag, void* stream) {
  LB_CHECK_ARG(n >= 0 && offset && coord_bits > 0 && coord_bits <= 20 && err_flag, "bad arguments");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(coords_f64 && coords && keys && (((uintptr_t)coords) & 15) == 0, "null or unaligned pointer");
  Mat3 M;
  for (int i = 0; i < 9; ++i) M.m[i] = 0;
  for (int i = 0; i < 3; ++i) M.off[i] = offset[i];
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  tta_quantize_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, as_stream(stream)>>>(coords_f64, n, M, batch, coord_bits, (int4*)coords, keys, err_flag); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
