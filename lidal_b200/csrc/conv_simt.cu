// Sparse convolution: weight packing, the CUDA-core (SIMT) output-stationary kernel and the lb_conv_fwd dispatcher.
//
// The SIMT kernel serves the shapes the tensor-core path cannot tile (c_in = 4 stem conv, odd channel counts)
// and is the independent cross-check for the tcgen05 kernel in conv_tc.cu: same 16-bit operands, same fp32
// accumulation, different summation order.
#include "common.cuh"

namespace lb {

int conv_tc_supported(int k_vol, int c_in, int c_out, int act_dtype);
int conv_tc_launch(const lb_conv_args& a, cudaStream_t st);
int conv_tc_pack8_supported(int k_vol, int c_in, int c_out, int act_dtype);

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }

// kernel fp32 [k, c_in, c_out] -> packed 16-bit [k, c_out, c_in]  (K-major B operand: c_in contiguous)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, int k, int cin, int cout, T* __restrict__ out) {
  int64_t total = (int64_t)k * cin * cout;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int ci = (int)(t % cin);
    int co = (int)((t / cin) % cout);
    int kk = (int)(t / ((int64_t)cin * cout));
    out[t] = from_f<T>(w[((int64_t)kk * cin + ci) * cout + co]);
  }
}

constexpr int ST_ROWS = 32;      // output rows per block
constexpr int ST_CK = 32;        // c_in chunk
constexpr int ST_THREADS = 256;  // 8 threads per row
constexpr int ST_MAXJ = 32;      // c_out <= 256 -> up to 32 outputs per thread

template <typename T, typename O>
__global__ void __launch_bounds__(ST_THREADS)
conv_simt_kernel(const T* __restrict__ in, int64_t ld_in, int64_t n_in, O* __restrict__ out, int64_t ld_out, int64_t n_out_cap,
                 const int* __restrict__ n_out_dev, const int* __restrict__ nbr, int64_t nbr_ld,
                 const int* __restrict__ out_rows, const T* __restrict__ w, int k_vol, int cin, int cout,
                 const float* __restrict__ scale, const float* __restrict__ shift, const T* __restrict__ res,
                 int64_t ld_res, int relu) {
  extern __shared__ float smem[];
  float* sA = smem;                          // [ST_ROWS][ST_CK + 1]
  float* sW = smem + ST_ROWS * (ST_CK + 1);  // [ST_CK][cout]
  __shared__ int s_nb[ST_ROWS];
  const int64_t n_out = n_out_dev ? (int64_t)*n_out_dev : n_out_cap;
  const int r = threadIdx.x >> 3, jl = threadIdx.x & 7;
  const int nj = (cout + 7) / 8;
  for (int64_t tile = blockIdx.x; tile * ST_ROWS < n_out; tile += gridDim.x) {
    const int64_t o = tile * ST_ROWS + r;
    float acc[ST_MAXJ];
#pragma unroll
    for (int i = 0; i < ST_MAXJ; ++i) acc[i] = 0.f;
    for (int k = 0; k < k_vol; ++k) {
      __syncthreads();
      if (threadIdx.x < ST_ROWS) {
        int64_t oo = tile * ST_ROWS + threadIdx.x;
        int nb = -1;
        if (oo < n_out) nb = nbr ? __ldg(&nbr[(int64_t)k * nbr_ld + oo]) : (int)oo;
        if (nb >= n_in) nb = -1;
        s_nb[threadIdx.x] = nb;
      }
      __syncthreads();
      int any = 0;
      if (threadIdx.x < ST_ROWS) any = s_nb[threadIdx.x] >= 0;
      if (!__syncthreads_or(any)) continue;           // whole tile has no neighbour at this offset
      for (int c0 = 0; c0 < cin; c0 += ST_CK) {
        const int ck = min(ST_CK, cin - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < ST_ROWS * ST_CK; t += ST_THREADS) {
          int rr = t / ST_CK, cc = t % ST_CK;
          int nb = s_nb[rr];
          sA[rr * (ST_CK + 1) + cc] = (nb >= 0 && cc < ck) ? to_f<T>(in[(int64_t)nb * ld_in + c0 + cc]) : 0.f;
        }
        for (int t = threadIdx.x; t < ck * cout; t += ST_THREADS) {
          int co = t / ck, cc = t % ck;                // c_in contiguous in the packed weight
          sW[cc * cout + co] = to_f<T>(w[((int64_t)k * cout + co) * cin + c0 + cc]);
        }
        __syncthreads();
        if (s_nb[r] >= 0) {
          for (int cc = 0; cc < ck; ++cc) {
            float a = sA[r * (ST_CK + 1) + cc];
#pragma unroll
            for (int i = 0; i < ST_MAXJ; ++i)
              if (i < nj) {
                int j = jl + 8 * i;
                if (j < cout) acc[i] = fmaf(a, sW[cc * cout + j], acc[i]);
              }
          }
        }
      }
    }
    if (o < n_out) {
      const int64_t orow = out_rows ? (int64_t)__ldg(&out_rows[o]) : o;
#pragma unroll
      for (int i = 0; i < ST_MAXJ; ++i)
        if (i < nj) {
          int j = jl + 8 * i;
          if (j < cout) {
            float v = acc[i];
            if (scale) v *= __ldg(&scale[j]);
            if (shift) v += __ldg(&shift[j]);
            if (relu == 2) v = fmaxf(v, 0.f);
            if (res) v += to_f<T>(res[orow * ld_res + j]);
            if (relu == 1) v = fmaxf(v, 0.f);
            out[orow * ld_out + j] = from_f<O>(v);
          }
        }
    }
  }
}

template <typename T, typename O>
static int launch_simt(const lb_conv_args& a, cudaStream_t st) {
  size_t smem = (size_t)(ST_ROWS * (ST_CK + 1) + ST_CK * a.c_out) * sizeof(float);
  int64_t tiles = (a.n_out + ST_ROWS - 1) / ST_ROWS;
  int64_t cap = (int64_t)sm_count() * 8;
  int grid = (int)(tiles > cap ? cap : (tiles < 1 ? 1 : tiles));
  conv_simt_kernel<T, O><<<grid, ST_THREADS, smem, st>>>(
      (const T*)a.in, a.ld_in, a.n_in, (O*)a.out, a.ld_out, a.n_out, a.n_out_dev, a.nbr, a.nbr_ld, a.out_rows,
      (const T*)a.weight, a.k_vol, a.c_in, a.c_out, a.scale, a.shift, (const T*)a.residual, a.ld_res,
      (a.flags & LB_CONV_RELU) ? ((a.flags & LB_CONV_RELU_FIRST) ? 2 : 1) : 0); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename T>
static int launch_simt_out(const lb_conv_args& a, cudaStream_t st) {
  if (a.out_dtype == LB_DT_F32) return launch_simt<T, float>(a, st);
  if (a.out_dtype == a.act_dtype) return launch_simt<T, T>(a, st);
  set_error("lb_conv_fwd: out_dtype must be F32 or equal act_dtype");
  return LB_EINVAL;
}

}  // namespace lb
using namespace lb;

extern "C" int lb_conv_pack_weight(const float* kernel, int k, int cin, int cout, int dt, void* packed, void* stream) {
  LB_CHECK_ARG(kernel && packed, "null pointer");
  LB_CHECK_ARG(k > 0 && cin > 0 && cout > 0, "bad shape");
  int64_t total = (int64_t)k * cin * cout, blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  int g = (int)(blocks > cap ? cap : blocks);
  if (dt == LB_DT_BF16) { pack_weight_kernel<<<g, 256, 0, as_stream(stream)>>>(kernel, k, cin, cout, (__nv_bfloat16*)packed); LB_LAUNCHED(1); }
  else if (dt == LB_DT_F16) { pack_weight_kernel<<<g, 256, 0, as_stream(stream)>>>(kernel, k, cin, cout, (__half*)packed); LB_LAUNCHED(1); }
  else { set_error("lb_conv_pack_weight: act_dtype must be BF16 or F16"); return LB_EINVAL; }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_conv_sched_ws_bytes(void) { return 256; }
extern "C" int lb_conv_uses_tensor_cores(int k_vol, int c_in, int c_out, int act_dtype) {
  return conv_tc_supported(k_vol, c_in, c_out, act_dtype);
}

extern "C" int lb_conv_fwd(const lb_conv_args* a, void* stream) {
  LB_CHECK_ARG(a, "null args");
  LB_CHECK_ARG(a->n_in >= 0 && a->n_out >= 0, "negative row count");
  if (a->n_out == 0) return LB_OK;                  // empty outputs: nothing to do (strides of empty tensors are arbitrary)
  LB_CHECK_ARG(a->in && a->out && a->weight, "null tensor");
  LB_CHECK_ARG(a->k_vol >= 1 && a->c_in >= 1 && a->c_out >= 1 && a->c_out <= 256, "bad k_vol/c_in/c_out (c_out <= 256)");
  LB_CHECK_ARG(a->ld_in >= a->c_in && a->ld_out >= a->c_out, "row stride smaller than channel count");
  LB_CHECK_ARG(a->nbr || a->k_vol == 1, "nbr may be NULL only for k_vol == 1");
  LB_CHECK_ARG(!a->nbr || a->nbr_ld >= a->n_out, "nbr_ld < n_out");
  LB_CHECK_ARG(a->act_dtype == LB_DT_BF16 || a->act_dtype == LB_DT_F16, "act_dtype must be BF16 or F16");
  LB_CHECK_ARG(!a->residual || a->ld_res >= a->c_out, "ld_res < c_out");
  cudaStream_t st = as_stream(stream);
  if (a->flags & LB_CONV_PACK8) {
    LB_CHECK_ARG(conv_tc_pack8_supported(a->k_vol, a->c_in, a->c_out, a->act_dtype) && (a->ld_in % 8 == 0) &&
                     (((uintptr_t)a->in) & 15) == 0 && a->nbr,
                 "LB_CONV_PACK8 needs c_in == 8, 16-byte rows, c_out % 32 == 0 and a neighbour table");
    return conv_tc_launch(*a, st);
  }
  if (!(a->flags & LB_CONV_FORCE_SIMT) && conv_tc_supported(a->k_vol, a->c_in, a->c_out, a->act_dtype) &&
      (a->ld_in % 8 == 0) && (((uintptr_t)a->in) & 15) == 0)
    return conv_tc_launch(*a, st);
  if (a->act_dtype == LB_DT_BF16) return launch_simt_out<__nv_bfloat16>(*a, st);
  return launch_simt_out<__half>(*a, st);
}
