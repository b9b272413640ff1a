// Weight gradient of the sparse convolution on the 5th-generation tensor cores (training):
//
//   gW[k][ci][co] = sum over the pairs (i, o) of offset k   x[i][ci] * g[o][co]
//
// Per offset this is a tall-skinny contraction  D[Cin x Cout] += X_k^T [Cin x P] * G_k [P x Cout]  whose reduction dimension
// is the PAIR index.  Both operands are gathered rows, and a gathered row tile [pairs][64 channels] is exactly the
// MN-major 128B-swizzled UMMA layout (64 contiguous M/N elements per 128-byte line, 8-line swizzle atoms along K), so the
// forward kernel's gather machinery is reused unchanged: 8 gather warps cp.async the x rows (A) and g rows (B) of 128
// pairs per stage, one thread issues eight tcgen05.mma (M = 128 input channels, N = Cout, K = 16 pairs each), the
// accumulator stays in TMEM across a chunk of pair tiles (split-K) and 4 epilogue warps flush it with float4 atomics.
#include "tc_ptx.cuh"

namespace lb {

constexpr int WG_PAIRS = 128;          // pairs per pipeline stage
constexpr int WG_M = 128;              // input channels per accumulator tile (UMMA M)
constexpr int WG_EPI_THREADS = 128;
constexpr int WG_PROD_THREADS = 256;
constexpr int WG_THREADS = WG_EPI_THREADS + WG_PROD_THREADS + 32;
constexpr int WG_MMA_WARP = (WG_EPI_THREADS + WG_PROD_THREADS) / 32;
constexpr int WG_MAX_STAGES = 4;
constexpr int WG_MAX_K = 27;
constexpr int WG_PANEL_BYTES = WG_PAIRS * 128;   // [128 pairs][64 channels] 16-bit

struct WgParams {
  const char* x; int64_t n_x, ld_x;
  const char* g; int64_t n_g, ld_g;
  const int2* pairs;                   // (x_row, g_row), offset-major
  int pair_begin[WG_MAX_K + 1];
  int item_begin[WG_MAX_K + 1];        // work items (chunk, m-tile) per offset, prefix sum
  int k_vol, c_in, c_out, n_mt, chunk_tiles, b_panels;
  float* gw;
  int is_bf16, stages, tmem_cols, n_acc;
};

// MN-major swizzled descriptor: LBO = stride between 64-element M/N panels, SBO = stride between 8-row K groups
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_kernel(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_bytes = 2 * WG_PANEL_BYTES;                       // channels [mt*128, +128) = two 64-channel panels
  const int stage_bytes = a_bytes + p.b_panels * WG_PANEL_BYTES;
  uint8_t* ring = smem;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = (uint64_t*)tail;
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + WG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* s_flags = (uint32_t*)(tempty_bar + 2);
  uint32_t* s_tmem = s_flags + WG_MAX_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = p.item_begin[p.k_vol];

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], WG_PROD_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], WG_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == WG_MMA_WARP) tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // item -> (offset k, m-tile, pair-tile range)
  auto decode = [&](int item, int& k, int& mt, int& t0, int& t1) {
    k = 0;
    while (item >= p.item_begin[k + 1]) ++k;
    const int local = item - p.item_begin[k];
    mt = local % p.n_mt;
    const int chunk = local / p.n_mt;
    const int tiles = (p.pair_begin[k + 1] - p.pair_begin[k] + WG_PAIRS - 1) / WG_PAIRS;
    t0 = chunk * p.chunk_tiles;
    t1 = min(t0 + p.chunk_tiles, tiles);
  };

  if (warp >= 4 && warp < WG_MMA_WARP) {
    // =============================================================== GATHER WARPS
    const int t = threadIdx.x - WG_EPI_THREADS;
    const int chunk = t & 7, row0 = t >> 3;             // 8 x 16-byte chunks per 128-byte line, 32 rows per pass
    int stage = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int k, mt, t0, t1;
      decode(item, k, mt, t0, t1);
      const int pb = p.pair_begin[k], pe = p.pair_begin[k + 1];
      for (int tile = t0; tile < t1; ++tile) {
        mbar_wait(&empty_bar[stage], ph ^ 1);
        uint8_t* st_base = ring + (size_t)stage * stage_bytes;
        if (t == 0) s_flags[stage] = (tile == t0 ? 1u : 0u) | (tile == t1 - 1 ? 2u : 0u);
#pragma unroll
        for (int i = 0; i < WG_PAIRS / 32; ++i) {
          const int r = row0 + 32 * i;
          const int pi = pb + tile * WG_PAIRS + r;
          int2 pr = make_int2(-1, -1);
          if (pi < pe) pr = __ldg(&p.pairs[pi]);
          if (pr.x >= p.n_x || pr.y >= p.n_g) pr = make_int2(-1, -1);
          const bool ok = pr.x >= 0 && pr.y >= 0;
          const uint32_t line = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) * 16));
          // A: two panels of 64 input channels
#pragma unroll
          for (int sp = 0; sp < 2; ++sp) {
            const int ch = mt * WG_M + sp * 64 + chunk * 8;
            const bool v = ok && ch < p.c_in;
            cp_async16(smem_u32(st_base + sp * WG_PANEL_BYTES) + line, p.x + ((int64_t)(v ? pr.x : 0) * p.ld_x + (v ? ch : 0)) * 2,
                       v ? 16u : 0u);
          }
          // B: ceil(c_out / 64) panels of 64 output channels
          for (int sp = 0; sp < p.b_panels; ++sp) {
            const int ch = sp * 64 + chunk * 8;
            const bool v = ok && ch < p.c_out;
            cp_async16(smem_u32(st_base + a_bytes + sp * WG_PANEL_BYTES) + line,
                       p.g + ((int64_t)(v ? pr.y : 0) * p.ld_g + (v ? ch : 0)) * 2, v ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(&full_bar[stage]);
        if (++stage == p.stages) { stage = 0; ph ^= 1; }
      }
    }
    cp_async_wait<0>();
  } else if (warp == WG_MMA_WARP) {
    // =============================================================== MMA ISSUER
    const uint32_t idesc = make_idesc_mn(WG_M, p.c_out, p.is_bf16 ? 1 : 0);
    int stage = 0;
    uint32_t ph = 0;
    int64_t icount = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++icount) {
      const int acc = (int)(icount % p.n_acc);
      const uint32_t acc_ph = (uint32_t)((icount / p.n_acc) & 1);
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.c_out);
      while (true) {
        mbar_wait(&full_bar[stage], ph);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t flags = s_flags[stage];
        if (lane == 0) {
          const uint32_t st_u32 = smem_u32(ring + (size_t)stage * stage_bytes);
#pragma unroll
          for (int kk = 0; kk < WG_PAIRS / 16; ++kk) {      // 16 pairs = two 8-line swizzle atoms = 2048 bytes along K
            const uint64_t a_desc = make_smem_desc_mn(st_u32 + kk * 2048, WG_PANEL_BYTES, 1024);
            const uint64_t b_desc = make_smem_desc_mn(st_u32 + a_bytes + kk * 2048, WG_PANEL_BYTES, 1024);
            umma_f16(d_tmem, a_desc, b_desc, idesc, ((flags & 1u) && kk == 0) ? 0u : 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (flags & 2u) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; ph ^= 1; }
        if (flags & 2u) break;
      }
    }
  } else {
    // =============================================================== EPILOGUE: TMEM -> float4 atomics into gW[k]
    int64_t icount = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++icount) {
      int k, mt, t0, t1;
      decode(item, k, mt, t0, t1);
      const int acc = (int)(icount % p.n_acc);
      const uint32_t acc_ph = (uint32_t)((icount / p.n_acc) & 1);
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const int ci = mt * WG_M + warp * 32 + lane;
      float* dst = p.gw + ((int64_t)k * p.c_in + (ci < p.c_in ? ci : 0)) * p.c_out;
      for (int c0 = 0; c0 < p.c_out; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * p.c_out + c0), v);
        if (ci < p.c_in) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (c0 + q * 4 < p.c_out)
              atomicAdd((float4*)(dst + c0 + q * 4), make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                                 __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace lb
using namespace lb;

extern "C" int lb_conv_wgrad(const void* x, int64_t n_x, int64_t ld_x, const void* g, int64_t n_g, int64_t ld_g,
                             const int32_t* pairs, const int32_t* pair_begin, int k_vol, int c_in, int c_out, int act_dtype,
                             float* grad_kernel, void* stream) {
  LB_CHECK_ARG(k_vol >= 1 && k_vol <= WG_MAX_K && pair_begin, "k_vol must be in [1,27]");
  LB_CHECK_ARG(c_in >= 8 && c_in % 8 == 0, "c_in must be a multiple of 8 (pad the input rows)");
  LB_CHECK_ARG(c_out >= 32 && c_out % 32 == 0 && c_out <= 256, "c_out must be a multiple of 32, <= 256");
  LB_CHECK_ARG(act_dtype == LB_DT_BF16 || act_dtype == LB_DT_F16, "act_dtype must be BF16 or F16");
  LB_CHECK_ARG(grad_kernel && (((uintptr_t)grad_kernel) & 15) == 0, "grad_kernel must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  LB_CUDA(cudaMemsetAsync(grad_kernel, 0, (size_t)k_vol * c_in * c_out * 4, st));
  const int total_pairs = pair_begin[k_vol];
  if (total_pairs == 0) return LB_OK;
  LB_CHECK_ARG(x && g && pairs, "null pointer");
  LB_CHECK_ARG(ld_x % 8 == 0 && ld_g % 8 == 0 && ld_x >= c_in && ld_g >= c_out && ((((uintptr_t)x) | ((uintptr_t)g)) & 15) == 0 &&
                   (((uintptr_t)pairs) & 7) == 0,
               "rows must be 16-byte aligned");
  WgParams p;
  p.x = (const char*)x; p.n_x = n_x; p.ld_x = ld_x;
  p.g = (const char*)g; p.n_g = n_g; p.ld_g = ld_g;
  p.pairs = (const int2*)pairs;
  p.k_vol = k_vol; p.c_in = c_in; p.c_out = c_out;
  p.n_mt = (c_in + WG_M - 1) / WG_M;
  p.b_panels = (c_out + 63) / 64;
  p.gw = grad_kernel;
  p.is_bf16 = act_dtype == LB_DT_BF16;
  int64_t total_tiles = 0;
  for (int k = 0; k < k_vol; ++k) {
    LB_CHECK_ARG(pair_begin[k + 1] >= pair_begin[k], "pair_begin must be non-decreasing");
    total_tiles += (pair_begin[k + 1] - pair_begin[k] + WG_PAIRS - 1) / WG_PAIRS;
  }
  // split-K chunk: enough work items to fill the machine ~4x, but long enough to amortise the atomic flush
  int64_t chunk = total_tiles * p.n_mt / ((int64_t)4 * sm_count());
  p.chunk_tiles = (int)(chunk < 1 ? 1 : (chunk > 64 ? 64 : chunk));
  for (int k = 0; k <= WG_MAX_K; ++k) p.pair_begin[k] = p.item_begin[k] = 0;
  int items = 0;
  for (int k = 0; k < k_vol; ++k) {
    p.pair_begin[k] = pair_begin[k];
    p.item_begin[k] = items;
    const int tiles = (pair_begin[k + 1] - pair_begin[k] + WG_PAIRS - 1) / WG_PAIRS;
    items += ((tiles + p.chunk_tiles - 1) / p.chunk_tiles) * p.n_mt;
  }
  p.pair_begin[k_vol] = pair_begin[k_vol];
  p.item_begin[k_vol] = items;
  p.n_acc = (2 * c_out <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.n_acc * c_out) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage_bytes = (size_t)(2 + p.b_panels) * WG_PANEL_BYTES;
  const size_t tail = 512;
  int stages = (int)((227 * 1024 - 1024 - tail) / stage_bytes);
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  if (stages < 2) { set_error("lb_conv_wgrad: shared memory too small"); return LB_ECAP; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + tail + 1024;
  int grid = items < sm_count() ? items : sm_count();
  if (grid < 1) grid = 1;
  LB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(p); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
