// Coordinate hashing, hash table, radix sort / scan primitives, strided downsample and kernel-map build.
// Integer-only work: every result here is bit-exact against the oracle.
// All kernels are HBM/L2-latency bound; grids are sized in multiples of the SM count.
#include <stdarg.h>
#include "common.cuh"

namespace lb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void add_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
static inline int grid_for(int64_t n, int block, int per_thread = 1) {
  int64_t blocks = (n + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------ hash
__global__ void hash_kernel(const int4* __restrict__ coords, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(&coords[i]);
    out[i] = fnv60(c.x, c.y, c.z, c.w);
  }
}
__global__ void kernel_hash_kernel(const int4* __restrict__ coords, int64_t n, const int* __restrict__ offsets, int k,
                                   int64_t* __restrict__ out) {
  extern __shared__ int s_off[];
  for (int i = threadIdx.x; i < 3 * k; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(&coords[i]);
    for (int j = 0; j < k; ++j)   // coalesced across threads for every j
      out[(int64_t)j * n + i] = fnv60(c.x + s_off[3 * j], c.y + s_off[3 * j + 1], c.z + s_off[3 * j + 2], c.w);
  }
}

// ------------------------------------------------------------------------------------------ table
__global__ void table_build_kernel(const int64_t* __restrict__ keys, int64_t n, TableView t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    table_insert(t, (uint64_t)__ldg(&keys[i]), (int)i);
}
__global__ void table_query_kernel(TableView t, const int64_t* __restrict__ q, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int64_t)table_find(t, (uint64_t)__ldg(&q[i]));
}

// ------------------------------------------------------------------------------------------ scan
// Three-phase exclusive scan: per-tile sums -> single-block scan of tile sums -> per-tile scan + offset.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    s_warp[lane] = wi - w;                 // exclusive warp offsets
    if (lane == 31) s_warp[32] = wi;       // total
  }
  __syncthreads();
  uint32_t res = s_warp[warp] + incl - v;
  block_total = s_warp[32];
  __syncthreads();
  return res;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t* __restrict__ in, int64_t n,
                                                               uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_warp[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j * SCAN_THREADS + threadIdx.x;
    if (i < n) acc += __ldg(&in[i]);
  }
  uint32_t total;
  block_exclusive_scan(acc, s_warp, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) scan_sums_single(uint32_t* sums, int64_t m, uint32_t* total_out) {
  __shared__ uint32_t s_warp[33];
  uint32_t carry = 0;
  for (int64_t base = 0; base < m; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    uint32_t v = i < m ? sums[i] : 0, tot;
    uint32_t ex = block_exclusive_scan(v, s_warp, tot);
    if (i < m) sums[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                int64_t n, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_warp[33];
  // each thread owns SCAN_ITEMS consecutive elements so the scan order is the memory order
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], acc = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? __ldg(&in[base + j]) : 0;
    acc += v[j];
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan(acc, s_warp, tot) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = ex;
    ex += v[j];
  }
}
size_t scan_ws_bytes(int64_t n) { return (size_t)(ceil_div(n, SCAN_TILE) + 1) * 4 + 256; }
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total, void* ws, cudaStream_t st) {
  if (n <= 0) {
    if (total) LB_CUDA(cudaMemsetAsync(total, 0, 4, st));
    return LB_OK;
  }
  int tiles = ceil_div(n, SCAN_TILE);
  uint32_t* sums = (uint32_t*)ws;
  scan_tile_sums<<<tiles, SCAN_THREADS, 0, st>>>(in, n, sums); LB_LAUNCHED(1);
  scan_sums_single<<<1, 1024, 0, st>>>(sums, tiles, total); LB_LAUNCHED(1);
  scan_tile_apply<<<tiles, SCAN_THREADS, 0, st>>>(in, out, n, sums); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ radix sort
// Stable LSD radix sort, 8 bits per pass.  Each block owns a contiguous tile; each warp a contiguous
// sub-range of it, ranked with match_any so equal digits keep their input order.
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PER_WARP = 512;                 // keys per warp per tile
constexpr int RS_TILE = RS_WARPS * RS_PER_WARP;  // 4096 keys per block

__global__ void __launch_bounds__(RS_THREADS) rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                      uint32_t* __restrict__ hist /*[256][tiles]*/, int tiles) {
  __shared__ uint32_t s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
  for (int j = threadIdx.x; j < RS_TILE; j += RS_THREADS) {
    int64_t i = base + j;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = s_h[threadIdx.x];
}
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const uint64_t* __restrict__ keys_in,
                                                         const uint32_t* __restrict__ vals_in,
                                                         uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                         int64_t n, int shift, const uint32_t* __restrict__ hist_scan,
                                                         int tiles) {
  __shared__ uint32_t s_cnt[RS_WARPS][256];   // per-warp digit counts, then per-warp running offsets
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_THREADS) (&s_cnt[0][0])[j] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_PER_WARP;
  // pass A: per-warp digit histogram
  for (int j = lane; j < RS_PER_WARP; j += 32) {
    int64_t i = wbase + j;
    if (i < n) atomicAdd(&s_cnt[warp][(keys_in[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  // exclusive prefix over warps per digit + global base of this tile
  {
    int d = threadIdx.x;   // RS_THREADS == 256 digits
    uint32_t run = hist_scan[(int64_t)d * tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = s_cnt[w][d];
      s_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass B: stable ranking inside each warp's sub-range, 32 keys at a time
  for (int j0 = 0; j0 < RS_PER_WARP; j0 += 32) {
    int64_t i = wbase + j0 + lane;
    bool ok = i < n;
    uint64_t key = ok ? keys_in[i] : 0;
    uint32_t digit = ok ? (uint32_t)((key >> shift) & 255) : 256u + lane;   // inactive lanes never match
    uint32_t peers = __match_any_sync(0xffffffffu, digit);
    uint32_t rank = __popc(peers & ((1u << lane) - 1));
    uint32_t pos = 0;
    if (ok) pos = s_cnt[warp][digit] + rank;
    __syncwarp();
    if (ok && rank == __popc(peers) - 1) s_cnt[warp][digit] = pos + 1;   // last peer advances the counter
    __syncwarp();
    if (ok) {
      keys_out[pos] = key;
      if (vals_in) vals_out[pos] = vals_in[i];
    }
  }
}
// ---- single-sweep variant: one kernel per pass.  Digit totals of every pass come from one up-front histogram of the
// unsorted keys (digit counts do not depend on order); the per-tile prefix inside a pass is resolved by decoupled
// look-back over a status word per (tile, digit) = flag (2 bits: 0 empty, 1 tile aggregate, 2 inclusive prefix) | count
// (30 bits).  Tiles take their index from an atomic ticket so a tile only ever waits on tiles that already started.
constexpr int RS_MAX_PASSES = 8;
constexpr int RS_LOOKBACK = 8;                   // predecessors polled per round (independent loads, one L2 latency)
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(RS_THREADS) rs_digit_totals(const uint64_t* __restrict__ keys, int64_t n, int passes,
                                                              uint32_t* __restrict__ totals /*[passes][256]*/) {
  __shared__ uint32_t s_h[RS_MAX_PASSES][256];
  for (int j = threadIdx.x; j < passes * 256; j += RS_THREADS) (&s_h[0][0])[j] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
    const uint64_t key = keys[i];
    for (int ps = 0; ps < passes; ++ps) atomicAdd(&s_h[ps][(key >> (8 * ps)) & 255], 1u);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < passes * 256; j += RS_THREADS) {
    const uint32_t c = (&s_h[0][0])[j];
    if (c) atomicAdd(&totals[j], c);
  }
}

__global__ void __launch_bounds__(RS_THREADS) rs_sweep(const uint64_t* __restrict__ keys_in,
                                                       const uint32_t* __restrict__ vals_in,
                                                       uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                       int64_t n, int shift, const uint32_t* __restrict__ totals /*[256]*/,
                                                       uint32_t* __restrict__ status /*[tiles][256]*/,
                                                       uint32_t* __restrict__ ticket) {
  constexpr int ITEMS = RS_PER_WARP / 32;        // keys per lane, kept in registers across the look-back
  __shared__ uint32_t s_cnt[RS_WARPS][256];
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_THREADS) (&s_cnt[0][0])[j] = 0;
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t wbase = tile * RS_TILE + (int64_t)warp * RS_PER_WARP;
  uint64_t key[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    key[j] = i < n ? keys_in[i] : 0;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (wbase + j * 32 + lane < n) atomicAdd(&s_cnt[warp][(key[j] >> shift) & 255], 1u);
  __syncthreads();
  {
    const int d = threadIdx.x;   // RS_THREADS == 256 digits
    uint32_t mine = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) mine += s_cnt[w][d];
    uint32_t* my_status = status + tile * 256 + d;
    uint32_t excl = 0;
    if (tile == 0) {
      st_relaxed_gpu(my_status, RS_FLAG_PREFIX | mine);
    } else {
      st_relaxed_gpu(my_status, RS_FLAG_AGG | mine);
      int64_t t = tile - 1;
      bool done = false;
      while (!done) {
        uint32_t v[RS_LOOKBACK];
#pragma unroll
        for (int j = 0; j < RS_LOOKBACK; ++j)
          v[j] = (t - j >= 0) ? ld_relaxed_gpu(status + (t - j) * 256 + d) : RS_FLAG_PREFIX;
        int used = 0;
        bool stop = false;
#pragma unroll
        for (int j = 0; j < RS_LOOKBACK; ++j) {
          if (!stop) {
            const uint32_t f = v[j] & ~RS_VAL_MASK;
            if (f == 0) {
              stop = true;                         // not published yet: poll again from this tile
            } else {
              excl += v[j] & RS_VAL_MASK;
              ++used;
              if (f == RS_FLAG_PREFIX) { done = true; stop = true; }
            }
          }
        }
        t -= used;
      }
      st_relaxed_gpu(my_status, RS_FLAG_PREFIX | (excl + mine));
    }
    uint32_t all;
    uint32_t run = block_exclusive_scan(totals[d], s_warp, all) + excl;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = s_cnt[w][d];
      s_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // stable ranking inside each warp's sub-range, 32 keys at a time (match_any keeps equal digits in input order)
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    const bool ok = i < n;
    const uint32_t digit = ok ? (uint32_t)((key[j] >> shift) & 255) : 256u + lane;   // inactive lanes never match
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1));
    uint32_t pos = 0;
    if (ok) pos = s_cnt[warp][digit] + rank;
    __syncwarp();
    if (ok && rank == __popc(peers) - 1) s_cnt[warp][digit] = pos + 1;
    __syncwarp();
    if (ok) {
      keys_out[pos] = key[j];
      if (vals_in) vals_out[pos] = __ldg(&vals_in[i]);
    }
  }
}
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static bool rs_use_sweep(int64_t n) {
  static const bool off = getenv("LIDAL_SORT_3PHASE") != nullptr;    // A/B switch: the older hist/scan/scatter passes
  // beyond ~512 tiles the first wave's look-back chain outweighs the saved histogram/scan launches (measured, 4M keys)
  return !off && n <= (int64_t)512 * RS_TILE;
}
static size_t rs_sweep_bytes(int tiles) {    // ticket[8] + totals[8][256] + status[8][tiles][256]
  return align256(256 + (size_t)RS_MAX_PASSES * 256 * 4 + (size_t)RS_MAX_PASSES * tiles * 256 * 4);
}

}  // namespace lb

using namespace lb;

extern "C" size_t lb_sort_pairs_ws_bytes(int64_t n) {
  int tiles = ceil_div(n > 0 ? n : 1, RS_TILE);
  return align256((size_t)n * 8) + align256((size_t)n * 4) + align256((size_t)256 * tiles * 4) +
         align256(scan_ws_bytes((int64_t)256 * tiles)) + rs_sweep_bytes(tiles) + 1024;
}
extern "C" int lb_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int end_bit, void* ws, size_t ws_bytes,
                             void* stream) {
  LB_CHECK_ARG(keys && ws, "null pointer");
  LB_CHECK_ARG(end_bit > 0 && end_bit <= 64, "end_bit out of range");
  if (ws_bytes < lb_sort_pairs_ws_bytes(n)) { set_error("lb_sort_pairs: workspace too small"); return LB_ECAP; }
  if (n <= 1) return LB_OK;
  cudaStream_t st = as_stream(stream);
  int tiles = ceil_div(n, RS_TILE);
  char* p = (char*)ws;
  uint64_t* k2 = (uint64_t*)p; p += align256((size_t)n * 8);
  uint32_t* v2 = (uint32_t*)p; p += align256((size_t)n * 4);
  uint32_t* hist = (uint32_t*)p; p += align256((size_t)256 * tiles * 4);
  void* sws = p; p += align256(scan_ws_bytes((int64_t)256 * tiles));
  uint32_t* ticket = (uint32_t*)p;                       // [8] (padded to 256 bytes)
  uint32_t* totals = (uint32_t*)(p + 256);               // [passes][256]
  uint32_t* status = totals + RS_MAX_PASSES * 256;       // [passes][tiles][256]
  int passes = (end_bit + 7) / 8;
  uint64_t *ka = keys, *kb = k2;
  uint32_t *va = vals, *vb = v2;
  const bool sweep = rs_use_sweep(n);
  if (sweep) {
    LB_CUDA(cudaMemsetAsync(p, 0, 256 + (size_t)RS_MAX_PASSES * 256 * 4 + (size_t)passes * tiles * 256 * 4, st));
    int grid = tiles < 148 * 4 ? tiles : 148 * 4;
    rs_digit_totals<<<grid, RS_THREADS, 0, st>>>(keys, n, passes, totals); LB_LAUNCHED(1);
  }
  for (int ps = 0; ps < passes; ++ps) {
    if (sweep) {
      rs_sweep<<<tiles, RS_THREADS, 0, st>>>(ka, vals ? va : nullptr, kb, vb, n, ps * 8, totals + ps * 256,
                                             status + (size_t)ps * tiles * 256, ticket + ps); LB_LAUNCHED(1);
    } else {
      rs_hist<<<tiles, RS_THREADS, 0, st>>>(ka, n, ps * 8, hist, tiles); LB_LAUNCHED(1);
      int rc = exclusive_scan_u32(hist, hist, (int64_t)256 * tiles, nullptr, sws, st);
      if (rc != LB_OK) return rc;
      rs_scatter<<<tiles, RS_THREADS, 0, st>>>(ka, vals ? va : nullptr, kb, vb, n, ps * 8, hist, tiles); LB_LAUNCHED(1);
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  if (ka != keys) {
    LB_CUDA(cudaMemcpyAsync(keys, ka, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    if (vals) LB_CUDA(cudaMemcpyAsync(vals, va, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ C ABI: hash + table
extern "C" int lb_abi_version(void) { return LB_ABI_VERSION; }
extern "C" uint64_t lb_launch_count(void) { return lb::launches(); }
extern "C" const char* lb_last_error(void) { return lb::g_err; }
extern "C" int lb_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  LB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  LB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return LB_OK;
}
extern "C" int lb_hash(const int32_t* coords, int64_t n, int64_t* out, void* stream) {
  LB_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(coords && out, "null pointer");
  LB_CHECK_ARG(((uintptr_t)coords & 15) == 0, "coords must be 16-byte aligned");
  hash_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>((const int4*)coords, n, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_kernel_hash(const int32_t* coords, int64_t n, const int32_t* offsets, int k, int64_t* out,
                              void* stream) {
  LB_CHECK_ARG(n >= 0 && k > 0 && k <= 1024, "bad n or k");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(coords && offsets && out, "null pointer");
  LB_CHECK_ARG(((uintptr_t)coords & 15) == 0, "coords must be 16-byte aligned");
  kernel_hash_kernel<<<grid_for(n, 256), 256, 3 * k * sizeof(int), as_stream(stream)>>>((const int4*)coords, n, offsets,
                                                                                       k, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" size_t lb_hashtable_bytes(int64_t n) { return (size_t)table_capacity(n > 0 ? n : 1) * 12; }
extern "C" int lb_hashtable_build(const int64_t* keys, int64_t n, void* table, size_t bytes, void* stream) {
  LB_CHECK_ARG(table && n >= 0, "null table or n < 0");
  if (bytes < lb_hashtable_bytes(n) || (bytes / 12) & (bytes / 12 - 1)) {
    set_error("lb_hashtable_build: table_bytes must be lb_hashtable_bytes(n) (12 * power of two)");
    return LB_ECAP;
  }
  cudaStream_t st = as_stream(stream);
  uint64_t cap = bytes / 12;
  LB_CUDA(cudaMemsetAsync(table, 0xFF, cap * 8, st));
  LB_CUDA(cudaMemsetAsync((char*)table + cap * 8, 0x7F, cap * 4, st));
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(keys, "null keys");
  table_build_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, n, table_view(table, bytes)); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_hashtable_query(const void* table, size_t bytes, const int64_t* q, int64_t nq, int64_t* out,
                                  void* stream) {
  LB_CHECK_ARG(table && nq >= 0, "null table or nq < 0");
  if (nq == 0) return LB_OK;
  LB_CHECK_ARG(q && out, "null pointer");
  table_query_kernel<<<grid_for(nq, 256), 256, 0, as_stream(stream)>>>(table_view(table, bytes), q, nq, out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ downsample
namespace lb {
// key = (b << 48) | (x << 32) | (y << 16) | z  -- ascending key order == (b,x,y,z) lexicographic order.
__global__ void ds_pack(const int4* __restrict__ coords, int64_t n, int sx, int sy, int sz, uint64_t* __restrict__ keys,
                        int* __restrict__ err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(&coords[i]);
    if ((c.x | c.y | c.z) < 0 || (c.x | c.y | c.z) > 65535 || c.w < 0 || c.w > 32767) atomicOr(err, 1);
    uint32_t x = (uint32_t)(c.x / sx) * sx, y = (uint32_t)(c.y / sy) * sy, z = (uint32_t)(c.z / sz) * sz;
    keys[i] = ((uint64_t)(uint32_t)c.w << 48) | ((uint64_t)(x & 65535) << 32) | ((uint64_t)(y & 65535) << 16) |
              (uint64_t)(z & 65535);
  }
}
__global__ void ds_flags(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void ds_emit(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ flags,
                        const uint32_t* __restrict__ pos, int64_t n, int4* __restrict__ out,
                        const int* __restrict__ err, int* __restrict__ n_out) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && *err) *n_out = -1;   // out-of-range coordinate: reported as n_out = -1
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (flags[i]) {
      uint64_t k = keys[i];
      out[pos[i]] = make_int4((int)((k >> 32) & 65535), (int)((k >> 16) & 65535), (int)(k & 65535), (int)(k >> 48));
    }
  }
}
}  // namespace lb

extern "C" size_t lb_downsample_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return align256((size_t)n * 8) + 2 * align256((size_t)n * 4) + align256(scan_ws_bytes(n)) +
         align256(lb_sort_pairs_ws_bytes(n)) + 512;
}
extern "C" int lb_downsample(const int32_t* coords, int64_t n, const int32_t ss[3], int batch_bits, int32_t* out_coords,
                             int32_t* n_out, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && ss && n_out && ws, "null pointer or n < 0");
  LB_CHECK_ARG(ss[0] > 0 && ss[1] > 0 && ss[2] > 0, "sample stride must be positive");
  LB_CHECK_ARG(batch_bits >= 1 && batch_bits <= 15, "batch_bits in [1,15]");
  if (ws_bytes < lb_downsample_ws_bytes(n)) { set_error("lb_downsample: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (n == 0) { LB_CUDA(cudaMemsetAsync(n_out, 0, 4, st)); return LB_OK; }
  LB_CHECK_ARG(coords && out_coords, "null pointer");
  char* p = (char*)ws;
  uint64_t* keys = (uint64_t*)p; p += align256((size_t)n * 8);
  uint32_t* flags = (uint32_t*)p; p += align256((size_t)n * 4);
  uint32_t* pos = (uint32_t*)p; p += align256((size_t)n * 4);
  void* scan_ws = p; p += align256(scan_ws_bytes(n));
  void* sort_ws = p; p += align256(lb_sort_pairs_ws_bytes(n));
  int* err = (int*)p;
  LB_CUDA(cudaMemsetAsync(err, 0, 4, st));
  int g = grid_for(n, 256);
  ds_pack<<<g, 256, 0, st>>>((const int4*)coords, n, ss[0], ss[1], ss[2], keys, err); LB_LAUNCHED(1);
  int rc = lb_sort_pairs(keys, nullptr, n, 48 + batch_bits, sort_ws, lb_sort_pairs_ws_bytes(n), stream);
  if (rc != LB_OK) return rc;
  ds_flags<<<g, 256, 0, st>>>(keys, n, flags); LB_LAUNCHED(1);
  rc = exclusive_scan_u32(flags, pos, n, (uint32_t*)n_out, scan_ws, st);
  if (rc != LB_OK) return rc;
  ds_emit<<<g, 256, 0, st>>>(keys, flags, pos, n, (int4*)out_coords, err, n_out); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ kernel map
namespace lb {
__global__ void kmap_query_kernel(TableView t, const int4* __restrict__ out_coords, int64_t cap,
                                  const int* __restrict__ n_dev, const int* __restrict__ offsets, int k,
                                  int* __restrict__ nbr) {
  extern __shared__ int s_off[];
  for (int i = threadIdx.x; i < 3 * k; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  const int64_t n = n_dev ? (int64_t)*n_dev : cap;
  // one thread per (offset, row): consecutive threads walk consecutive rows -> coalesced nbr writes
  const int64_t total = cap * k;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(idx / cap);
    int64_t o = idx - (int64_t)j * cap;
    int r = -1;
    if (o < n) {
      int4 c = __ldg(&out_coords[o]);
      r = table_find(t, (uint64_t)fnv60(c.x + s_off[3 * j], c.y + s_off[3 * j + 1], c.z + s_off[3 * j + 2], c.w));
    }
    nbr[idx] = r;
  }
}
// Symmetric (submanifold) maps: in == out coordinates and offsets[k-1-j] == -offsets[j], so nbr[j][o] = r implies
// nbr[k-1-j][r] = o.  Only the first (k+1)/2 offsets are probed; the mirrored half (pre-filled with -1) is scattered.
__global__ void kmap_query_sym_kernel(TableView t, const int4* __restrict__ coords, int64_t n, const int* __restrict__ offsets,
                                      int k, int* __restrict__ nbr, int64_t ld) {
  extern __shared__ int s_off[];
  for (int i = threadIdx.x; i < 3 * k; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  const int half = (k + 1) / 2;
  const int64_t total = n * half;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx / n);
    const int64_t o = idx - (int64_t)j * n;
    const int4 c = __ldg(&coords[o]);
    const int r = table_find(t, (uint64_t)fnv60(c.x + s_off[3 * j], c.y + s_off[3 * j + 1], c.z + s_off[3 * j + 2], c.w));
    nbr[(int64_t)j * ld + o] = r;
    if (r >= 0 && j != k - 1 - j) nbr[(int64_t)(k - 1 - j) * ld + r] = (int)o;
  }
}
__global__ void kmap_flags(const int* __restrict__ nbr, int64_t total, uint32_t* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = nbr[i] >= 0 ? 1u : 0u;
}
__global__ void kmap_emit(const int* __restrict__ nbr, const uint32_t* __restrict__ pos, int64_t n_out, int k,
                          int2* __restrict__ nbmaps, int* __restrict__ nbsizes, const uint32_t* __restrict__ total) {
  const int64_t tot = n_out * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    int r = nbr[i];
    if (r >= 0) nbmaps[pos[i]] = make_int2(r, (int)(i % n_out));
    if (i % n_out == 0) {
      int j = (int)(i / n_out);
      uint32_t end = (j + 1 < k) ? pos[(int64_t)(j + 1) * n_out] : *total;
      nbsizes[j] = (int)(end - pos[i]);
    }
  }
}
}  // namespace lb

extern "C" int lb_kmap_query(const void* table, size_t bytes, const int32_t* out_coords, int64_t cap,
                             const int32_t* n_dev, const int32_t* offsets, int k, int32_t* nbr, void* stream) {
  LB_CHECK_ARG(cap >= 0 && k > 0 && k <= 1024, "bad sizes");
  if (cap == 0) return LB_OK;
  LB_CHECK_ARG(table && out_coords && offsets && nbr, "null pointer");
  kmap_query_kernel<<<grid_for(cap * k, 256), 256, 3 * k * sizeof(int), as_stream(stream)>>>(
      table_view(table, bytes), (const int4*)out_coords, cap, n_dev, offsets, k, nbr); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_kmap_query_sym(const void* table, size_t bytes, const int32_t* coords, int64_t n, const int32_t* offsets,
                                 int k, int32_t* nbr, int64_t nbr_ld, void* stream) {
  LB_CHECK_ARG(n >= 0 && k > 0 && (k & 1) && k <= 343 && nbr_ld >= n, "bad arguments (k must be odd)");
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(table && coords && offsets && nbr, "null pointer");
  cudaStream_t st = as_stream(stream);
  const int half = (k + 1) / 2;
  LB_CUDA(cudaMemsetAsync(nbr + (int64_t)half * nbr_ld, 0xFF, (size_t)(k - half) * nbr_ld * 4, st));
  kmap_query_sym_kernel<<<grid_for(n * half, 256), 256, 3 * k * sizeof(int), st>>>(table_view(table, bytes), (const int4*)coords, n,
                                                                              offsets, k, nbr, nbr_ld); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" size_t lb_kmap_compact_ws_bytes(int64_t n_out, int k) {
  int64_t t = (n_out > 0 ? n_out : 1) * (int64_t)k;
  return 2 * align256((size_t)t * 4) + align256(scan_ws_bytes(t)) + 256;
}
extern "C" int lb_kmap_compact(const int32_t* nbr, int64_t n_out, int k, int32_t* nbmaps, int32_t* nbsizes,
                               int32_t* total, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n_out >= 0 && k > 0 && nbsizes && total && ws, "bad arguments");
  if (ws_bytes < lb_kmap_compact_ws_bytes(n_out, k)) { set_error("lb_kmap_compact: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (n_out == 0) {
    LB_CUDA(cudaMemsetAsync(nbsizes, 0, (size_t)k * 4, st));
    LB_CUDA(cudaMemsetAsync(total, 0, 4, st));
    return LB_OK;
  }
  LB_CHECK_ARG(nbr && nbmaps, "null pointer");
  int64_t t = n_out * k;
  char* p = (char*)ws;
  uint32_t* flags = (uint32_t*)p; p += align256((size_t)t * 4);
  uint32_t* pos = (uint32_t*)p; p += align256((size_t)t * 4);
  void* sws = p;
  int g = grid_for(t, 256);
  kmap_flags<<<g, 256, 0, st>>>(nbr, t, flags); LB_LAUNCHED(1);
  int rc = exclusive_scan_u32(flags, pos, t, (uint32_t*)total, sws, st);
  if (rc != LB_OK) return rc;
  kmap_emit<<<g, 256, 0, st>>>(nbr, pos, n_out, k, (int2*)nbmaps, nbsizes, (const uint32_t*)total); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

namespace lb {
__global__ void kmap_transpose_kernel(const int* __restrict__ nbr, int64_t nbr_ld, int64_t n_out, int k,
                                      int* __restrict__ nbr_t, int64_t n_in) {
  const int64_t total = n_out * k;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(t / n_out);
    int64_t o = t - (int64_t)j * n_out;
    int i = __ldg(&nbr[(int64_t)j * nbr_ld + o]);
    if (i >= 0 && i < n_in) nbr_t[(int64_t)j * n_in + i] = (int)o;
  }
}
}  // namespace lb
extern "C" int lb_kmap_transpose(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* nbr_t, int64_t n_in,
                                 void* stream) {
  LB_CHECK_ARG(n_out >= 0 && n_in >= 0 && k > 0 && nbr_ld >= n_out, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (n_in > 0) { LB_CHECK_ARG(nbr_t, "null nbr_t"); LB_CUDA(cudaMemsetAsync(nbr_t, 0xFF, (size_t)n_in * k * 4, st)); }
  if (n_out == 0 || n_in == 0) return LB_OK;
  LB_CHECK_ARG(nbr, "null nbr");
  kmap_transpose_kernel<<<grid_for(n_out * k, 256), 256, 0, st>>>(nbr, nbr_ld, n_out, k, nbr_t, n_in); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ unique with inverse
namespace lb {
__global__ void uq_iota(uint32_t* v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}
__global__ void uq_emit(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                        const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, int64_t n,
                        int64_t* __restrict__ uniq, int* __restrict__ inverse, int* __restrict__ first_row) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t id = pos[i] + flags[i] - 1;          // exclusive scan of flags -> id of the run this element belongs to
    if (flags[i]) {
      uniq[id] = (int64_t)keys[i];
      if (first_row) first_row[id] = (int)vals[i];   // stable sort: the run's first element is the first occurrence
    }
    if (inverse) inverse[vals[i]] = (int)id;
  }
}
}  // namespace lb
extern "C" size_t lb_unique_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return align256((size_t)n * 8) + 3 * align256((size_t)n * 4) + align256(scan_ws_bytes(n)) +
         align256(lb_sort_pairs_ws_bytes(n)) + 256;
}
extern "C" int lb_unique_i64(const int64_t* keys, int64_t n, int key_bits, int64_t* uniq, int32_t* n_unique,
                             int32_t* inverse, int32_t* first_row, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && n_unique && ws && key_bits > 0 && key_bits <= 64, "bad arguments");
  if (ws_bytes < lb_unique_ws_bytes(n)) { set_error("lb_unique_i64: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (n == 0) { LB_CUDA(cudaMemsetAsync(n_unique, 0, 4, st)); return LB_OK; }
  LB_CHECK_ARG(keys && uniq, "null pointer");
  char* p = (char*)ws;
  uint64_t* k = (uint64_t*)p; p += align256((size_t)n * 8);
  uint32_t* v = (uint32_t*)p; p += align256((size_t)n * 4);
  uint32_t* flags = (uint32_t*)p; p += align256((size_t)n * 4);
  uint32_t* pos = (uint32_t*)p; p += align256((size_t)n * 4);
  void* scan_ws = p; p += align256(scan_ws_bytes(n));
  void* sort_ws = p;
  LB_CUDA(cudaMemcpyAsync(k, keys, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  int g = grid_for(n, 256);
  uq_iota<<<g, 256, 0, st>>>(v, n); LB_LAUNCHED(1);
  int rc = lb_sort_pairs(k, v, n, key_bits, sort_ws, lb_sort_pairs_ws_bytes(n), stream);
  if (rc != LB_OK) return rc;
  ds_flags<<<g, 256, 0, st>>>(k, n, flags); LB_LAUNCHED(1);
  rc = exclusive_scan_u32(flags, pos, n, (uint32_t*)n_unique, scan_ws, st);
  if (rc != LB_OK) return rc;
  uq_emit<<<g, 256, 0, st>>>(k, v, flags, pos, n, uniq, inverse, first_row); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ mask-sorted kernel maps
namespace lb {
// per-offset neighbour counts (how many rows have offset j)
__global__ void ks_count(const int* __restrict__ nbr, int64_t ld, int64_t n, int k, unsigned* __restrict__ counts,
                         uint32_t* __restrict__ masks) {
  __shared__ unsigned s_cnt[32];
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n_pad = (n + 31) & ~(int64_t)31;          // whole warps stay converged for the ballots
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_pad; o += (int64_t)gridDim.x * blockDim.x) {
    uint32_t m = 0;
    for (int j = 0; j < k; ++j) {
      const bool v = o < n && __ldg(&nbr[(int64_t)j * ld + o]) >= 0;
      const unsigned b = __ballot_sync(0xffffffffu, v);
      if (lane == 0 && b) atomicAdd(&s_cnt[j], __popc(b));
      m |= (uint32_t)v << j;
    }
    if (o < n) masks[o] = m;                               // bit j = offset j present
  }
  __syncthreads();
  if (threadIdx.x < k && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}
// bit position of offset j inside the sort key: the LEAST frequent offset becomes the most significant bit, so the
// many rows that lack the rare offsets form long runs and 128/256-row tiles share more of their mask
// (measured: 17-30 % fewer active (tile, offset) blocks than sorting by the plain mask value).
__global__ void ks_bitpos(const unsigned* __restrict__ counts, int k, int* __restrict__ bitpos) {
  const int j = threadIdx.x;
  if (j >= k) return;
  int rank = 0;
  for (int i = 0; i < k; ++i)
    if (counts[i] < counts[j] || (counts[i] == counts[j] && i < j)) ++rank;
  bitpos[j] = k - 1 - rank;
}
// `drop`: the most frequent offsets (lowest key bits) are left out of the key - nearly every tile needs them anyway, and a
// 27-offset map then sorts in three 8-bit passes instead of four.
// `chunk_shift` > 0: rows are first grouped into chunks of 2^chunk_shift consecutive rows (key = chunk : mask) -- when the
// row order is spatially coherent (scan order), the tiles in flight then gather from a slice of the input that stays in
// L2 instead of every mask group sweeping the whole feature matrix.
__global__ void ks_keys(const uint32_t* __restrict__ masks, int64_t n, int k, const int* __restrict__ bitpos, int drop,
                        int chunk_shift, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  __shared__ int s_pos[32];
  if (threadIdx.x < 32) s_pos[threadIdx.x] = threadIdx.x < k ? bitpos[threadIdx.x] : 0;
  __syncthreads();
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t nat = masks[o];
    uint64_t m = 0;
    for (int j = 0; j < k; ++j) m |= (uint64_t)((nat >> j) & 1u) << s_pos[j];
    m >>= drop;
    if (chunk_shift > 0) m |= (uint64_t)(o >> chunk_shift) << (k - drop);
    keys[o] = m;
    vals[o] = (uint32_t)o;
  }
}
__global__ void ks_permute(const int* __restrict__ nbr, int64_t ld, int64_t n, int k, const int* __restrict__ perm,
                           int* __restrict__ out, int64_t out_ld) {
  const int64_t total = n * k;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(t / n);
    int64_t i = t - (int64_t)j * n;
    out[(int64_t)j * out_ld + i] = __ldg(&nbr[(int64_t)j * ld + __ldg(&perm[i])]);
  }
}
// tile mask g = OR of the row masks of sorted rows [128 g, 128 g + 128): one warp per group, 4 rows per lane
__global__ void ks_tile_masks(const uint32_t* __restrict__ masks, const int* __restrict__ perm, int64_t n, int64_t groups,
                              uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < groups; g += warps) {
    uint32_t m = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = g * 128 + u * 32 + lane;
      if (i < n) m |= __ldg(&masks[__ldg(&perm[i])]);
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if (lane == 0) out[g] = m;
  }
}
}  // namespace lb
extern "C" size_t lb_kmap_sort_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return align256((size_t)n * 8) + align256(lb_sort_pairs_ws_bytes(n)) + 512 + align256((size_t)n * 4);
}
extern "C" int lb_kmap_sort_by_mask(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm,
                                    int32_t* nbr_sorted, void* ws, size_t ws_bytes, void* stream) {
  return lb_kmap_sort_by_mask_ld(nbr, nbr_ld, n_out, k, perm, nbr_sorted, n_out, ws, ws_bytes, stream);
}
extern "C" int lb_kmap_sort_by_mask_tm(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm,
                                       int32_t* nbr_sorted, int64_t sorted_ld, uint32_t* tile_masks, void* ws, size_t ws_bytes,
                                       void* stream);
extern "C" int lb_kmap_sort_by_mask_ld(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm,
                                       int32_t* nbr_sorted, int64_t sorted_ld, void* ws, size_t ws_bytes, void* stream) {
  return lb_kmap_sort_by_mask_tm(nbr, nbr_ld, n_out, k, perm, nbr_sorted, sorted_ld, nullptr, ws, ws_bytes, stream);
}
extern "C" int lb_kmap_sort_by_mask_tm(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm,
                                       int32_t* nbr_sorted, int64_t sorted_ld, uint32_t* tile_masks, void* ws, size_t ws_bytes,
                                       void* stream) {
  LB_CHECK_ARG(n_out >= 0 && k > 0 && k <= 32 && nbr_ld >= n_out && sorted_ld >= n_out && ws, "bad arguments");
  if (ws_bytes < lb_kmap_sort_ws_bytes(n_out)) { set_error("lb_kmap_sort_by_mask: workspace too small"); return LB_ECAP; }
  if (n_out == 0) return LB_OK;
  LB_CHECK_ARG(nbr && perm && nbr_sorted, "null pointer");
  cudaStream_t st = as_stream(stream);
  uint64_t* keys = (uint64_t*)ws;
  void* sort_ws = (char*)ws + align256((size_t)n_out * 8);
  unsigned* counts = (unsigned*)((char*)sort_ws + align256(lb_sort_pairs_ws_bytes(n_out)));   // [32] + bitpos [32]
  int* bitpos = (int*)(counts + 32);
  uint32_t* masks = (uint32_t*)((char*)counts + 512);
  LB_CUDA(cudaMemsetAsync(counts, 0, 256, st));
  ks_count<<<grid_for(n_out, 256), 256, 0, st>>>(nbr, nbr_ld, n_out, k, counts, masks); LB_LAUNCHED(1);
  ks_bitpos<<<1, 32, 0, st>>>(counts, k, bitpos); LB_LAUNCHED(1);
  static const int key_bits = getenv("LIDAL_MASK_KEY_BITS") ? atoi(getenv("LIDAL_MASK_KEY_BITS")) : LB_MASK_KEY_BITS;
  const int drop = (key_bits > 0 && k > key_bits) ? k - key_bits : 0;
  static const int chunk_shift = getenv("LIDAL_MASK_CHUNK_SHIFT") ? atoi(getenv("LIDAL_MASK_CHUNK_SHIFT")) : LB_MASK_CHUNK_SHIFT;
  int chunk_bits = 0;
  if (chunk_shift > 0)
    while (((n_out - 1) >> chunk_shift) >> chunk_bits) ++chunk_bits;
  ks_keys<<<grid_for(n_out, 256), 256, 0, st>>>(masks, n_out, k, bitpos, drop, chunk_bits ? chunk_shift : 0, keys, (uint32_t*)perm); LB_LAUNCHED(1);
  int rc = lb_sort_pairs(keys, (uint32_t*)perm, n_out, k - drop + chunk_bits, sort_ws, lb_sort_pairs_ws_bytes(n_out), stream);
  if (rc != LB_OK) return rc;
  ks_permute<<<grid_for(n_out * k, 256), 256, 0, st>>>(nbr, nbr_ld, n_out, k, perm, nbr_sorted, sorted_ld); LB_LAUNCHED(1);
  if (tile_masks) {          // the row masks are still in the workspace: 4 bytes per row instead of re-reading the table
    const int64_t groups = (n_out + 127) / 128;
    ks_tile_masks<<<grid_for(groups * 32, 256), 256, 0, st>>>(masks, perm, n_out, groups, tile_masks); LB_LAUNCHED(1);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ per-tile offset masks
namespace lb {
// one warp per 128-row group: lane l owns rows 4 l .. 4 l + 3 (one 16-byte load per offset when the table rows are aligned)
__global__ void kmap_tile_masks_kernel(const int* __restrict__ nbr, int64_t ld, int64_t n, int k, uint32_t* __restrict__ out,
                                       int64_t groups, int vec_ok) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < groups; g += warps) {
    const int64_t o = g * 128 + lane * 4;
    uint32_t m = 0;
    for (int j = 0; j < k; ++j) {
      const int* src = nbr + (int64_t)j * ld + o;
      bool any = false;
      if (vec_ok && o + 3 < n) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(src));
        any = (v.x >= 0) | (v.y >= 0) | (v.z >= 0) | (v.w >= 0);
      } else {
        for (int u = 0; u < 4; ++u)
          if (o + u < n && __ldg(src + u) >= 0) any = true;
      }
      m |= (uint32_t)any << j;
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if (lane == 0) out[g] = m;
  }
}
}  // namespace lb
extern "C" int lb_kmap_tile_masks(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, uint32_t* tile_masks, void* stream) {
  LB_CHECK_ARG(n_out >= 0 && k > 0 && k <= 32 && nbr_ld >= n_out, "bad arguments");
  if (n_out == 0) return LB_OK;
  LB_CHECK_ARG(nbr && tile_masks, "null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t groups = (n_out + 127) / 128;
  const int vec_ok = ((nbr_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(nbr) & 15) == 0) ? 1 : 0;
  kmap_tile_masks_kernel<<<grid_for(groups * 32, 256), 256, 0, st>>>(nbr, nbr_ld, n_out, k, tile_masks, groups, vec_ok); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ hash group-by (engine)
// Groups equal keys without sorting: the group of a key is owned by its smallest row (atomicMin in the table), groups
// are numbered in order of their owner row (exclusive scan) -> deterministic, first-occurrence order.
namespace lb {
__global__ void gb_flags(TableView t, const int64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags,
                         int* __restrict__ owner) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int o = table_find(t, (uint64_t)__ldg(&keys[i]));
    owner[i] = o;
    flags[i] = (o == (int)i) ? 1u : 0u;
  }
}
__global__ void gb_emit(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, const int* __restrict__ owner,
                        int64_t n, int* __restrict__ inverse, int* __restrict__ first_row) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    inverse[i] = (int)pos[owner[i]];
    if (flags[i] && first_row) first_row[pos[i]] = (int)i;
  }
}
__global__ void dsm_keys(const int4* __restrict__ coords, int64_t n, int s2, int64_t* __restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(&coords[i]);
    keys[i] = fnv60(c.x / s2 * s2, c.y / s2 * s2, c.z / s2 * s2, c.w);
  }
}
// parent coordinates + both stride-2 maps from the child -> parent assignment (no hash queries)
__global__ void dsm_emit(const int4* __restrict__ coords, const int* __restrict__ inverse, const uint32_t* __restrict__ flags,
                         int64_t n, int ts, int4* __restrict__ out_coords, int* __restrict__ nbr_dn, int64_t ld_dn,
                         int* __restrict__ nbr_up) {
  const int s2 = 2 * ts;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = __ldg(&coords[i]);
    const int g = inverse[i];
    const int px = c.x / s2 * s2, py = c.y / s2 * s2, pz = c.z / s2 * s2;
    const int k = ((c.x - px) / ts) * 4 + ((c.y - py) / ts) * 2 + (c.z - pz) / ts;     // even-kernel offset order: x slowest
    if (flags[i]) out_coords[g] = make_int4(px, py, pz, c.w);
    nbr_dn[(int64_t)k * ld_dn + g] = (int)i;
    nbr_up[(int64_t)k * n + i] = g;
  }
}
}  // namespace lb

extern "C" size_t lb_group_by_key_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return align256(lb_hashtable_bytes(n)) + 3 * align256((size_t)n * 4) + align256(scan_ws_bytes(n)) + 256;
}
extern "C" int lb_group_by_key(const int64_t* keys, int64_t n, int32_t* inverse, int32_t* first_row, int32_t* n_groups,
                               void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && n_groups && ws, "bad arguments");
  if (ws_bytes < lb_group_by_key_ws_bytes(n)) { set_error("lb_group_by_key: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (n == 0) { LB_CUDA(cudaMemsetAsync(n_groups, 0, 4, st)); return LB_OK; }
  LB_CHECK_ARG(keys && inverse, "null pointer");
  char* p = (char*)ws;
  void* table = p; const size_t tb = lb_hashtable_bytes(n); p += align256(tb);
  uint32_t* flags = (uint32_t*)p; p += align256((size_t)n * 4);
  uint32_t* pos = (uint32_t*)p; p += align256((size_t)n * 4);
  int* owner = (int*)p; p += align256((size_t)n * 4);
  void* scan_ws = p;
  int rc = lb_hashtable_build(keys, n, table, tb, stream);
  if (rc != LB_OK) return rc;
  int g = grid_for(n, 256);
  gb_flags<<<g, 256, 0, st>>>(table_view(table, tb), keys, n, flags, owner); LB_LAUNCHED(1);
  rc = exclusive_scan_u32(flags, pos, n, (uint32_t*)n_groups, scan_ws, st);
  if (rc != LB_OK) return rc;
  gb_emit<<<g, 256, 0, st>>>(flags, pos, owner, n, inverse, first_row); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------ row counts of all levels
// How many voxels will every coarser level have?  unique (coords / 2^l * 2^l) for l = 1..levels, counted by inserting the
// SAME 60-bit hashes lb_downsample_maps groups by into one key-only table per level.  It lets a caller learn every row
// count of the pyramid in ONE host round trip (before building anything) instead of one round trip per level.
namespace lb {
__global__ void level_counts_kernel(const int4* __restrict__ coords, int64_t n, int levels, unsigned long long* __restrict__ tables,
                                    uint64_t cap, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t n_pad = (n + 31) & ~(int64_t)31;            // whole warps stay converged for the ballots
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = i < n ? __ldg(&coords[i]) : make_int4(0, 0, 0, 0);
    // hierarchical: only the row that inserted a cell at level l - 1 carries it to level l (all rows of a cell share their
    // parent), so the atomics shrink with the pyramid: n + n_1 + n_2 + ... instead of levels * n
    bool alive = i < n;
    for (int l = 1; l <= levels; ++l) {
      const int s = 1 << l;
      bool fresh = false;
      if (alive) {
        const unsigned long long key = (unsigned long long)fnv60(c.x / s * s, c.y / s * s, c.z / s * s, c.w);
        unsigned long long* t = tables + (uint64_t)(l - 1) * cap;
        uint64_t slot = mix64(key) & (cap - 1);
        while (true) {
          const unsigned long long prev = atomicCAS(&t[slot], LB_EMPTY_KEY, key);
          if (prev == LB_EMPTY_KEY) { fresh = true; break; }
          if (prev == key) break;
          slot = (slot + 1) & (cap - 1);
        }
      }
      const unsigned b = __ballot_sync(0xffffffffu, fresh);
      if (lane == 0 && b) atomicAdd(&counts[l - 1], __popc(b));
      alive = fresh;
    }
  }
}
__global__ void level_counts_publish(const int* __restrict__ dev_counts, int levels, int* __restrict__ out) {
  if ((int)threadIdx.x < levels) out[threadIdx.x] = dev_counts[threadIdx.x];
}
}  // namespace lb
extern "C" size_t lb_level_counts_ws_bytes(int64_t n, int levels) {
  if (n < 1) n = 1;
  if (levels < 1) levels = 1;
  return (size_t)levels * table_capacity(n) * 8 + 256;
}
extern "C" int lb_level_counts(const int32_t* coords, int64_t n, int levels, int32_t* counts, void* ws, size_t ws_bytes,
                               void* stream) {
  LB_CHECK_ARG(n >= 0 && levels >= 1 && levels <= 16 && counts && ws, "bad arguments");
  if (ws_bytes < lb_level_counts_ws_bytes(n, levels)) { set_error("lb_level_counts: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  const uint64_t cap = table_capacity(n);
  int* dev_counts = (int*)ws;
  unsigned long long* tables = (unsigned long long*)((char*)ws + 256);
  LB_CUDA(cudaMemsetAsync(dev_counts, 0, 256, st));
  if (n > 0) {
    LB_CHECK_ARG(coords, "null coords");
    LB_CUDA(cudaMemsetAsync(tables, 0xFF, (size_t)levels * cap * 8, st));
    level_counts_kernel<<<grid_for(n, 256), 256, 0, st>>>((const int4*)coords, n, levels, tables, cap, dev_counts); LB_LAUNCHED(1);
  }
  level_counts_publish<<<1, 32, 0, st>>>(dev_counts, levels, counts); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_downsample_maps_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return align256((size_t)n * 8) + align256((size_t)n * 4) + align256(lb_group_by_key_ws_bytes(n)) + 256;
}
extern "C" int lb_downsample_maps(const int32_t* coords, int64_t n, int tensor_stride, int32_t* out_coords, int32_t* n_out,
                                  int32_t* nbr_dn, int64_t ld_dn, int32_t* nbr_up, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && tensor_stride > 0 && n_out && ws && ld_dn >= n, "bad arguments");
  if (ws_bytes < lb_downsample_maps_ws_bytes(n)) { set_error("lb_downsample_maps: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  if (n == 0) { LB_CUDA(cudaMemsetAsync(n_out, 0, 4, st)); return LB_OK; }
  LB_CHECK_ARG(coords && out_coords && nbr_dn && nbr_up, "null pointer");
  char* p = (char*)ws;
  int64_t* keys = (int64_t*)p; p += align256((size_t)n * 8);
  int* inverse = (int*)p; p += align256((size_t)n * 4);
  void* gws = p;
  int g = grid_for(n, 256);
  dsm_keys<<<g, 256, 0, st>>>((const int4*)coords, n, 2 * tensor_stride, keys); LB_LAUNCHED(1);
  int rc = lb_group_by_key(keys, n, inverse, nullptr, n_out, gws, lb_group_by_key_ws_bytes(n), stream);
  if (rc != LB_OK) return rc;
  LB_CUDA(cudaMemsetAsync(nbr_dn, 0xFF, (size_t)8 * ld_dn * 4, st));
  LB_CUDA(cudaMemsetAsync(nbr_up, 0xFF, (size_t)8 * n * 4, st));
  // flags live at a fixed place inside the group-by workspace: [table][flags]...
  const uint32_t* flags = (const uint32_t*)((char*)gws + align256(lb_hashtable_bytes(n)));
  dsm_emit<<<g, 256, 0, st>>>((const int4*)coords, inverse, flags, n, tensor_stride, (int4*)out_coords, nbr_dn, ld_dn, nbr_up); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
