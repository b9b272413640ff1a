// LiDAL inter-frame uncertainty scoring on device (SURVEY.md section 8 rows A9, A10, A11-device part).
//
// Restates score/sv_level/LiDAL.py:59-103.  The pickled float64 KD-tree of the reference is replaced by a
// per-frame uniform hash grid, but the match rule is unchanged: the exact float64 nearest neighbour of
// every query point in each neighbouring frame, accepted iff sqrt(d2) <= dis_thresh.  All floating-point
// steps follow the reference's evaluation order and precisions so scores are reproducible to the last bit
// wherever libm agrees:
//   kl_div / entr : evaluated in double, rounded to float  (scipy ufunc loops d_dd__As_ff_f)
//   np.sum(axis=1) over classes: float32 pairwise order of numpy (8 accumulators, tree, sequential tail)
//   interd accumulates in double; sum_prob in float, neighbours in nei_ids order.
// One warp owns one query point: lanes 0..26 probe the 27 surrounding cells, then lane c owns class c.
#include "common.cuh"
#include "np_sum.cuh"

namespace lb {

// Per-frame uniform grid, cell-sorted:
//   header | slots[cap] {key, start, count} (open addressing on the packed cell key) | xyz_sorted f64 [n,3] | orig i32 [n]
//          | rel f32 [n,3] (coordinates relative to the point's own cell origin: |rel| < cell, so float32 keeps ~1e-8 m)
// Points are radix-sorted by cell key, so the candidates of a cell are one contiguous run and consecutive points are
// spatial neighbours (probes of consecutive query points hit the same cache lines).
struct GridHeader {
  double cell;
  int64_t n;
  uint64_t cap;
  uint64_t pad;
};
struct GridSlot {
  unsigned long long key;
  int start;
  int count;
};
struct GridView {
  double cell;
  uint64_t mask;
  int64_t n;
  const GridSlot* slots;
  const double* xyz;     // cell-sorted coordinates
  const int* orig;       // cell-sorted position -> original row
  const float* rel;      // cell-sorted coordinates minus their cell origin, float32 (prefilter operand)
};
__host__ __device__ inline size_t grid_bytes_for(int64_t n) {
  uint64_t cap = table_capacity(n);
  return sizeof(GridHeader) + cap * sizeof(GridSlot) + (size_t)(n > 0 ? n : 1) * 40 + 64;
}
__host__ __device__ __forceinline__ GridView grid_view(const void* g) {
  const GridHeader* h = (const GridHeader*)g;
  GridView v;
  v.cell = h->cell;
  v.mask = h->cap - 1;
  v.n = h->n;
  v.slots = (const GridSlot*)((const char*)g + sizeof(GridHeader));
  v.xyz = (const double*)(v.slots + h->cap);
  v.orig = (const int*)(v.xyz + 3 * (h->n > 0 ? h->n : 1));
  v.rel = (const float*)(v.orig + (h->n > 0 ? h->n : 1));
  return v;
}
constexpr long long CELL_BIAS = 1 << 20;
// (every user -- grid build, cell-relative coordinates, queries -- goes through this one function, so cell membership is
// consistent by construction; a multiply by the reciprocal replaces three double divisions per query)
__device__ __forceinline__ long long cell_of(double x, double cell) {
  long long c = (long long)floor(x * (1.0 / cell));
  return c < -CELL_BIAS + 2 ? -CELL_BIAS + 2 : (c > CELL_BIAS - 2 ? CELL_BIAS - 2 : c);
}
__device__ __forceinline__ uint64_t cell_key(long long ix, long long iy, long long iz) {
  return ((uint64_t)(ix + CELL_BIAS) << 42) | ((uint64_t)(iy + CELL_BIAS) << 21) | (uint64_t)(iz + CELL_BIAS);
}
// slot index of a packed cell key: 32-bit murmur finaliser over the folded key (cheaper than the 64-bit mix)
__device__ __forceinline__ uint64_t cell_slot(uint64_t key) {
  uint32_t h = (uint32_t)key ^ (uint32_t)(key >> 21) ^ (uint32_t)(key >> 42) * 0x9E3779B1u;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return (uint64_t)h;
}
// slot lookup: (start, count) of a cell or count = 0
__device__ __forceinline__ int2 grid_find(const GridView& g, uint64_t key) {
  uint64_t slot = cell_slot(key) & g.mask;
  while (true) {
    const uint4 raw = __ldg((const uint4*)&g.slots[slot]);          // one 16-byte load per probe
    const unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
    if (k == key) return make_int2((int)raw.z, (int)raw.w);
    if (k == LB_EMPTY_KEY) return make_int2(0, 0);
    slot = (slot + 1) & g.mask;
  }
}

__global__ void grid_init_kernel(void* grid, double cell, int64_t n, uint64_t cap) {
  GridHeader* h = (GridHeader*)grid;
  if (blockIdx.x == 0 && threadIdx.x == 0) { h->cell = cell; h->n = n; h->cap = cap; h->pad = 0; }
  GridSlot* slots = (GridSlot*)((char*)grid + sizeof(GridHeader));
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    slots[i].key = LB_EMPTY_KEY;
    slots[i].start = 0;
    slots[i].count = 0;
  }
}
__global__ void grid_keys_kernel(const double* __restrict__ xyz, int64_t n, double cell, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = cell_key(cell_of(xyz[3 * i], cell), cell_of(xyz[3 * i + 1], cell), cell_of(xyz[3 * i + 2], cell));
    vals[i] = (uint32_t)i;
  }
}
// after the sort: copy coordinates in sorted order, register every run [start, end) of equal keys in the slot table
__global__ void grid_fill_kernel(const double* __restrict__ xyz, const uint64_t* __restrict__ keys,
                                 const uint32_t* __restrict__ vals, int64_t n, void* grid, uint64_t cap) {
  GridSlot* slots = (GridSlot*)((char*)grid + sizeof(GridHeader));
  double* sx = (double*)(slots + cap);
  int* orig = (int*)(sx + 3 * n);
  float* rel = (float*)(orig + n);
  const double cell = ((const GridHeader*)grid)->cell;
  const uint64_t mask = cap - 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t o = vals[i];
    sx[3 * i] = xyz[3 * (int64_t)o];
    sx[3 * i + 1] = xyz[3 * (int64_t)o + 1];
    sx[3 * i + 2] = xyz[3 * (int64_t)o + 2];
    orig[i] = (int)o;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = xyz[3 * (int64_t)o + a];
      rel[3 * i + a] = (float)(v - (double)cell_of(v, cell) * cell);
    }
    const uint64_t key = keys[i];
    if (i == 0 || keys[i - 1] != key) {                  // run start: claim the slot (keys are unique per run)
      int64_t e = i + 1;
      while (e < n && keys[e] == key) ++e;               // runs are a handful of points
      uint64_t slot = cell_slot(key) & mask;
      while (atomicCAS(&slots[slot].key, LB_EMPTY_KEY, (unsigned long long)key) != LB_EMPTY_KEY) slot = (slot + 1) & mask;
      slots[slot].start = (int)i;
      slots[slot].count = (int)(e - i);
    }
  }
}

constexpr int MAX_NBR = 32;
struct FrameRef {
  const void* grid;
  const double* xyz;
  const float* prob;
  int64_t n;
};
struct ScoreParams {
  FrameRef nbr[MAX_NBR];
  int n_nbr;
};

// Phase 1 -- nearest-neighbour search, one THREAD per (cell-sorted query point, neighbour frame): blockIdx.y = frame.
// Consecutive threads are spatial neighbours, so their cell probes and candidate reads coalesce in L1/L2; every
// (point, frame) pair is independent, so a frame's search exposes 24x the parallelism of a per-point loop.
// Candidates are screened in float32 on cell-relative coordinates (error ~1e-7 m, screening radius inflated by 1e-5 m so a
// true match can never be dropped) and only survivors are evaluated in the reference's arithmetic: float64, no FMA
// contraction (sklearn's rdist), ties to the smaller original row.  Result: the exact nearest neighbour if it lies within
// the threshold, else -1 -- identical to the KD-tree query + `dists <= dis_thresh` of LiDAL.py:66-69.
// nn_fm[f * nq + sp] (frame-major, by SORTED position): coalesced writes; phase 2 reads it back by sorted position.
__global__ void __launch_bounds__(128)
nn_search_kernel(const void* __restrict__ q_grid, int64_t nq, const __grid_constant__ ScoreParams P, double thresh,
                 int* __restrict__ nn_fm) {
  const GridView qg = grid_view(q_grid);
  const int f = blockIdx.y;
  const GridView g = grid_view(P.nbr[f].grid);
  const float tf = (float)thresh + 1e-5f;                    // inflated screening radius: must never drop a true match
  const float t2f = tf * tf;
  for (int64_t sp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; sp < nq; sp += (int64_t)gridDim.x * blockDim.x) {
    const double qx = __ldg(&qg.xyz[3 * sp]), qy = __ldg(&qg.xyz[3 * sp + 1]), qz = __ldg(&qg.xyz[3 * sp + 2]);
    const long long cx = cell_of(qx, g.cell), cy = cell_of(qy, g.cell), cz = cell_of(qz, g.cell);
    // query relative to its own cell origin; a neighbouring cell's origin differs by exactly (ox, oy, oz) * cell
    const double rx = qx - (double)cx * g.cell, ry = qy - (double)cy * g.cell, rz = qz - (double)cz * g.cell;
    // Distance from the query to the lower / upper face of its own cell per axis: a neighbouring cell can hold a match only
    // if its box lies within the match radius.  The pruning runs in float32 against the inflated screening radius (the
    // gaps are < cell, so their float32 error is ~1e-8 m -- far inside the 1e-5 m inflation): conservative, so no cell that
    // could hold a match is skipped; typically ~6 of 27 cells remain.
    const float cellf = (float)g.cell;
    const float rxf = (float)rx, ryf = (float)ry, rzf = (float)rz;
    const float gap2x[3] = {rxf * rxf, 0.f, (cellf - rxf) * (cellf - rxf)};
    const float gap2y[3] = {ryf * ryf, 0.f, (cellf - ryf) * (cellf - ryf)};
    const float gap2z[3] = {rzf * rzf, 0.f, (cellf - rzf) * (cellf - rzf)};
    double best = INFINITY;
    int bi = 0x7fffffff;
    // which of the 27 cells can hold a match: straight-line predicated arithmetic (no divergence), one bit per cell
    uint32_t cells = 0;
#pragma unroll
    for (int ix = 0; ix < 3; ++ix)
#pragma unroll
      for (int iy = 0; iy < 3; ++iy)
#pragma unroll
        for (int iz = 0; iz < 3; ++iz)
          cells |= (gap2x[ix] + gap2y[iy] + gap2z[iz] <= t2f) ? (1u << (ix * 9 + iy * 3 + iz)) : 0u;
    // every lane walks only its own surviving cells (~6): the warp iterates max-popcount times instead of 27
    while (cells) {
      const int c = __ffs(cells) - 1;
      cells &= cells - 1;
      const int ox = c / 9 - 1, oy = (c / 3) % 3 - 1, oz = c % 3 - 1;
      const int2 run = grid_find(g, cell_key(cx + ox, cy + oy, cz + oz));
      if (run.y == 0) continue;
      const float fx = rxf - (float)ox * cellf, fy = ryf - (float)oy * cellf, fz = rzf - (float)oz * cellf;
      for (int j = run.x; j < run.x + run.y; ++j) {
        const float ex = fx - __ldg(&g.rel[3 * (int64_t)j]), ey = fy - __ldg(&g.rel[3 * (int64_t)j + 1]),
                    ez = fz - __ldg(&g.rel[3 * (int64_t)j + 2]);
        if (ex * ex + ey * ey + ez * ez > t2f) continue;
        // sklearn euclidean rdist: d = 0; d += t*t per axis, no FMA contraction
        const double tx = __dsub_rn(qx, __ldg(&g.xyz[3 * (int64_t)j]));
        const double ty = __dsub_rn(qy, __ldg(&g.xyz[3 * (int64_t)j + 1]));
        const double tz = __dsub_rn(qz, __ldg(&g.xyz[3 * (int64_t)j + 2]));
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz));
        const int oj = __ldg(&g.orig[j]);
        if (d2 < best || (d2 == best && oj < bi)) { best = d2; bi = oj; }
      }
    }
    nn_fm[(int64_t)f * nq + sp] = (bi != 0x7fffffff && __dsqrt_rn(best) <= thresh) ? bi : -1;
  }
}

// Phase 2 -- one WARP per query point (walked by sorted position; results land at the point's original row), lane = class:
// accumulate over the matched neighbours in nei_ids order.
__global__ void __launch_bounds__(256)
interframe_kernel(const void* __restrict__ q_grid, const float* __restrict__ q_prob, int64_t nq, int n_cls,
                  const __grid_constant__ ScoreParams P, const int* __restrict__ nn_fm, double* __restrict__ interd_out,
                  float* __restrict__ intere_out, int* __restrict__ count_out, int* __restrict__ nn_out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float eps = 0.00001f;
  const int* __restrict__ q_orig = grid_view(q_grid).orig;
  for (int64_t sp = warp; sp < nq; sp += nwarps) {
    const int64_t p = __ldg(&q_orig[sp]);
    const float q = lane < n_cls ? __ldg(&q_prob[p * n_cls + lane]) : 0.f;
    const int my_nn = lane < P.n_nbr ? __ldg(&nn_fm[(int64_t)lane * nq + sp]) : -1;
    if (nn_out && lane < P.n_nbr) nn_out[p * P.n_nbr + lane] = my_nn;
    float sum = q;
    double interd = 0.0;
    int cnt = 1;
    unsigned hits = __ballot_sync(full, my_nn >= 0);
    while (hits) {                                        // matched neighbour frames only, in nei_ids order
      const int f = __ffs(hits) - 1;
      hits &= hits - 1;
      const int bi = __shfl_sync(full, my_nn, f);
      const float pn = lane < n_cls ? __ldg(&P.nbr[f].prob[(int64_t)bi * n_cls + lane]) : 0.f;
      sum = __fadd_rn(sum, pn);
      float kl = 0.f;
      if (lane < n_cls) {
        const double x = (double)__fadd_rn(q, eps), y = (double)__fadd_rn(pn, eps);
        kl = (float)__dadd_rn(__dsub_rn(__dmul_rn(x, log(__ddiv_rn(x, y))), x), y);
      }
      const float ks = np_pairwise_sum32(kl, n_cls, lane);
      interd += (double)ks;      // only lane 0's value is meaningful
      cnt += 1;
    }
    // LiDAL.py:75-76  sum_prob /= map_count (f32 / f64 -> f32);  entropy(pk) renormalises, natural log
    const float pm = lane < n_cls ? (float)__ddiv_rn((double)sum, (double)cnt) : 0.f;
    float s = np_pairwise_sum32(pm, n_cls, lane);
    s = __shfl_sync(full, s, 0);
    const float pk = __fdiv_rn(pm, s);
    float en = 0.f;
    if (lane < n_cls) {
      const double pd = (double)pk;
      en = pd > 0.0 ? (float)__dmul_rn(-pd, log(pd)) : (pd == 0.0 ? 0.f : -INFINITY);
    }
    const float H = np_pairwise_sum32(en, n_cls, lane);
    if (lane == 0) {
      // LiDAL.py:79-81
      const int m = cnt - 1;
      interd_out[p] = m > 0 ? __ddiv_rn(interd, (double)m) : interd;
      intere_out[p] = H;
      if (count_out) count_out[p] = m;
    }
  }
}

// One block per region.  sv_interds = interd_points[p_ids].mean() (float64) and sv_interes = intere_points[p_ids].mean()
// (float32) follow numpy's pairwise order exactly (np_sum.cuh) -- the selection compares these values, so the last bit
// matters; the centre is a fixed-order float64 tree (numpy adds the rows of query_points[p_ids] sequentially in float64:
// the two differ far below float32 resolution of the stored centre).  Deterministic.
__global__ void __launch_bounds__(NP_BLOCK_THREADS)
region_reduce_kernel(const double* __restrict__ interd, const float* __restrict__ intere, const double* __restrict__ xyz,
                     const int* __restrict__ ptr, const int* __restrict__ pts, float* __restrict__ sv_d,
                     float* __restrict__ sv_e, int64_t* __restrict__ sv_n, float* __restrict__ sv_c) {
  __shared__ NpLeafLayout L;
  __shared__ double sums[NP_MAX_LEAVES];
  __shared__ double sh[3][NP_BLOCK_THREADS];
  const int r = blockIdx.x;
  const int b = ptr[r], e = ptr[r + 1];
  np_build_leaves(e - b, L);
  const double sd = np_sum_leaves<double>(interd, pts + b, L, sums);
  const float se = np_sum_leaves<float>(intere, pts + b, L, reinterpret_cast<float*>(sums));
  if (sv_c) {
    double a[3] = {0, 0, 0};
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int p = __ldg(&pts[i]);
      a[0] += xyz[3 * (int64_t)p];
      a[1] += xyz[3 * (int64_t)p + 1];
      a[2] += xyz[3 * (int64_t)p + 2];
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) sh[v][threadIdx.x] = a[v];
    __syncthreads();
    for (int s = NP_BLOCK_THREADS / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s)
#pragma unroll
        for (int v = 0; v < 3; ++v) sh[v][threadIdx.x] += sh[v][threadIdx.x + s];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const double n = (double)(e - b);
    sv_d[r] = (float)__ddiv_rn(sd, n);                  // float64 mean stored into a float32 array (LiDAL.py:88,97)
    sv_e[r] = __fdiv_rn(se, (float)(e - b));            // float32 mean (LiDAL.py:89,98)
    if (sv_n) sv_n[r] = e - b;
    if (sv_c) {
      sv_c[3 * r] = (float)(sh[0][0] / n);
      sv_c[3 * r + 1] = (float)(sh[1][0] / n);
      sv_c[3 * r + 2] = (float)(sh[2][0] / n);
    }
  }
}

// float -> order-preserving uint32
__global__ void argsort_prepare(const float* __restrict__ keys, int64_t n, uint64_t* __restrict__ k64,
                                uint32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t u = __float_as_uint(keys[i]);
    if (u == 0x80000000u) u = 0;                   // -0.0 == +0.0 for numpy's comparison sort
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    k64[i] = u;
    vals[i] = (uint32_t)i;
  }
}

__global__ void f32_to_f64_xyz(const float* __restrict__ c, int64_t n3, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)c[i];
}

// LiDAL.py:252  dist = np.sqrt(np.square(a - b).sum()) in float32; in range iff dist < radius.
__global__ void region_pairs_kernel(const float* __restrict__ c, int64_t n, float radius, const void* grid,
                                    int* __restrict__ row_ptr, int* __restrict__ nbr_idx, int fill) {
  const GridView g = grid_view(grid);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float ax = c[3 * i], ay = c[3 * i + 1], az = c[3 * i + 2];
    const long long cx = cell_of((double)ax, g.cell), cy = cell_of((double)ay, g.cell), cz = cell_of((double)az, g.cell);
    int cnt = 0;
    const int base = fill ? row_ptr[i] : 0;
    for (int t = 0; t < 27; ++t) {
      const int2 run = grid_find(g, cell_key(cx + t / 9 - 1, cy + (t / 3) % 3 - 1, cz + t % 3 - 1));
      for (int jj = run.x; jj < run.x + run.y; ++jj) {
        const int j = __ldg(&g.orig[jj]);
        if (j == i) continue;
        float ex = __fsub_rn(ax, c[3 * (int64_t)j]), ey = __fsub_rn(ay, c[3 * (int64_t)j + 1]),
              ez = __fsub_rn(az, c[3 * (int64_t)j + 2]);
        float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
        if (d < radius) {
          if (fill) nbr_idx[base + cnt] = j;
          ++cnt;
        }
      }
    }
    if (!fill) row_ptr[i] = cnt;
  }
}

}  // namespace lb
using namespace lb;

static inline int grid1d(int64_t n, int block) {
  int64_t b = (n + block - 1) / block, cap = (int64_t)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" size_t lb_frame_grid_bytes(int64_t n) { return grid_bytes_for(n); }
extern "C" size_t lb_frame_grid_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return (((size_t)n * 8 + 255) & ~(size_t)255) + (((size_t)n * 4 + 255) & ~(size_t)255) + lb_sort_pairs_ws_bytes(n) + 256;
}
extern "C" int lb_frame_grid_build(const double* xyz, int64_t n, double cell, void* grid, size_t bytes, void* ws,
                                   size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && grid && cell > 0 && ws, "bad arguments");
  LB_CHECK_ARG(((uintptr_t)grid & 15) == 0, "grid must be 16-byte aligned");
  if (bytes < grid_bytes_for(n)) { set_error("lb_frame_grid_build: grid buffer too small"); return LB_ECAP; }
  if (ws_bytes < lb_frame_grid_ws_bytes(n)) { set_error("lb_frame_grid_build: workspace too small"); return LB_ECAP; }
  cudaStream_t st = as_stream(stream);
  uint64_t cap = table_capacity(n);
  grid_init_kernel<<<grid1d((int64_t)cap, 256), 256, 0, st>>>(grid, cell, n, cap); LB_LAUNCHED(1);
  if (n > 0) {
    LB_CHECK_ARG(xyz, "null xyz");
    uint64_t* keys = (uint64_t*)ws;
    uint32_t* vals = (uint32_t*)((char*)ws + (((size_t)n * 8 + 255) & ~(size_t)255));
    void* sort_ws = (char*)vals + (((size_t)n * 4 + 255) & ~(size_t)255);
    grid_keys_kernel<<<grid1d(n, 256), 256, 0, st>>>(xyz, n, cell, keys, vals); LB_LAUNCHED(1);
    int rc = lb_sort_pairs(keys, vals, n, 63, sort_ws, lb_sort_pairs_ws_bytes(n), stream);
    if (rc != LB_OK) return rc;
    grid_fill_kernel<<<grid1d(n, 256), 256, 0, st>>>(xyz, keys, vals, n, grid, cap); LB_LAUNCHED(1);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_interframe_score_ws_bytes(int64_t nq, int n_nbr) {
  return (size_t)(nq > 0 ? nq : 1) * (size_t)(n_nbr > 0 ? n_nbr : 1) * 4 + 256;
}
extern "C" int lb_interframe_score(const void* q_grid, const float* q_prob, int64_t nq, int n_cls,
                                   const lb_frame_ref* nbrs, int n_nbr, double dis_thresh, double cell,
                                   double* interd, float* intere, int32_t* count, int32_t* nn, void* ws, size_t ws_bytes,
                                   void* stream) {
  LB_CHECK_ARG(nq >= 0 && n_cls > 0 && n_cls <= 32, "n_cls must be in [1,32]");
  LB_CHECK_ARG(n_nbr >= 0 && n_nbr <= MAX_NBR, "at most 32 neighbour frames");
  LB_CHECK_ARG(cell >= dis_thresh * 1.001, "grid cell must exceed dis_thresh (27-cell probe exactness)");
  if (nq == 0) return LB_OK;
  LB_CHECK_ARG(q_grid && q_prob && interd && intere && ws && (nbrs || n_nbr == 0), "null pointer");
  if (ws_bytes < lb_interframe_score_ws_bytes(nq, n_nbr)) { set_error("lb_interframe_score: workspace too small"); return LB_ECAP; }
  ScoreParams P;
  P.n_nbr = n_nbr;
  for (int i = 0; i < n_nbr; ++i) {
    LB_CHECK_ARG(nbrs[i].grid && nbrs[i].prob, "null neighbour frame");
    P.nbr[i].grid = nbrs[i].grid; P.nbr[i].xyz = nbrs[i].xyz; P.nbr[i].prob = nbrs[i].prob; P.nbr[i].n = nbrs[i].n;
  }
  cudaStream_t st = as_stream(stream);
  int* nn_fm = (int*)ws;
  if (n_nbr > 0) {
    int64_t b1 = (nq + 127) / 128, cap1 = (int64_t)sm_count() * 32;
    nn_search_kernel<<<dim3((unsigned)(b1 > cap1 ? cap1 : b1), (unsigned)n_nbr), 128, 0, st>>>(q_grid, nq, P, dis_thresh, nn_fm); LB_LAUNCHED(1);
  }
  int64_t blocks = (nq + 7) / 8, cap = (int64_t)sm_count() * 8;
  interframe_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(q_grid, q_prob, nq, n_cls, P, nn_fm, interd, intere, count, nn); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_region_reduce(const double* interd, const float* intere, const double* xyz, const int32_t* region_ptr,
                                const int32_t* region_pts, int n_regions, float* sv_d, float* sv_e, int64_t* sv_n,
                                float* sv_c, void* stream) {
  LB_CHECK_ARG(n_regions >= 0, "n_regions < 0");
  if (n_regions == 0) return LB_OK;
  LB_CHECK_ARG(interd && intere && xyz && region_ptr && region_pts && sv_d && sv_e, "null pointer");
  region_reduce_kernel<<<n_regions, NP_BLOCK_THREADS, 0, as_stream(stream)>>>(interd, intere, xyz, region_ptr, region_pts, sv_d, sv_e,
                                                                 sv_n, sv_c); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_argsort_ws_bytes(int64_t n) { return lb_sort_pairs_ws_bytes(n) + (size_t)(n > 0 ? n : 1) * 12 + 512; }
extern "C" int lb_argsort_f32(const float* keys, int64_t n, int32_t* order, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && ws, "bad arguments");
  if (ws_bytes < lb_argsort_ws_bytes(n)) { set_error("lb_argsort_f32: workspace too small"); return LB_ECAP; }
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(keys && order, "null pointer");
  cudaStream_t st = as_stream(stream);
  uint64_t* k64 = (uint64_t*)ws;
  size_t off = ((size_t)n * 8 + 255) & ~(size_t)255;
  void* sws = (char*)ws + off;
  argsort_prepare<<<grid1d(n, 256), 256, 0, st>>>(keys, n, k64, (uint32_t*)order); LB_LAUNCHED(1);
  int rc = lb_sort_pairs(k64, (uint32_t*)order, n, 32, sws, ws_bytes - off, stream);
  if (rc != LB_OK) return rc;
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" size_t lb_region_pairs_ws_bytes(int64_t n) {
  return (((size_t)(n > 0 ? n : 1) * 24 + 255) & ~(size_t)255) + ((grid_bytes_for(n) + 255) & ~(size_t)255) +
         lb_frame_grid_ws_bytes(n) + 256;
}
extern "C" int lb_region_pairs(const float* centers, int64_t n, float radius, int32_t* row_counts_or_ptr,
                               int32_t* nbr_idx, void* ws, size_t ws_bytes, void* stream) {
  LB_CHECK_ARG(n >= 0 && radius > 0 && ws, "bad arguments");
  if (ws_bytes < lb_region_pairs_ws_bytes(n)) { set_error("lb_region_pairs: workspace too small"); return LB_ECAP; }
  if (n == 0) return LB_OK;
  LB_CHECK_ARG(centers && row_counts_or_ptr, "null pointer");
  cudaStream_t st = as_stream(stream);
  double* xyz = (double*)ws;
  void* grid = (char*)ws + (((size_t)n * 24 + 255) & ~(size_t)255);
  f32_to_f64_xyz<<<grid1d(n * 3, 256), 256, 0, st>>>(centers, n * 3, xyz); LB_LAUNCHED(1);
  void* gws = (char*)grid + ((grid_bytes_for(n) + 255) & ~(size_t)255);
  int rc = lb_frame_grid_build(xyz, n, (double)radius * 1.01, grid, grid_bytes_for(n), gws, lb_frame_grid_ws_bytes(n), stream);
  if (rc != LB_OK) return rc;
  region_pairs_kernel<<<grid1d(n, 128), 128, 0, st>>>(centers, n, radius, grid, row_counts_or_ptr, nbr_idx,
                                                      nbr_idx ? 1 : 0); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---------------------------------------------------------------------------------------- frame-level baselines (F3)
// score/frame_level/{softmax_entropy.py:34, margin_sampling.py:33-34, least_confidence_sampling.py}: three per-frame means
// over the points of one prob map: entropy(prob) (scipy: renormalise, entr in double -> float, float pairwise sum),
// top1 - top2, top1.  One warp per point, lane = class; fixed-order block partials -> deterministic.
namespace lb {
__global__ void __launch_bounds__(256)
frame_level_kernel(const float* __restrict__ prob, int64_t n, int n_cls, double* __restrict__ partial /*[grid,3]*/) {
  __shared__ double sh[3][8];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double acc_e = 0, acc_m = 0, acc_c = 0;
  for (int64_t p = warp; p < n; p += nwarps) {
    const float x = lane < n_cls ? __ldg(&prob[p * n_cls + lane]) : 0.f;
    float s = np_pairwise_sum32(x, n_cls, lane);
    s = __shfl_sync(full, s, 0);
    const float pk = __fdiv_rn(x, s);
    float en = 0.f;
    if (lane < n_cls) {
      const double pd = (double)pk;
      en = pd > 0.0 ? (float)__dmul_rn(-pd, log(pd)) : (pd == 0.0 ? 0.f : -INFINITY);
    }
    const float H = np_pairwise_sum32(en, n_cls, lane);
    float v1 = lane < n_cls ? x : -INFINITY;              // top-1
#pragma unroll
    for (int d = 16; d; d >>= 1) v1 = fmaxf(v1, __shfl_xor_sync(full, v1, d));
    const unsigned at_max = __ballot_sync(full, lane < n_cls && x == v1);
    const int first = __ffs(at_max) - 1;
    float v2 = (lane < n_cls && lane != first) ? x : -INFINITY;   // top-2 (duplicates of the max count, as np.sort gives)
#pragma unroll
    for (int d = 16; d; d >>= 1) v2 = fmaxf(v2, __shfl_xor_sync(full, v2, d));
    if (lane == 0) { acc_e += (double)H; acc_m += (double)__fsub_rn(v1, v2); acc_c += (double)v1; }
  }
  if (lane == 0) { sh[0][w] = acc_e; sh[1][w] = acc_m; sh[2][w] = acc_c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sh[threadIdx.x][i];
    partial[(int64_t)blockIdx.x * 3 + threadIdx.x] = t;
  }
}
__global__ void frame_level_final(const double* __restrict__ partial, int blocks, int64_t n, double* __restrict__ out) {
  if (threadIdx.x < 3) {
    double t = 0;
    for (int b = 0; b < blocks; ++b) t += partial[(int64_t)b * 3 + threadIdx.x];
    out[threadIdx.x] = n > 0 ? t / (double)n : 0.0;
  }
}
}  // namespace lb
extern "C" size_t lb_frame_level_ws_bytes(void) { return (size_t)sm_count() * 8 * 3 * 8 + 64; }
extern "C" int lb_frame_level_scores(const float* prob, int64_t n, int n_cls, double* out3, void* ws, size_t ws_bytes,
                                     void* stream) {
  LB_CHECK_ARG(n >= 0 && n_cls > 1 && n_cls <= 32 && out3 && ws, "bad arguments");
  if (ws_bytes < lb_frame_level_ws_bytes()) { set_error("lb_frame_level_scores: workspace too small"); return LB_ECAP; }
  LB_CHECK_ARG(prob || n == 0, "null prob");
  const int blocks = sm_count() * 8;
  frame_level_kernel<<<blocks, 256, 0, as_stream(stream)>>>(prob, n, n_cls, (double*)ws); LB_LAUNCHED(1);
  frame_level_final<<<1, 32, 0, as_stream(stream)>>>((const double*)ws, blocks, n, out3); LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
