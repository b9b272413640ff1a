// Micro-benchmark: how fast can one SM fill a ring of UMMA-layout A tiles with gathered rows?
//   mode 0: TMA tile::gather4 -- a stage (128 rows) is ONE warp instruction, 32 lanes x 4 rows; warp w owns stages s % W == w
//   mode 1: cp.async lock-step (what conv_tc_kernel does in round 1): all W warps fill every stage, 16 B per thread-copy,
//           completion through cp.async.mbarrier.arrive.noinc
//   mode 2: cp.async, warp w owns stages s % W == w entirely (decoupled stages)
// A consumer warp waits on each full barrier and releases the slot at once (stands in for the MMA warp), so the number
// printed is the producer-side ceiling for the sparse-conv gather.  Not part of the library build.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/gather_bw tools/experiments/gather_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
static __device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory"); }
static __device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
static __device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
static __device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t n) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory"); }
static __device__ __forceinline__ void cp_arrive_noinc(uint64_t* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
static __device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

constexpr int STAGES = 12;

template <int MODE, int ROW_BYTES>
__global__ void __launch_bounds__(1024, 1)
bench_kernel(const __grid_constant__ CUtensorMap map, const char* __restrict__ in, const int* __restrict__ idx, int tiles_per_cta, int W, int WPS) {
  extern __shared__ uint8_t raw[];
  uint8_t* ring = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  constexpr int STAGE_BYTES = 128 * ROW_BYTES;
  constexpr int CHUNKS = ROW_BYTES / 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], MODE == 0 ? 1 : (MODE == 1 ? W * 32 : 32 * WPS));   // modes 2-4: the owning warp's 32 lanes
      mbar_init(&empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int* my_idx = idx + (size_t)blockIdx.x * tiles_per_cta * 128;
  if (warp == W) {                                           // consumer
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int s = t % STAGES;
      mbar_wait(&full[s], (t / STAGES) & 1);
      if (lane == 0) mbar_arrive(&empty[s]);
      __syncwarp();
    }
  } else if (warp < W) {
    if (MODE == 0) {
      for (int t = warp; t < tiles_per_cta; t += W) {
        const int s = t % STAGES;
        const int4 r = __ldg((const int4*)(my_idx + (size_t)t * 128) + lane);
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        if (lane == 0) mbar_expect(&full[s], STAGE_BYTES);
        __syncwarp();
        gather4(smem_u32(ring + s * STAGE_BYTES + lane * 4 * ROW_BYTES), &map, &full[s], 0, r.x, r.y, r.z, r.w);
      }
    } else if (MODE == 1) {
      const int tt = warp * 32 + lane, nthr = W * 32;
      const int chunk = tt % CHUNKS, row0 = tt / CHUNKS, rpp = nthr / CHUNKS;
      for (int t = 0; t < tiles_per_cta; ++t) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        for (int r = row0; r < 128; r += rpp) {
          const int row = __ldg(my_idx + (size_t)t * 128 + r);
          const uint32_t sw = ROW_BYTES == 128 ? (uint32_t)(chunk ^ (r & 7)) : (uint32_t)(chunk ^ ((r >> 1) & 3));
          cp_async16(smem_u32(ring + s * STAGE_BYTES + r * ROW_BYTES + sw * 16), in + (size_t)(row >= 0 ? row : 0) * ROW_BYTES + chunk * 16, row >= 0 ? 16u : 0u);
        }
        cp_arrive_noinc(&full[s]);
      }
    } else if (MODE == 3) {
      // register-staged: LDG.128 x8 in flight, then STS.128 (generic-proxy writes; consumer would need a proxy fence)
      const int sub = warp % WPS, grp = warp / WPS, ngrp = W / WPS;
      const int chunk = lane % CHUNKS, row0 = sub * (32 / CHUNKS) + lane / CHUNKS, rpp = WPS * 32 / CHUNKS;
      for (int t = grp; t < tiles_per_cta; t += ngrp) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        for (int rb = row0; rb < 128; rb += rpp * 8) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = rb + u * rpp;
            const int row = __ldg(my_idx + (size_t)t * 128 + r);
            v[u] = row >= 0 ? __ldg((const uint4*)(in + (size_t)row * ROW_BYTES + chunk * 16)) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = rb + u * rpp;
            const uint32_t sw = ROW_BYTES == 128 ? (uint32_t)(chunk ^ (r & 7)) : (uint32_t)(chunk ^ ((r >> 1) & 3));
            *(uint4*)(ring + s * STAGE_BYTES + r * ROW_BYTES + sw * 16) = v[u];
          }
        }
        mbar_arrive(&full[s]);
      }
    } else if (MODE == 4) {
      // cp.async only for present rows; absent rows get an STS.128 of zeros
      const int sub = warp % WPS, grp = warp / WPS, ngrp = W / WPS;
      const int chunk = lane % CHUNKS, row0 = sub * (32 / CHUNKS) + lane / CHUNKS, rpp = WPS * 32 / CHUNKS;
      for (int t = grp; t < tiles_per_cta; t += ngrp) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
#pragma unroll 8
        for (int r = row0; r < 128; r += rpp) {
          const int row = __ldg(my_idx + (size_t)t * 128 + r);
          const uint32_t sw = ROW_BYTES == 128 ? (uint32_t)(chunk ^ (r & 7)) : (uint32_t)(chunk ^ ((r >> 1) & 3));
          uint8_t* dst = ring + s * STAGE_BYTES + r * ROW_BYTES + sw * 16;
          if (row >= 0) cp_async16(smem_u32(dst), in + (size_t)row * ROW_BYTES + chunk * 16, 16u);
          else *(uint4*)dst = make_uint4(0, 0, 0, 0);
        }
        cp_arrive_noinc(&full[s]);
      }
    } else {
      const int sub = warp % WPS, grp = warp / WPS, ngrp = W / WPS;
      const int chunk = lane % CHUNKS, row0 = sub * (32 / CHUNKS) + lane / CHUNKS, rpp = WPS * 32 / CHUNKS;
      for (int t = grp; t < tiles_per_cta; t += ngrp) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
#pragma unroll 8
        for (int r = row0; r < 128; r += rpp) {
          const int row = __ldg(my_idx + (size_t)t * 128 + r);
          const uint32_t sw = ROW_BYTES == 128 ? (uint32_t)(chunk ^ (r & 7)) : (uint32_t)(chunk ^ ((r >> 1) & 3));
          cp_async16(smem_u32(ring + s * STAGE_BYTES + r * ROW_BYTES + sw * 16), in + (size_t)(row >= 0 ? row : 0) * ROW_BYTES + chunk * 16, row >= 0 ? 16u : 0u);
        }
        cp_arrive_noinc(&full[s]);
      }
    }
  }
  __syncthreads();
  asm volatile("cp.async.wait_all;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int ROW_BYTES>
static void run(EncodeFn enc, char* d_in, int n_rows, const int* d_idx, int tiles_per_cta, int W, double miss, const char* label, int WPS = 1) {
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)(ROW_BYTES / 2), (cuuint64_t)n_rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ROW_BYTES};
  cuuint32_t box[2] = {(cuuint32_t)(ROW_BYTES / 2), 1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_in, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
  const int smem = STAGES * 128 * ROW_BYTES + 1024;
  cudaFuncSetAttribute(bench_kernel<MODE, ROW_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    bench_kernel<MODE, ROW_BYTES><<<148, 32 * (W + 1), smem>>>(map, d_in, d_idx, tiles_per_cta, W, WPS);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("%s: %s\n", label, cudaGetErrorString(e)); exit(2); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  const double bytes = 148.0 * tiles_per_cta * 128 * ROW_BYTES * (1.0 - miss);
  const double stage_cyc = best * 1e-3 * 1.9e9 / tiles_per_cta;
  printf("%-24s rowB=%3d W=%2d rows=%8d miss=%.2f : %7.3f ms  %7.1f GB/s gathered  %6.0f cyc/stage(128 rows)\n", label, ROW_BYTES, W, n_rows, miss, best,
         bytes / best * 1e-6, stage_cyc);
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode entry point\n"); return 1; }
  EncodeFn enc = (EncodeFn)fn;
  const int tiles_per_cta = 2000;
  const size_t n_idx = (size_t)148 * tiles_per_cta * 128;
  for (int n_rows : {1000000}) {
    char* d_in; cudaMalloc(&d_in, (size_t)n_rows * 128); cudaMemset(d_in, 1, (size_t)n_rows * 128);
    for (double miss : {0.0, 0.75}) {
      std::vector<int> h(n_idx);
      srand(7);
      // rows walk roughly in order with local scatter (like a kernel map of a mask-sorted tile), some missing
      for (size_t i = 0; i < n_idx; ++i) {
        const bool m = (rand() % 1000) < (int)(miss * 1000);
        long long base = (long long)((double)i / n_idx * n_rows);
        long long r = base + (rand() % 4096) - 2048;
        if (r < 0) r = 0; if (r >= n_rows) r = n_rows - 1;
        h[i] = m ? -1 : (int)r;
      }
      int* d_idx; cudaMalloc(&d_idx, n_idx * 4); cudaMemcpy(d_idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
      for (int W : {8, 12}) run<0, 128>(enc, d_in, n_rows, d_idx, tiles_per_cta, W, miss, "tma gather4");
      for (int W : {8, 12}) run<0, 64>(enc, d_in, n_rows * 2, d_idx, tiles_per_cta, W, miss, "tma gather4");
      for (int W : {8, 12, 16, 24}) run<2, 128>(enc, d_in, n_rows, d_idx, tiles_per_cta, W, miss, W > 12 ? "cp.async 2 warps/stage" : "cp.async 1 warp/stage", W > 12 ? 2 : 1);
      for (int W : {8, 12, 16, 24}) run<2, 64>(enc, d_in, n_rows * 2, d_idx, tiles_per_cta, W, miss, W > 12 ? "cp.async 2 warps/stage" : "cp.async 1 warp/stage", W > 12 ? 2 : 1);
      for (int W : {8, 16, 24}) run<3, 128>(enc, d_in, n_rows, d_idx, tiles_per_cta, W, miss, W > 12 ? "ldg+sts 2 warps/stage" : "ldg+sts 1 warp/stage", W > 12 ? 2 : 1);
      for (int W : {8, 16, 24}) run<3, 64>(enc, d_in, n_rows * 2, d_idx, tiles_per_cta, W, miss, W > 12 ? "ldg+sts 2 warps/stage" : "ldg+sts 1 warp/stage", W > 12 ? 2 : 1);
      for (int W : {8, 16, 24}) run<4, 128>(enc, d_in, n_rows, d_idx, tiles_per_cta, W, miss, W > 12 ? "cp.async+sts0 2w/stage" : "cp.async+sts0 1w/stage", W > 12 ? 2 : 1);
      for (int W : {8, 16, 24}) run<4, 64>(enc, d_in, n_rows * 2, d_idx, tiles_per_cta, W, miss, W > 12 ? "cp.async+sts0 2w/stage" : "cp.async+sts0 1w/stage", W > 12 ? 2 : 1);
      cudaFree(d_idx);
    }
    cudaFree(d_in);
  }
  return 0;
}
