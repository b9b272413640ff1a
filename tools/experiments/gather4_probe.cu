// Stand-alone probe for next round's A-operand path: does `cp.async.bulk.tensor.2d...tile::gather4` (sm_100a) do what the
// sparse-conv gather needs?  For each tensor-map variant it gathers 32 x 4 rows of a [n_rows, 64] bf16 matrix into a
// 128-row SWIZZLE_128B smem tile (the UMMA K-major layout conv_tc_kernel builds with cp.async today), with some row
// indices = -1 and some >= n_rows, then un-swizzles the tile and compares with the expected gather (zeros for invalid rows).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/gather4_probe tools/experiments/gather4_probe.cu -lcuda
//   /tmp/gather4_probe
// Not part of the library build.  Instruction text from cute/arch/copy_sm100_tma.hpp (SM100_TMA_LOAD_2D_GATHER4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ rows /*[128]*/,
                             uint16_t* __restrict__ out /*[128][64] un-swizzled*/, int expect_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* tile = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);     // 128 rows x 128 B, SWIZZLE_128B atoms of 8 rows
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(expect_bytes) : "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {                                                 // lane l gathers rows 4l .. 4l+3 -> 512 B at tile + 512 l
    const int l = threadIdx.x;
    const int r0 = rows[4 * l], r1 = rows[4 * l + 1], r2 = rows[4 * l + 2], r3 = rows[4 * l + 3];
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(tile + 512 * l)), "l"(&map), "r"(smem_u32(&bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
  }
  // wait (bounded: a wrong byte count must not hang the box)
  uint32_t done = 0;
  for (int it = 0; it < 2000000 && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  __syncthreads();
  if (threadIdx.x == 0 && !done) printf("  (barrier never completed: transaction byte count differs from %d)\n", expect_bytes);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {               // un-swizzle: chunk c of row r sits at c ^ (r & 7)
    const int r = i / 8, c = i % 8;
    const uint4 v = *(const uint4*)(tile + r * 128 + ((c ^ (r & 7)) * 16));
    *(uint4*)(out + r * 64 + c * 8) = v;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int n_rows = 1000, C = 64;
  std::vector<uint16_t> h((size_t)n_rows * C);
  for (int r = 0; r < n_rows; ++r) for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (uint16_t)((r * 64 + c) & 0xffff);
  std::vector<int> rows(128);
  srand(1);
  for (int i = 0; i < 128; ++i) rows[i] = rand() % n_rows;
  rows[5] = -1; rows[6] = -1; rows[40] = n_rows; rows[41] = n_rows + 7; rows[127] = -1;     // missing neighbours / out of range
  uint16_t *d_in, *d_out; int* d_rows;
  cudaMalloc(&d_in, h.size() * 2); cudaMalloc(&d_out, 128 * 64 * 2); cudaMalloc(&d_rows, 128 * 4);
  cudaMemcpy(d_in, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_rows, rows.data(), 128 * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode entry point\n"); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 1024);
  for (int box_rows : {1, 4}) {
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)n_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_in, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box rows = %d: encode -> %d\n", box_rows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(d_out, 0xEE, 128 * 64 * 2);
    probe_kernel<<<1, 128, 128 * 128 + 1024>>>(map, d_rows, d_out, 128 * 128);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<uint16_t> o(128 * 64);
    cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad_valid = 0, bad_invalid = 0;
    for (int i = 0; i < 128; ++i) {
      const bool valid = rows[i] >= 0 && rows[i] < n_rows;
      for (int c = 0; c < C; ++c) {
        const uint16_t want = valid ? h[(size_t)rows[i] * C + c] : 0;
        if (o[i * 64 + c] != want) (valid ? bad_valid : bad_invalid)++;
      }
    }
    printf("  mismatching elements: valid rows %d, invalid rows (expected zero fill) %d\n", bad_valid, bad_invalid);
  }
  return 0;
}
