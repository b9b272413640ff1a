// Shared helpers for the lidal_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/lidal_b200.h"

namespace lb {

void set_error(const char* fmt, ...);

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#define LB_CHECK_ARG(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      lb::set_error("%s: invalid argument: %s", __func__, msg);   \
      return LB_EINVAL;                                           \
    }                                                             \
  } while (0)

#define LB_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      lb::set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(_e),     \
                    __FILE__, __LINE__);                                                \
      return LB_ECUDA;                                                                  \
    }                                                                                   \
  } while (0)

// every kernel launch site is followed by LB_LAUNCHED(n): n kernels were just enqueued (feeds lb_launch_count)
void add_launches(int n);
#define LB_LAUNCHED(n) lb::add_launches(n)
#define LB_LAUNCH_CHECK() LB_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached per device).
int sm_count();

// 64-bit FNV-1a over four uint32 words folded to 60 bits (torchsparse 1.4.0 backend/hash).
__host__ __device__ __forceinline__ int64_t fnv60(int x, int y, int z, int b) {
  uint64_t h = 14695981039346656037ULL;
  h ^= (uint32_t)x; h *= 1099511628211ULL;
  h ^= (uint32_t)y; h *= 1099511628211ULL;
  h ^= (uint32_t)z; h *= 1099511628211ULL;
  h ^= (uint32_t)b; h *= 1099511628211ULL;
  h = (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFFULL);
  return (int64_t)h;
}

// ---- open-addressing hash table: [cap x int64 key][cap x int32 row] ----
struct TableView {
  unsigned long long* keys;
  int* rows;
  uint64_t mask;
};
__host__ __device__ __forceinline__ uint64_t table_capacity(int64_t n) {
  uint64_t cap = 1024;
  while (cap < (uint64_t)(2 * n)) cap <<= 1;
  return cap;
}
__host__ __device__ __forceinline__ TableView table_view(const void* table, size_t bytes) {
  uint64_t cap = bytes / 12;
  TableView t;
  t.keys = (unsigned long long*)table;
  t.rows = (int*)((char*)table + cap * 8);
  t.mask = cap - 1;
  return t;
}
__device__ __forceinline__ uint64_t mix64(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
  return h;
}
#define LB_EMPTY_KEY 0xFFFFFFFFFFFFFFFFULL
__device__ __forceinline__ void table_insert(const TableView& t, uint64_t key, int row) {
  uint64_t slot = mix64(key) & t.mask;
  while (true) {
    unsigned long long prev = atomicCAS(&t.keys[slot], LB_EMPTY_KEY, (unsigned long long)key);
    if (prev == LB_EMPTY_KEY || prev == key) {
      atomicMin(&t.rows[slot], row);
      return;
    }
    slot = (slot + 1) & t.mask;
  }
}
__device__ __forceinline__ int table_find(const TableView& t, uint64_t key) {
  uint64_t slot = mix64(key) & t.mask;
  while (true) {
    unsigned long long k = __ldg(&t.keys[slot]);
    if (k == key) return __ldg(&t.rows[slot]);
    if (k == LB_EMPTY_KEY) return -1;
    slot = (slot + 1) & t.mask;
  }
}

// device-wide exclusive scan of uint32 (scan.cu); ws >= scan_ws_bytes(n)
size_t scan_ws_bytes(int64_t n);
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total /*device, optional*/, void* ws,
                       cudaStream_t st);

}  // namespace lb
