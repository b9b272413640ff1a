// Sparse convolution as an output-stationary implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   out[o, :] = epilogue( sum_{k in offsets} in[nbr[k][o], :] @ W[k] )
//
// One persistent CTA per SM walks 128- or 256-row output tiles.  Warp roles (416 threads):
//   warps 0-3  epilogue : tcgen05.ld accumulator rows from TMEM -> scale/shift/residual/ReLU -> global store
//   warps 4-11 producers: gather the A operand -- 128 neighbour rows x BLOCK_K channels -- straight into the
//                         swizzled K-major shared-memory layout with 16-byte cp.async (zero-fill for missing
//                         neighbours); thread 128 also issues the TMA load of the weight tile W[k][:, c0:c0+BK]
//   warp  12   MMA      : one elected thread issues tcgen05.mma (M=128, N=c_out, K=16) per 16 channels,
//                         accumulating over all active offsets and channel blocks in TMEM (fp32)
// Producers never block on their own copies: each thread posts `cp.async.mbarrier.arrive.noinc` on the stage's full
// barrier, which the hardware fires when that thread's copies have landed, so the whole ring depth is in flight.
// Pipelines: a STAGES-deep smem ring (full/empty mbarriers) between producers and MMA, and a 2-deep TMEM
// accumulator ring (tmem_full/tmem_empty) between MMA and epilogue, so the epilogue of tile t overlaps the
// gathers and MMAs of tile t+1.  Offsets for which no row of the tile has a neighbour are skipped entirely.
// Tiles are handed out by a global ticket counter (dynamic scheduler, expensive tiles first); the tile id and its
// stage count travel through the ring in one smem word with the tile's first stage (producers -> MMA warp ->
// epilogue via tstart barriers), an all-ones word is the end-of-work sentinel.
// No atomics in the data path: every output row is written exactly once, by one thread.
//
// Measured on B200 (ncu, round 1): the kernel is bound by instructions per ring stage in the producer warps and in
// the single issuing warp, not by DRAM, L2 or the tensor pipe on the 32/96-channel layers -- hence the care taken to
// keep both per-stage paths short (running addresses, uniform control flow through REDUX, elect.sync issue).
#include <type_traits>
#include "tc_ptx.cuh"

namespace lb {

constexpr int TILE_M = 128;
constexpr int NUM_EPI_THREADS = 256;     // two epilogue groups of four warps (TMEM lane quadrant = warp % 4)
constexpr int EPI_WARPS = NUM_EPI_THREADS / 32;
#ifndef LIDAL_PROD_WARPS
#define LIDAL_PROD_WARPS 8
#endif
constexpr int NUM_PROD_THREADS = 32 * LIDAL_PROD_WARPS;   // 8 gather warps: two per scheduler, so dependent address/LDS/cp.async chains overlap
static_assert(LIDAL_PROD_WARPS == 8 || LIDAL_PROD_WARPS == 16, "the thread -> row mapping of the gather warps assumes 8 or 16 warps");
constexpr int NUM_THREADS = NUM_EPI_THREADS + NUM_PROD_THREADS + 64;
constexpr int MMA_WARP = (NUM_EPI_THREADS + NUM_PROD_THREADS) / 32;
constexpr int WEIGHT_WARP = MMA_WARP + 1;   // issues the TMA weight-tile loads (and arms the stage barriers) so that the gather
                                            // warps' per-stage instruction stream carries nothing but the gather itself
constexpr int PROD_BAR_THREADS = NUM_PROD_THREADS + 32;   // named barrier 1: the gather warps + the weight warp
constexpr int MAX_STAGES = 12;
constexpr int MAX_KVOL = 27;

struct TcParams {
  const char* in;          // 16-bit activations
  int64_t n_in, ld_in;     // ld in elements
  char* out;
  int64_t n_out, ld_out;
  const int* n_out_dev;
  const int* nbr;
  int64_t nbr_ld;
  const int* out_rows;
  int k_vol, c_in, c_out;
  const float* scale;
  const float* shift;
  const char* residual;
  int64_t ld_res;
  int out_dtype;           // LB_DT_*
  int relu;
  int is_bf16;
  int stages;
  int tmem_cols;
  int nb;                  // (offset, channel-block) blocks carried by one pipeline stage (1..4)
  int T;                   // 128-row sub-tiles per CTA tile (1 or 2): T accumulators share every weight tile
  int staged;              // smem-staged epilogue (16-bit output): 1 = per-row bulk async copies (any row order),
                           // 2 = TMA tile boxes [32 rows x 32 channels] through out_map / res_map (rows in natural order)
  int stg_bufs;            // staging buffers per epilogue warp (2: the stores of one sub-tile drain while the next is built)
  int n_acc;               // TMEM accumulator sets (2 = epilogue overlaps the next tile, 1 when 2*T*c_out > 512)
  int pack8;               // LB_CONV_PACK8: K axis = (offset, 8 channels), 8 offsets per 64-wide K block
  int wwarp;               // 1: the weight warp loads the weight tiles (lock-step cp.async producer); 0: producer thread 0 does
  int dbg;                 // LIDAL_DBG knock-out bits for bottleneck hunting (results are WRONG when set): 1 = no gather copies,
                           // 2 = no MMA issue, 4 = no output stores
  int zero_row, zero_mask; // prod_mode 3: rows [zero_row, zero_row + zero_mask] of the input tensor are all zero: absent neighbours
                           // are gathered from there (spread over the pool so no single L2 line is hammered)
  int prod_warps;          // prod_mode 1: warps that own stages = min(8, stages) -- a warp's consecutive stages must be at most
                           // one ring lap apart, or its parity wait on the empty barrier could be satisfied by an older phase
  int prod_mode;           // 0: all producer warps fill every stage in lock-step (cp.async); 1: each producer warp owns whole stages
                           // (cp.async); 2: owned stages, register-staged LDG.128 -> STS.128 with a writer-side proxy fence;
                           // 3: TMA tile::gather4, all warps on every stage (16 rows of each sub-tile per warp)
                           // (stage q belongs to warp q % prod_warps), so eight dependent fill chains run side by side
  unsigned* sched;         // dynamic tile scheduler: [0] next ticket, [1] retired CTAs (both zero between launches);
                           // nullptr = static round-robin
  const uint32_t* tile_masks;   // lean producer: active-offset mask of every 128-row group of the (mask-sorted) neighbour table
  int lean;                // 0: per-tile prologue through s_idx (modes above); 1: prologue-free gather (tile masks, indices read
                           // straight from the table, static snake schedule); 2: identity input rows -> the A operand arrives by TMA
                           // tile loads issued by the weight warp, the gather warps stay idle
  int lean_arrive;         // lean == 1: 0 = every gather thread posts cp.async.mbarrier.arrive.noinc (256 arrivals per stage);
                           // 1 = commit groups, a stage is published `lean_lag` stages later by ONE arrival per warp after
                           // cp.async.wait_group + a writer-side proxy fence
  int lean_lag;            // stages a warp keeps unpublished (1..3, < stages - 1)
  int idx_bytes;           // shared memory reserved for s_idx (0 in lean mode: the bytes go to the ring)
  int fastpro;             // short per-tile prologue: offset masks from tile_masks, index rows staged unchecked (see the gather warps)
};

// Bottleneck-hunting switches (knock-outs, cycle accounting) exist only in a -DLIDAL_CONV_DEBUG build: in the production
// kernel every `DBG(p)` test folds to zero and the instrumentation disappears.
#ifdef LIDAL_CONV_DEBUG
#define DBG(p) ((p).dbg)
#else
#define DBG(p) 0
#endif
// LIDAL_DBG & 128: cycle accounting of CTA 0 (lane 0 of one warp per role) into the scheduler cell, printed by the host
#define DBG_ON ((DBG(p) & 128) && p.sched && blockIdx.x == 0)
#define DBG_ADD(slot, val) atomicAdd(reinterpret_cast<unsigned long long*>(p.sched + 16) + (slot), (unsigned long long)(val))

template <typename T> __device__ __forceinline__ float cvt_in(uint16_t raw);
template <> __device__ __forceinline__ float cvt_in<__nv_bfloat16>(uint16_t raw) { return __uint_as_float((uint32_t)raw << 16); }
template <> __device__ __forceinline__ float cvt_in<__half>(uint16_t raw) { return __half2float(__ushort_as_half(raw)); }

__device__ __forceinline__ uint32_t pack2(float a, float b, int is_bf16) {
  if (is_bf16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// BK = channels per pipeline stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
// KMAX = compile-time bound of the per-tile offset loops (1, 8 or 27): the index registers and the mask votes are unrolled
// over it, so a 1x1 or 2x2x2 layer does not pay the 27-offset prologue on every tile.
template <int BK, int T, int KMAX>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap out_map,
               const __grid_constant__ CUtensorMap res_map, const __grid_constant__ CUtensorMap in_map, const TcParams p) {
  constexpr int ROW_BYTES = BK * 2;
  constexpr int CHUNKS = ROW_BYTES / 16;                 // 16-byte chunks per row: 8 or 4
  constexpr int A_BYTES = TILE_M * ROW_BYTES;            // 16 KB or 8 KB
  constexpr uint32_t LAYOUT = (BK == 64) ? 2u : 4u;      // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t SBO = 8 * ROW_BYTES;                // 8-row group pitch
  constexpr int ROWS_PER_PASS = NUM_PROD_THREADS / CHUNKS;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align by OFFSET (not by integer round-trip of the pointer) so the compiler keeps the shared address space -> LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = p.c_out * ROW_BYTES;
  const int b_pad = (b_bytes + 1023) & ~1023;
  const int a_blk = T * A_BYTES;                        // A operand of one block: T sub-tiles of 128 rows
  const int stage_bytes = p.nb * (a_blk + b_pad);         // [nb A blocks][nb B sub-tiles]
  constexpr int TM = T * TILE_M;                          // output rows per CTA tile
  uint8_t* ring = smem;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  int* s_idx = (int*)tail;                                         // [MAX_KVOL][T * TILE_M]
  float* s_scale = (float*)(tail + p.idx_bytes);                   // [256]
  float* s_shift = s_scale + 256;                                  // [256]
  uint64_t* full_bar = (uint64_t*)(s_shift + 256);                 // [MAX_STAGES]
  uint64_t* empty_bar = full_bar + MAX_STAGES;                     // [MAX_STAGES]
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;                    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                            // [2]
  uint32_t* s_tmem = (uint32_t*)(tempty_bar + 2) + MAX_STAGES;     // [1] (MAX_STAGES words of slack kept in front: tail_bytes() layout)
  uint32_t* s_mask = s_tmem + 1;                                   // [2] active-offset mask of the producers' tile (by parity)
  uint64_t* res_bar = (uint64_t*)(s_mask + 2 + 1);                 // [8] residual rows landed (one per epilogue warp), 8-byte aligned
  uint64_t* tstart_bar = res_bar + EPI_WARPS;                              // [2] tile id of an accumulator set published (MMA warp -> epilogue)
  int* s_next = (int*)(tstart_bar + 2);                            // [1] next ticket of this CTA (-1: none), producers only
  int* s_stage_tile = s_next + 2;                                  // [MAX_STAGES] tile id carried by a tile's first stage
  int* s_acc_tile = s_stage_tile + MAX_STAGES;                     // [2] tile id held by each accumulator set (-1: no more work)
  int* s_dbg = s_acc_tile + 2;                                     // [2 * MAX_STAGES] debug build only: clock of a stage's gather issue / MMA commit
  const int stg_pitch = p.c_out * 2 + 16;                          // staged epilogue: row pitch (+16 B: conflict-free 128-bit LDS)
  uint8_t* s_stage = tail + (((size_t)p.idx_bytes + 2 * 256 * 4 + (2 * MAX_STAGES + 4) * 8 + MAX_STAGES * 4 + 256 + 1023) & ~(size_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_out = p.n_out_dev ? (int64_t)*p.n_out_dev : p.n_out;
  const int64_t num_tiles = (n_out + TM - 1) / TM;
  const int kc_blocks = p.c_in / BK;
  // Work is handed out as tickets.  Static mode: CTA b owns tickets b, b + grid, ...  and ticket == tile.  Dynamic mode
  // (p.sched): the first ticket is b, later ones come from a global counter, and tickets walk the tiles from the LAST one
  // down -- with mask-sorted rows the late tiles carry the most offsets, so the expensive tiles start first and the cheap
  // ones fill the tail (longest-processing-time order).
  auto tile_of = [&](int64_t ticket) -> int64_t { return p.sched ? num_tiles - 1 - ticket : ticket; };

  // ---------------- one-time setup
  for (int i = threadIdx.x; i < 256; i += NUM_THREADS) {
    s_scale[i] = (p.scale && i < p.c_out) ? p.scale[i] : 1.f;
    s_shift[i] = (p.shift && i < p.c_out) ? p.shift[i] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      // every gather thread of the stage posts one asynchronous arrival + 1 expect_tx arrive for the TMA weight tile
      mbar_init(&full_bar[s], p.lean == 2 ? 1 : (p.lean == 1 && p.lean_arrive) ? LIDAL_PROD_WARPS + 1
                              : p.prod_mode == 3 ? 1 : (p.prod_mode >= 1 ? 32 : NUM_PROD_THREADS) + 1);
      mbar_init(&empty_bar[s], 1);                     // released by tcgen05.commit
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&res_bar[w], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tstart_bar[a], 1);
      mbar_init(&tempty_bar[a], T == 2 ? NUM_EPI_THREADS : NUM_EPI_THREADS / 2);   // every epilogue thread that drains the set arrives once per tile
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // ---------------- lean mode: no per-tile prologue.  Every role derives the same static tile sequence on its own -- round i
  // of CTA b takes ticket i * grid + (i odd ? grid - 1 - b : b) (a snake, so every CTA gets expensive and cheap tiles alike)
  // and tickets walk the mask-sorted tiles from the last (most offsets) down -- and reads the tile's offset mask from the
  // table lb_kmap_tile_masks built once per map.  Nothing is staged in shared memory and no CTA-wide barrier is left on the
  // gather path; the roles meet only at the ring's full / empty barriers.
  const int n_groups128 = (int)((n_out + TILE_M - 1) / TILE_M);
  auto lean_tile = [&](int round) -> int {
    const int64_t tk = (int64_t)round * gridDim.x + ((round & 1) ? (int64_t)(gridDim.x - 1 - blockIdx.x) : (int64_t)blockIdx.x);
    return tk < num_tiles ? (int)(num_tiles - 1 - tk) : -1;
  };
  auto lean_mask = [&](int tile) -> uint32_t {
    if (p.lean == 2) return 1u;
    uint32_t m = __ldg(p.tile_masks + tile * T);
    if (T == 2 && tile * 2 + 1 < n_groups128) m |= __ldg(p.tile_masks + tile * 2 + 1);
    if (p.k_vol < 32) m &= (1u << p.k_vol) - 1u;
    return m ? m : 1u;                                      // keep the pipeline uniform: one all-zero block
  };

  if (p.lean == 1 && warp >= EPI_WARPS && warp < MMA_WARP) {
    // =============================================================== LEAN GATHER WARPS
    // Thread t copies the 16-byte chunk (t % CHUNKS) of rows row0 + i * ROWS_PER_PASS of every 128-row sub-tile.  The row
    // indices of an offset come straight from the neighbour table (one 4-byte load per slot, all CHUNKS lanes of a row read
    // the same word) and are fetched TWO offsets ahead of their use, across tile boundaries, so the copy stream never waits
    // for them; the tile's offset mask is fetched one tile ahead.
    const int t = threadIdx.x - NUM_EPI_THREADS;
    const int chunk = t % CHUNKS, row0 = t / CHUNKS;
    constexpr int PASSES = TILE_M / ROWS_PER_PASS;
    constexpr int NS = T * PASSES;                          // gather slots of a thread per block
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t st_u32 = ring_u32;
    int stage = 0;
    uint32_t ph = 0;
    const int64_t ld_in_b = p.ld_in * 2;
    const char* in_col = p.in + chunk * 16;
    const int n_out_i = (int)n_out;
    uint32_t dst_off[NS];                                   // smem offset of each slot inside a block (swizzled chunk position)
#pragma unroll
    for (int sub = 0; sub < T; ++sub)
#pragma unroll
      for (int i = 0; i < PASSES; ++i) {
        const int r = row0 + i * ROWS_PER_PASS;
        const uint32_t sw = (BK == 64) ? (uint32_t)(chunk ^ (r & 7)) : (uint32_t)(chunk ^ ((r >> 1) & 3));
        dst_off[sub * PASSES + i] = (uint32_t)(sub * A_BYTES + r * ROW_BYTES) + sw * 16;
      }
    // iterator over the (tile, offset) pairs of this CTA, in issue order
    int round = 0;
    int it_tile = lean_tile(0);
    uint32_t it_mask = it_tile >= 0 ? lean_mask(it_tile) : 0u;
    int nx_tile = lean_tile(1);
    uint32_t nx_mask = nx_tile >= 0 ? lean_mask(nx_tile) : 0u;
    bool it_first = true;
    // descriptor of a fetched pair: -1 = no more work; else bit 30 = first pair of its tile, bits 8.. = offsets of the tile
    auto fetch = [&](int (&dst)[NS]) -> int {
      if (it_mask == 0u) {
        it_tile = nx_tile;
        it_mask = nx_mask;
        it_first = true;
        if (it_tile >= 0) {
          ++round;
          nx_tile = lean_tile(round + 1);
          nx_mask = nx_tile >= 0 ? lean_mask(nx_tile) : 0u;
        }
      }
      if (it_tile < 0) return -1;
      const int k = __ffs(it_mask) - 1;
      const int desc = (it_first ? (1 << 30) | (__popc(it_mask) << 8) : 0);
      it_first = false;
      it_mask &= it_mask - 1u;
      const int* src = p.nbr + (int64_t)k * p.nbr_ld + (int64_t)it_tile * TM + row0;
#pragma unroll
      for (int sub = 0; sub < T; ++sub)
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int o = sub * TILE_M + i * ROWS_PER_PASS;
          dst[sub * PASSES + i] = (it_tile * TM + row0 + o < n_out_i) ? __ldg(src + o) : -1;
        }
      return desc;
    };
    int nb_a[NS], nb_b[NS], nb_c[NS];
    int d_a = fetch(nb_a);
    int d_b = d_a >= 0 ? fetch(nb_b) : -1;
    int d_c = d_b >= 0 ? fetch(nb_c) : -1;
    int j = 0, remaining = 0;
    int pend = 0, arr_stage = 0;                            // lean_arrive: committed but unpublished stages, oldest of them
    while (d_a >= 0) {
      if (d_a & (1 << 30)) remaining = ((d_a >> 8) & 0xff) * kc_blocks;
      for (int cb = 0; cb < kc_blocks; ++cb) {
        if (j == 0) mbar_wait(&empty_bar[stage], ph ^ 1);   // slot free (first lap passes immediately)
        const char* in_cb = in_col + cb * (BK * 2);
        const uint32_t blk_u32 = st_u32 + (uint32_t)(j * a_blk);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const int nb = nb_a[s];
          cp_async16(blk_u32 + dst_off[s], in_cb + (int64_t)(nb >= 0 ? nb : 0) * ld_in_b, nb >= 0 ? 16u : 0u);
        }
        --remaining;
        if (++j == p.nb || remaining == 0) {
          if (p.lean_arrive) {
            cp_async_commit();
            if (++pend > p.lean_lag) {
              if (p.lean_lag == 1) cp_async_wait<1>(); else if (p.lean_lag == 2) cp_async_wait<2>(); else cp_async_wait<3>();
              fence_proxy_async();                          // my landed copies -> visible to the tensor core's async proxy
              __syncwarp();
              if (lane == 0) mbar_arrive(&full_bar[arr_stage]);
              if (++arr_stage == p.stages) arr_stage = 0;
              --pend;
            }
          } else {
            cp_async_arrive_noinc(&full_bar[stage]);        // asynchronous: fires when this thread's copies have landed
          }
          j = 0;
          st_u32 += stage_bytes;
          if (++stage == p.stages) { stage = 0; ph ^= 1; st_u32 = ring_u32; }
        }
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) { nb_a[s] = nb_b[s]; nb_b[s] = nb_c[s]; }
      d_a = d_b;
      d_b = d_c;
      d_c = d_c >= 0 ? fetch(nb_c) : -1;
    }
    if (p.lean_arrive) {                                    // publish what is still pending
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      for (; pend > 0; --pend) {
        if (lane == 0) mbar_arrive(&full_bar[arr_stage]);
        if (++arr_stage == p.stages) arr_stage = 0;
      }
    }
    // sentinel stage (the weight warp writes its tile word): complete the barrier's arrival count
    mbar_wait(&empty_bar[stage], ph ^ 1);
    if (!p.lean_arrive || lane == 0) mbar_arrive(&full_bar[stage]);
    cp_async_wait<0>();
  } else if (p.lean && warp == WEIGHT_WARP) {
    // =============================================================== LEAN WEIGHT / TILE LOADER
    // Opens every stage: arms the full barrier with the stage's TMA byte count, publishes the tile word with a tile's first
    // stage and issues the weight-tile loads; for identity input rows (lean == 2) also the A operand as plain TMA tile loads
    // (box = BK channels x 128 rows, same swizzle as the gather would produce), so such layers run without gather warps.
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&w_map) : "memory");
      if (p.lean == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
    }
    int stage = 0;
    uint32_t ph = 0, st_u32 = smem_u32(ring);
    const uint32_t ring_u32 = st_u32;
    const uint32_t blk_tx = (uint32_t)b_bytes + (p.lean == 2 ? (uint32_t)a_blk : 0u);
    int round = 0;
    int tile = lean_tile(0);
    uint32_t mask = tile >= 0 ? lean_mask(tile) : 0u;
    while (tile >= 0) {
      const int ntile = lean_tile(round + 1);               // next tile's mask: in flight while this tile is issued
      const uint32_t nmask = ntile >= 0 ? lean_mask(ntile) : 0u;
      int remaining = __popc(mask) * kc_blocks;
      const int cur_word = (int)((uint32_t)tile | ((uint32_t)remaining << 24));
      bool first = true;
      int j = 0;
      for (uint32_t m = mask; m; m &= m - 1u) {
        const int k = __ffs(m) - 1;
        for (int cb = 0; cb < kc_blocks; ++cb) {
          if (j == 0) {
            mbar_wait(&empty_bar[stage], ph ^ 1);
            if (lane == 0) {
              if (first) s_stage_tile[stage] = cur_word;
              const int in_stage = remaining < p.nb ? remaining : p.nb;
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)in_stage * blk_tx);
            }
            first = false;
          }
          if (lane == 0) {
            tma_load_2d(st_u32 + (uint32_t)(p.nb * a_blk + j * b_pad), &w_map, cb * BK, k * p.c_out, &full_bar[stage]);
            if (p.lean == 2) {
#pragma unroll
              for (int sub = 0; sub < T; ++sub)
                tma_load_2d(st_u32 + (uint32_t)(j * a_blk + sub * A_BYTES), &in_map, cb * BK, tile * TM + sub * TILE_M, &full_bar[stage]);
            }
          }
          --remaining;
          if (++j == p.nb || remaining == 0) {
            j = 0;
            st_u32 += stage_bytes;
            if (++stage == p.stages) { stage = 0; ph ^= 1; st_u32 = ring_u32; }
          }
        }
      }
      tile = ntile;
      mask = nmask;
      ++round;
    }
    mbar_wait(&empty_bar[stage], ph ^ 1);
    if (lane == 0) {
      s_stage_tile[stage] = -1;                             // sentinel: this CTA is out of work
      mbar_arrive(&full_bar[stage]);
    }
  } else if (p.lean && warp >= EPI_WARPS && warp < MMA_WARP) {
    // lean == 2: the A operand comes by TMA, the gather warps have nothing to do
  } else if (warp >= EPI_WARPS && warp < MMA_WARP) {
    // =============================================================== PRODUCERS
    const int t = threadIdx.x - NUM_EPI_THREADS;          // 0..127
    const int chunk = t % CHUNKS, row0 = t / CHUNKS;
    if (t == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&w_map) : "memory");
    int stage = 0;                                        // ring position of the open stage
    uint32_t ph = 0;
    // One stage = 128*T gathered rows x BK channels of one offset + its weight tile.  Every instruction of this path
    // runs once per stage in each of the 8 producer warps, so it is kept minimal: running stage address, precomputed
    // per-thread smem offsets, no per-stage bookkeeping beyond the flag word.
    int cur_word = 0;                                     // tile being issued | its stage count << 24 (published with its first stage)
    const int pw = warp - NUM_EPI_THREADS / 32;           // producer warp 0..7
    int seq = 0;                                          // (stages issued so far) % prod_warps -- prod_mode 1: owner of the open stage
    constexpr int PASSES = TILE_M / ROWS_PER_PASS;
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t st_u32 = ring_u32;                           // smem address of the open stage
    const int64_t ld_in_b = p.ld_in * 2;
    const char* in_col = p.in + (p.pack8 ? 0 : chunk * 16);
    auto issue = [&](const int (&nbv)[T][PASSES], int cb, int b_col, int b_row, uint32_t first_last) {
      const long long dbg_w = (DBG_ON && t == 0) ? clock64() : 0;
      mbar_wait(&empty_bar[stage], ph ^ 1);               // slot free (first lap passes immediately)
      if (DBG_ON && t == 0) DBG_ADD(5, clock64() - dbg_w);
      if (t == 0 && !p.wwarp) {
        if (first_last & 1u) s_stage_tile[stage] = cur_word;     // tile id | stage count << 24, read once per tile
        if (DBG(p) & 16) mbar_arrive(&full_bar[stage]);
        else {
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)b_bytes);
          tma_load_2d(st_u32 + a_blk, &w_map, b_col, b_row, &full_bar[stage]);
        }
      }
      const char* in_cb = in_col + (p.pack8 ? 0 : cb * (BK * 2));
#pragma unroll
      for (int sub = 0; sub < T; ++sub) {
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int r = row0 + i * ROWS_PER_PASS;
          const int nb = nbv[sub][i];
          // pack8: 16 bytes = the 8 (padded) channels of one offset; otherwise channels [cb*BK + chunk*8, +8) of the row
          const char* src = in_cb + (int64_t)(nb >= 0 ? nb : 0) * ld_in_b;
          const uint32_t sw = (BK == 64) ? (uint32_t)(chunk ^ (r & 7)) : (uint32_t)(chunk ^ ((r >> 1) & 3));
          if (!(DBG(p) & 1)) cp_async16(st_u32 + sub * A_BYTES + r * ROW_BYTES + sw * 16, src, nb >= 0 ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(&full_bar[stage]);            // asynchronous: fires when this thread's copies have landed
      st_u32 += stage_bytes;
      if (++stage == p.stages) { stage = 0; ph ^= 1; st_u32 = ring_u32; }
    };
    // Neighbour indices of the NEXT tile are fetched into registers while the current tile's stages are being issued
    // (27 independent loads in flight), so the tile prologue never waits on global memory.
    // Work item q = (offset k, group of 4 consecutive tile rows): one 16-byte load from the table (when its rows are
    // 16-byte aligned: nbr_ld % 4 == 0), one 16-byte store into s_idx -- a quarter of the instructions of a load per row.
    constexpr int R4 = TM / 4;                            // 4-row groups per offset
    constexpr int ITEMS = (KMAX * R4 + NUM_PROD_THREADS - 1) / NUM_PROD_THREADS;
    int4 nb_reg[ITEMS];
    const bool vec_ok = p.nbr && (p.nbr_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.nbr) & 15) == 0;
    // Short prologue (p.fastpro: the caller supplied the tile-mask table, so the table is one this library built and its
    // entries are -1 or valid rows): the tile's offset mask is ONE prefetched word instead of compares, a warp reduction, a
    // shared-memory atomic and a re-read; the index rows of full tiles are fetched with per-thread constant offsets (running
    // pointer, no division, no bounds tests) and staged as they are.  ~60 instructions per tile instead of ~350 -- on the
    // 32-channel layers the prologue was longer than the tile's main loop.
    constexpr int KJ = NUM_PROD_THREADS / R4;             // offsets advanced per work item
    const int k_t = t / R4;                               // offset of this thread's item 0
    const int64_t f_step = (int64_t)KJ * p.nbr_ld;        // ints between consecutive items of a thread
    const int* f_base = p.nbr ? p.nbr + (int64_t)k_t * p.nbr_ld + (t % R4) * 4 : nullptr;
    int* const st_base = s_idx + k_t * TM + (t % R4) * 4;
    uint32_t mask_next = 0;                               // fastpro: mask of the tile whose indices sit in nb_reg
    const int n_groups128_p = (int)((n_out + TILE_M - 1) / TILE_M);
    auto table_mask = [&](int64_t tile) -> uint32_t {
      uint32_t m = __ldg(p.tile_masks + tile * T);
      if (T == 2 && tile * 2 + 1 < n_groups128_p) m |= __ldg(p.tile_masks + tile * 2 + 1);
      if (p.k_vol < 32) m &= (1u << p.k_vol) - 1u;
      return m;
    };
    auto fetch_indices = [&](int64_t tile) {
      const int64_t o0 = tile * TM;
      if (p.fastpro) {
        mask_next = table_mask(tile);
        if (o0 + TM <= n_out) {                           // full tile: no bounds tests
          const int* src = f_base + o0;
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            if (k_t + j * KJ < KMAX && k_t + j * KJ < p.k_vol) nb_reg[j] = __ldg(reinterpret_cast<const int4*>(src));
            src += f_step;
          }
          return;
        }
      }
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int q = t + j * NUM_PROD_THREADS;
        const int k = q / R4;
        const int64_t o = o0 + (q % R4) * 4;
        int4 v = make_int4(-1, -1, -1, -1);
        if (k < KMAX && k < p.k_vol && o < n_out) {
          if (!p.nbr || (DBG(p) & 64)) {
            v.x = (int)o;
            if (o + 1 < n_out) v.y = (int)o + 1;
            if (o + 2 < n_out) v.z = (int)o + 2;
            if (o + 3 < n_out) v.w = (int)o + 3;
          } else {
            const int* src = p.nbr + (int64_t)k * p.nbr_ld + o;
            if (vec_ok && o + 3 < n_out) {
              v = __ldg(reinterpret_cast<const int4*>(src));
            } else {
              v.x = __ldg(src);
              if (o + 1 < n_out) v.y = __ldg(src + 1);
              if (o + 2 < n_out) v.z = __ldg(src + 2);
              if (o + 3 < n_out) v.w = __ldg(src + 3);
            }
          }
        }
        nb_reg[j] = v;
      }
    };
    const uint32_t n_in_u = p.n_in > 0x7fffffff ? 0x7fffffffu : (uint32_t)p.n_in;
    int64_t cur = blockIdx.x;                             // this CTA's current ticket
    if (cur < num_tiles) fetch_indices(tile_of(cur));
    if (t < 2) s_mask[t] = 0;
    unsigned ahead = 0;                                   // thread 0: ticket drawn one tile ahead (its latency is hidden)
    if (t == 0 && p.sched && cur < num_tiles) ahead = atomicAdd(&p.sched[0], 1u);
    uint32_t par = 0;
    const bool dbg_me = DBG_ON && t == 0;
    const long long dbg_p0 = dbg_me ? clock64() : 0;
    for (; cur < num_tiles; par ^= 1) {
      const int64_t tile = tile_of(cur);
      // (A) every producer has finished reading s_idx / s_mask[par^1] of the previous tile
      long long dbg_t = dbg_me ? clock64() : 0;
      asm volatile("bar.sync 1, %0;" ::"n"(PROD_BAR_THREADS) : "memory");
      if (dbg_me) { DBG_ADD(6, clock64() - dbg_t); DBG_ADD(8, 1); }
      uint32_t my_bits = 0;
      if (p.fastpro) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
          if (k_t + j * KJ < KMAX && k_t + j * KJ < p.k_vol) *reinterpret_cast<int4*>(st_base + j * KJ * TM) = nb_reg[j];
      } else
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int q = t + j * NUM_PROD_THREADS;
        const int k = q / R4;
        if (k < KMAX && k < p.k_vol) {
          int4 v = nb_reg[j];
          // one unsigned compare rejects both "no neighbour" (-1) and out-of-range rows
          const bool ox = (uint32_t)v.x < n_in_u, oy = (uint32_t)v.y < n_in_u, oz = (uint32_t)v.z < n_in_u, ow = (uint32_t)v.w < n_in_u;
          if (p.prod_mode == 3) {                          // absent neighbour -> a row of the zero pool (TMA cannot skip rows)
            const int z = p.zero_row + ((q * 4 + k * 13) & p.zero_mask);
            v.x = ox ? v.x : z; v.y = oy ? v.y : p.zero_row + ((z + 1) & p.zero_mask);
            v.z = oz ? v.z : p.zero_row + ((z + 2) & p.zero_mask); v.w = ow ? v.w : p.zero_row + ((z + 3) & p.zero_mask);
          } else {
            v.x = ox ? v.x : -1; v.y = oy ? v.y : -1; v.z = oz ? v.z : -1; v.w = ow ? v.w : -1;
          }
          *reinterpret_cast<int4*>(&s_idx[k * TM + (q % R4) * 4]) = v;
          if (ox | oy | oz | ow) my_bits |= 1u << k;
        }
      }
      if (!p.fastpro) {
        const uint32_t my_mask = __reduce_or_sync(0xffffffffu, my_bits);   // one warp reduction instead of a vote per offset
        if (lane == 0 && my_mask) atomicOr(&s_mask[par], my_mask);
      }
      if (t == 0) {
        s_mask[par ^ 1] = 0;
        const int64_t nx = p.sched ? (int64_t)gridDim.x + ahead : cur + gridDim.x;
        s_next[0] = nx < num_tiles ? (int)nx : -1;
        if (p.sched && nx < num_tiles) ahead = atomicAdd(&p.sched[0], 1u);
      }
      // (B) indices and mask of this tile are complete
      dbg_t = dbg_me ? clock64() : 0;
      asm volatile("bar.sync 1, %0;" ::"n"(PROD_BAR_THREADS) : "memory");
      if (dbg_me) DBG_ADD(7, clock64() - dbg_t);
      uint32_t mask = __reduce_or_sync(0xffffffffu, p.fastpro ? mask_next : s_mask[par]);   // same value in every lane; REDUX makes it provably uniform
      if (mask == 0) mask = 1;                            // keep the pipeline uniform: one all-zero k-block
      const int nx = (int)__reduce_or_sync(0xffffffffu, (uint32_t)s_next[0]);   // stable until barrier (A) of the next tile
      if (nx >= 0) fetch_indices(tile_of(nx));
      int nbv[T][PASSES];                                 // neighbour rows of this thread's gather slots, one offset at a time
      if (p.pack8) {
        const int nkb = (p.k_vol * 8 + BK - 1) / BK;
        cur_word = (int)((uint32_t)tile | ((uint32_t)nkb << 24));
        for (int kb = 0; kb < nkb; ++kb) {
          const int kk = kb * 8 + chunk;                  // chunk <-> offset 8*kb + chunk
#pragma unroll
          for (int sub = 0; sub < T; ++sub)
#pragma unroll
            for (int i = 0; i < PASSES; ++i)
              nbv[sub][i] = kk < p.k_vol ? s_idx[kk * TM + sub * TILE_M + row0 + i * ROWS_PER_PASS] : -1;
          issue(nbv, kb, kb * BK, 0, (kb == 0 ? 1u : 0u) | (kb == nkb - 1 ? 2u : 0u));
        }
      } else if (p.prod_mode == 3) {
        // TMA gather, all producer warps on every stage: warp w brings rows [16w, 16w + 16) of each 128-row sub-tile with
        // four tile::gather4 instructions (lanes 0-3; lanes 4-7 serve the second sub-tile), so a stage is issued in the
        // time of 4-8 TMA instructions per warp.  Thread 0 arms the stage's barrier with the byte count of the whole stage
        // (gathered rows + weight tile); absent neighbours were redirected to the zero pool when the indices were staged.
        int remaining = __popc(mask) * kc_blocks;
        cur_word = (int)((uint32_t)tile | ((uint32_t)remaining << 24));
        bool first = true;
        const int g_sub = lane >> 2, g_row = pw * 16 + (lane & 3) * 4;
        for (int k = __ffs(mask) - 1; k < 32 && (mask >> k); ++k) {
          if (!((mask >> k) & 1)) continue;
          int4 v = make_int4(0, 0, 0, 0);
          if (lane < 4 * T) v = *reinterpret_cast<const int4*>(&s_idx[k * TM + g_sub * TILE_M + g_row]);
          for (int cb = 0; cb < kc_blocks; ++cb, --remaining, first = false) {
            mbar_wait(&empty_bar[stage], ph ^ 1);
            if (t == 0) {
              if (first) s_stage_tile[stage] = cur_word;
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(a_blk + b_bytes));
              tma_load_2d(st_u32 + a_blk, &w_map, cb * BK, k * p.c_out, &full_bar[stage]);
            }
            if (lane < 4 * T && !(DBG(p) & 1))
              tma_gather4(st_u32 + g_sub * A_BYTES + g_row * ROW_BYTES, &in_map, cb * BK, v.x, v.y, v.z, v.w, &full_bar[stage]);
            st_u32 += stage_bytes;
            if (++stage == p.stages) { stage = 0; ph ^= 1; st_u32 = ring_u32; }
          }
        }
      } else if (p.prod_mode >= 1) {
        // Per-warp stage ownership: the CTA's stages are numbered in issue order and stage q is filled entirely by
        // producer warp q % prod_warps (lane = 16-byte chunk of a row x a group of consecutive rows, four row indices per LDS.128).
        // Every warp walks the same (offset, channel block) sequence and skips the stages it does not own, so eight
        // fill chains -- slot wait, index loads, address arithmetic, cp.async issue -- overlap instead of one.
        int remaining = __popc(mask) * kc_blocks;
        cur_word = (int)((uint32_t)tile | ((uint32_t)remaining << 24));
        bool first = true;
        constexpr int G = 32 / CHUNKS;                     // row groups per warp instruction: 4 (BK=64) or 8 (BK=32)
        constexpr int RPL = TM / G;                        // consecutive rows owned by a lane
        const int lchunk = lane % CHUNKS, lgrp = lane / CHUNKS;
        const uint32_t lane_dst = (uint32_t)((lgrp * RPL) / TILE_M) * A_BYTES + (uint32_t)((lgrp * RPL) % TILE_M) * ROW_BYTES;
        const char* lane_src = p.in + lchunk * 16;
        for (int k = __ffs(mask) - 1; k < 32 && (mask >> k); ++k) {
          if (!((mask >> k) & 1)) continue;
          for (int cb = 0; cb < kc_blocks; ++cb, --remaining, first = false) {
            if (seq == pw) {
              const long long dbg_w = dbg_me ? clock64() : 0;
              mbar_wait(&empty_bar[stage], ph ^ 1);
              if (dbg_me) DBG_ADD(5, clock64() - dbg_w);
              if (lane == 0) {
                if (first) s_stage_tile[stage] = cur_word;
                if (DBG(p) & 16) mbar_arrive(&full_bar[stage]);
                else {
                  mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)b_bytes);
                  tma_load_2d(st_u32 + a_blk, &w_map, cb * BK, k * p.c_out, &full_bar[stage]);
                }
              }
              const char* src_cb = lane_src + cb * (BK * 2);
              const int4* idx4 = reinterpret_cast<const int4*>(&s_idx[k * TM + lgrp * RPL]);
              const uint32_t dst0 = st_u32 + lane_dst;
              if (p.prod_mode == 2) {
                // Register-staged gather: 8 independent 16-byte loads in flight per lane, then 8 shared-memory stores
                // (absent rows store zeros and load nothing).  Measured on B200 (tools/experiments/gather_bw.cu): LDG + STS
                // moves twice the bytes per warp that cp.async does for scattered 16-byte pieces.  The stores go through the
                // generic proxy, so the writer fences towards the async proxy (tcgen05 reads) before it arrives.
#pragma unroll 1
                for (int j8 = 0; j8 < RPL / 8; ++j8) {
                  const int4 va = idx4[2 * j8], vb = idx4[2 * j8 + 1];
                  const int nb8[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
                  uint4 v[8];
#pragma unroll
                  for (int u = 0; u < 8; ++u)
                    v[u] = nb8[u] >= 0 ? __ldg(reinterpret_cast<const uint4*>(src_cb + (int64_t)nb8[u] * ld_in_b)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    const int j = j8 * 8 + u;
                    const uint32_t sw = (BK == 64) ? (uint32_t)(lchunk ^ (j & 7)) : (uint32_t)(lchunk ^ ((j >> 1) & 3));
                    if (!(DBG(p) & 1))
                      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + (uint32_t)j * ROW_BYTES + sw * 16), "r"(v[u].x),
                                   "r"(v[u].y), "r"(v[u].z), "r"(v[u].w)
                                   : "memory");
                  }
                }
                fence_proxy_async();
                mbar_arrive(&full_bar[stage]);
              } else {
#pragma unroll 4
              for (int j4 = 0; j4 < RPL / 4; ++j4) {
                const int4 v = idx4[j4];
                const int nb4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int j = j4 * 4 + u;                // row within the lane's group; (lgrp * RPL) % 8 == 0
                  const uint32_t sw = (BK == 64) ? (uint32_t)(lchunk ^ (j & 7)) : (uint32_t)(lchunk ^ ((j >> 1) & 3));
                  const int nb = nb4[u];
                  if (!(DBG(p) & 1)) cp_async16(dst0 + (uint32_t)j * ROW_BYTES + sw * 16, src_cb + (int64_t)(nb >= 0 ? nb : 0) * ld_in_b, nb >= 0 ? 16u : 0u);
                }
              }
              cp_async_arrive_noinc(&full_bar[stage]);
              }
            }
            if (++seq == p.prod_warps) seq = 0;
            st_u32 += stage_bytes;
            if (++stage == p.stages) { stage = 0; ph ^= 1; st_u32 = ring_u32; }
          }
        }
      } else if (p.wwarp) {
        // Lock-step gather trimmed to the instructions that cannot be avoided.  (ncu source view, round 2: every role of this
        // kernel is a chain of dependent instructions that issues one every ~10 cycles, so the time of a ring stage is the
        // LENGTH of the longest per-stage chain.  The gather warps ran ~150 instructions per block -- 64-bit address
        // arithmetic per slot and block, four branches of stage bookkeeping; now ~30.)  Per offset: each slot's source row
        // pointer and fill size are formed once (one IMAD.WIDE per slot).  Per block: T * PASSES LDGSTS whose global and
        // shared offsets are immediates (the channel-block loop is unrolled for c_in / BK in {1, 2, 3, 4, 6}).  Per stage of
        // p.nb blocks: one slot wait, one arrival, a branch-free ring advance on running 32-bit addresses.
        constexpr int NS = T * PASSES;
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        uint32_t bar_off = (uint32_t)stage * 8u;
        const uint32_t bar_end = (uint32_t)p.stages * 8u;
        const uint32_t ld32 = (uint32_t)ld_in_b;                               // row pitch in bytes (< 2^32: ld_in is a row width)
        const uint32_t sw0 = (BK == 64) ? (uint32_t)(chunk ^ (row0 & 7)) : (uint32_t)(chunk ^ ((row0 >> 1) & 3));   // ROWS_PER_PASS % 8 == 0
        const uint32_t t_off = (uint32_t)(row0 * ROW_BYTES) + sw0 * 16u;
        const int* idx_t = s_idx + row0;
        int remaining = __popc(mask) * kc_blocks;                              // blocks of this tile still to issue
        int j = 0;                                                             // block slot inside the open stage
        uint32_t dst = st_u32 + t_off;                                         // this thread's first slot of the open block
        auto close_block = [&]() {
          --remaining;
          dst += (uint32_t)a_blk;
          if (++j == p.nb || remaining == 0) {
            cp_async_arrive_noinc_u32(full0 + bar_off);                        // asynchronous: fires when this thread's copies have landed
            if (DBG_ON && t == 0) s_dbg[bar_off >> 3] = (int)clock();
            j = 0;
            st_u32 += stage_bytes;
            bar_off += 8u;
            const bool wrap = bar_off == bar_end;
            bar_off = wrap ? 0u : bar_off;
            st_u32 = wrap ? ring_u32 : st_u32;
            ph ^= wrap ? 1u : 0u;
            dst = st_u32 + t_off;
          }
        };
        auto open_block = [&]() {
          if (j == 0) {
            const int dbg_c0 = (DBG_ON && t == 0) ? (int)clock() : 0;
            mbar_wait_u32(empty0 + bar_off, ph ^ 1);                           // slot free (first lap passes immediately)
            if (DBG_ON && t == 0) {
              const int now = (int)clock();
              DBG_ADD(5, now - dbg_c0);
              if (now - dbg_c0 > 150) { DBG_ADD(13, now - s_dbg[MAX_STAGES + (bar_off >> 3)]); DBG_ADD(14, 1); }   // blocked: MMA commit -> slot seen free
            }
          }
        };
        auto one_offset = [&](auto kc_c, int k) {
          constexpr int KC = decltype(kc_c)::value;
          const char* src[NS];
          uint32_t sz[NS];
#pragma unroll
          for (int sub = 0; sub < T; ++sub)
#pragma unroll
            for (int i = 0; i < PASSES; ++i) {
              const int nb = idx_t[k * TM + sub * TILE_M + i * ROWS_PER_PASS];
              sz[sub * PASSES + i] = nb >= 0 ? 16u : 0u;
              src[sub * PASSES + i] = in_col + (uint64_t)(uint32_t)(nb >= 0 ? nb : 0) * ld32;
            }
#pragma unroll
          for (int cb = 0; cb < KC; ++cb) {
            open_block();
            const int dbg_g0 = (DBG_ON && t == 0) ? (int)clock() : 0;
#pragma unroll
            for (int sub = 0; sub < T; ++sub)
#pragma unroll
              for (int i = 0; i < PASSES; ++i)
                if (!(DBG(p) & 1))
                  cp_async16(dst + (uint32_t)(sub * A_BYTES + i * ROWS_PER_PASS * ROW_BYTES), src[sub * PASSES + i] + cb * (BK * 2), sz[sub * PASSES + i]);
            const int dbg_g1 = (DBG_ON && t == 0) ? (int)clock() : 0;
            close_block();
            if (DBG_ON && t == 0) { DBG_ADD(19, dbg_g1 - dbg_g0); DBG_ADD(20, (int)clock() - dbg_g1); }
          }
        };
        for (uint32_t m = mask; m; m &= m - 1u) {
          const int k = __ffs(m) - 1;
          switch (kc_blocks) {
            case 1: one_offset(std::integral_constant<int, 1>{}, k); break;
            case 2: one_offset(std::integral_constant<int, 2>{}, k); break;
            case 3: one_offset(std::integral_constant<int, 3>{}, k); break;
            case 4: one_offset(std::integral_constant<int, 4>{}, k); break;
            case 6: one_offset(std::integral_constant<int, 6>{}, k); break;
            default:
              for (int cb = 0; cb < kc_blocks; ++cb) {                       // other channel counts: same sequence, runtime offsets
                open_block();
#pragma unroll
                for (int sub = 0; sub < T; ++sub)
#pragma unroll
                  for (int i = 0; i < PASSES; ++i) {
                    const int nb = idx_t[k * TM + sub * TILE_M + i * ROWS_PER_PASS];
                    cp_async16(dst + (uint32_t)(sub * A_BYTES + i * ROWS_PER_PASS * ROW_BYTES),
                               in_col + (uint64_t)(uint32_t)(nb >= 0 ? nb : 0) * ld32 + cb * (BK * 2), nb >= 0 ? 16u : 0u);
                  }
                close_block();
              }
          }
        }
        stage = (int)(bar_off >> 3);
      } else {
        int remaining = __popc(mask) * kc_blocks;
        cur_word = (int)((uint32_t)tile | ((uint32_t)remaining << 24));
        bool first = true;
        for (int k = __ffs(mask) - 1; k < 32 && (mask >> k); ++k) {
          if (!((mask >> k) & 1)) continue;
#pragma unroll
          for (int sub = 0; sub < T; ++sub)
#pragma unroll
            for (int i = 0; i < PASSES; ++i) nbv[sub][i] = s_idx[k * TM + sub * TILE_M + row0 + i * ROWS_PER_PASS];
          for (int cb = 0; cb < kc_blocks; ++cb, --remaining, first = false)
            issue(nbv, cb, cb * BK, k * p.c_out, (first ? 1u : 0u) | (remaining == 1 ? 2u : 0u));
        }
      }
      cur = nx >= 0 ? nx : num_tiles;
    }
    if (dbg_me) DBG_ADD(4, clock64() - dbg_p0);
    // sentinel stage: no data, tile word -1 tells the MMA warp (and through it the epilogue) that this CTA is out of work
    if (p.prod_mode == 3) {
      mbar_wait(&empty_bar[stage], ph ^ 1);
      if (t == 0) {
        s_stage_tile[stage] = -1;
        mbar_arrive(&full_bar[stage]);                      // the barrier expects one arrival (thread 0's expect_tx)
      }
    } else if (p.prod_mode >= 1) {
      if (seq == pw) {
        mbar_wait(&empty_bar[stage], ph ^ 1);
        if (lane == 0) {
          s_stage_tile[stage] = -1;
          mbar_arrive(&full_bar[stage]);                    // stands in for the expect_tx arrival of a normal stage
        }
        mbar_arrive(&full_bar[stage]);
      }
    } else {
      mbar_wait(&empty_bar[stage], ph ^ 1);
      if (t == 0 && !p.wwarp) {
        s_stage_tile[stage] = -1;
        mbar_arrive(&full_bar[stage]);                      // stands in for the expect_tx arrival of a normal stage
      }
      mbar_arrive(&full_bar[stage]);
    }
    cp_async_wait<0>();                                   // nothing of ours may still be in flight at teardown
  } else if (warp == WEIGHT_WARP) {
    // =============================================================== WEIGHT LOADER
    // Walks the same (tile, offset, channel block) sequence as the gather warps: it joins their two per-tile barriers to
    // learn the tile's offset mask and the next ticket, and for every stage arms the full barrier with the weight tile's
    // byte count and issues the TMA load.  The tile word (id | stage count) travels with the tile's first stage.
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&w_map) : "memory");
    int stage = 0;
    uint32_t ph = 0, st_u32 = smem_u32(ring);
    const uint32_t ring_u32 = st_u32;
    int64_t cur = blockIdx.x;
    uint32_t par = 0;
    const int n_groups128_w = (int)((n_out + TILE_M - 1) / TILE_M);
    auto table_mask = [&](int64_t tile) -> uint32_t {
      uint32_t m = __ldg(p.tile_masks + tile * T);
      if (T == 2 && tile * 2 + 1 < n_groups128_w) m |= __ldg(p.tile_masks + tile * 2 + 1);
      if (p.k_vol < 32) m &= (1u << p.k_vol) - 1u;
      return m;
    };
    uint32_t mask_w = (p.fastpro && cur < num_tiles) ? table_mask(tile_of(cur)) : 0u;   // fastpro: mask of the current tile, fetched a tile ahead
    for (; cur < num_tiles; par ^= 1) {
      const int64_t tile = tile_of(cur);
      asm volatile("bar.sync 1, %0;" ::"n"(PROD_BAR_THREADS) : "memory");   // (A)
      asm volatile("bar.sync 1, %0;" ::"n"(PROD_BAR_THREADS) : "memory");   // (B)
      uint32_t mask = __reduce_or_sync(0xffffffffu, p.fastpro ? mask_w : s_mask[par]);
      if (mask == 0) mask = 1;
      const int nx = (int)__reduce_or_sync(0xffffffffu, (uint32_t)s_next[0]);
      if (p.fastpro && nx >= 0) mask_w = table_mask(tile_of(nx));
      if (p.wwarp) {
        int remaining = __popc(mask) * kc_blocks;         // blocks of this tile
        const int cur_word = (int)((uint32_t)tile | ((uint32_t)remaining << 24));
        bool first = true;
        int j = 0;
        for (int k = __ffs(mask) - 1; k < 32 && (mask >> k); ++k) {
          if (!((mask >> k) & 1)) continue;
          for (int cb = 0; cb < kc_blocks; ++cb) {
            if (j == 0) {                                 // open a stage: arm its barrier with all of its weight bytes
              mbar_wait(&empty_bar[stage], ph ^ 1);
              if (lane == 0) {
                if (first) s_stage_tile[stage] = cur_word;
                const int in_stage = remaining < p.nb ? remaining : p.nb;
                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(in_stage * b_bytes));
              }
              first = false;
            }
            if (lane == 0) tma_load_2d(st_u32 + (uint32_t)(p.nb * a_blk + j * b_pad), &w_map, cb * BK, k * p.c_out, &full_bar[stage]);
            --remaining;
            if (++j == p.nb || remaining == 0) {
              j = 0;
              st_u32 += stage_bytes;
              if (++stage == p.stages) { stage = 0; ph ^= 1; st_u32 = ring_u32; }
            }
          }
        }
      }
      cur = nx >= 0 ? nx : num_tiles;
    }
    if (p.wwarp) {
      mbar_wait(&empty_bar[stage], ph ^ 1);
      if (lane == 0) {
        s_stage_tile[stage] = -1;                         // sentinel: this CTA is out of work
        mbar_arrive(&full_bar[stage]);
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================================================== MMA ISSUER
    // The issuing warp is a single instruction stream: every instruction here is on the critical path of every stage,
    // so descriptors are advanced with 32-bit adds on their low word and nothing is rebuilt per stage.
    const uint32_t idesc = make_idesc(TILE_M, p.c_out, p.is_bf16 ? 1 : 0);
    const uint32_t desc_hi = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);             // bits 32-45 SBO, 46 version, 61-63 swizzle
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);      // bits 0-13 address >> 4, 16-29 LBO = 1
    const uint32_t stage_units = (uint32_t)stage_bytes >> 4, b_units = (uint32_t)(p.nb * a_blk) >> 4, b_pad_units = (uint32_t)b_pad >> 4;
    constexpr uint32_t A_UNITS = A_BYTES >> 4;
    // The whole issue loop runs in ONE thread (lane 0): waits, tile hand-over, MMAs and commits need no warp-level
    // agreement, so there is no per-stage elect / warp sync and the barrier addresses are running 32-bit registers.
    // (elect.sync rather than `lane == 0`: the compiler then knows the region is single-threaded and issues the UTCHMMA /
    // UTCBAR instructions back to back instead of wrapping each one in an election loop.)
    if (elect_one()) {
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t bar_end = (uint32_t)p.stages * 8u;
    uint32_t bar_off = 0;
    uint32_t ph = 0, st_lo = ring_lo;
    const bool need_fence = p.prod_mode < 2 && !(p.lean == 2 || (p.lean == 1 && p.lean_arrive));
    const bool dbg_me = DBG_ON;
    const long long dbg_m0 = dbg_me ? clock64() : 0;
    for (int64_t tcount = 0;; ++tcount) {
      const int acc = p.n_acc == 2 ? (int)(tcount & 1) : 0;
      const uint32_t acc_ph = (uint32_t)((p.n_acc == 2 ? tcount >> 1 : tcount) & 1);
      long long dbg_t = dbg_me ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1);            // epilogue drained this accumulator set
      if (dbg_me) DBG_ADD(2, clock64() - dbg_t);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T * p.c_out);
      // first stage of the tile: its slot carries the tile word (tile id | block count << 24; -1 = this CTA is done)
      dbg_t = dbg_me ? clock64() : 0;
      mbar_wait_u32(full0 + bar_off, ph);
      if (dbg_me) DBG_ADD(1, clock64() - dbg_t);
      const uint32_t word = (uint32_t)*(volatile int*)&s_stage_tile[bar_off >> 3];
      if (word == 0xffffffffu) {
        s_acc_tile[acc] = -1;
        mbar_arrive(&tstart_bar[acc]);
        if (T == 1) {                                      // 128-row tiles: the other epilogue group owns the other accumulator set
          const int acc2 = (int)((tcount + 1) & 1);
          mbar_wait(&tempty_bar[acc2], (uint32_t)(((tcount + 1) >> 1) & 1) ^ 1);
          s_acc_tile[acc2] = -1;
          mbar_arrive(&tstart_bar[acc2]);
        }
        if (dbg_me) DBG_ADD(0, clock64() - dbg_m0);
        break;
      }
      const int n_blk = (int)(word >> 24);                       // blocks of the tile; a stage carries up to p.nb of them
      s_acc_tile[acc] = (int)(word & 0xffffffu);                 // the epilogue learns its tile before the MMAs finish
      mbar_arrive(&tstart_bar[acc]);
      if (dbg_me) DBG_ADD(3, n_blk);
      for (int e = 0, done = 0; done < n_blk; ++e) {
        const int nbs = (n_blk - done) < p.nb ? (n_blk - done) : p.nb;
        if (e) {
          dbg_t = dbg_me ? clock64() : 0;
          mbar_wait_u32(full0 + bar_off, ph);
          if (dbg_me) {
            const long long now = clock64();
            DBG_ADD(1, now - dbg_t);
            if (now - dbg_t > 150) { DBG_ADD(12, (int)now - s_dbg[bar_off >> 3]); DBG_ADD(15, 1); }   // blocked: gather issue -> data seen landed
          }
        }
        const long long dbg_s1 = dbg_me ? clock64() : 0;
        if (need_fence && !(DBG(p) & 8)) fence_proxy_async();   // cp.async wrote through the generic proxy and cannot fence on the writer side
        tc_fence_after();
        const long long dbg_s2 = dbg_me ? clock64() : 0;
#pragma unroll 1
        for (int j = 0; j < nbs; ++j) {                   // the blocks of this stage, back to back
          const uint32_t a_lo = st_lo + (uint32_t)j * (uint32_t)T * A_UNITS;
          const uint32_t b_lo = st_lo + b_units + (uint32_t)j * b_pad_units;
#pragma unroll
          for (int sub = 0; sub < T; ++sub) {             // every sub-tile reuses the same weight tile
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)          // +32 bytes along K inside the swizzle atom = +2 in the address field
              if (!(DBG(p) & 2)) umma_f16_lohi(d_tmem + (uint32_t)(sub * p.c_out), a_lo + (uint32_t)sub * A_UNITS + (uint32_t)kk * 2u,
                            b_lo + (uint32_t)kk * 2u, desc_hi, idesc, (kk == 0 && e == 0 && j == 0) ? 0u : 1u);   // first MMA overwrites
          }
        }
        const long long dbg_s3 = dbg_me ? clock64() : 0;
        umma_commit_u32(empty0 + bar_off);                // smem slot reusable once these MMAs retire
        if (done + nbs == n_blk) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
        if (dbg_me) {
          s_dbg[MAX_STAGES + (bar_off >> 3)] = (int)clock();
          DBG_ADD(16, dbg_s2 - dbg_s1); DBG_ADD(17, dbg_s3 - dbg_s2); DBG_ADD(18, clock64() - dbg_s3);
        }
        done += nbs;
        st_lo += stage_units;
        bar_off += 8u;
        if (bar_off == bar_end) { bar_off = 0; ph ^= 1; st_lo = ring_lo; }
      }
    }
    }
    __syncwarp();
  } else {
    // =============================================================== EPILOGUE (warps 0-7: TMEM lane quadrant = warp % 4)
    // Two groups of four warps.  With 256-row tiles (T == 2) group g drains sub-tile g of every tile; with 128-row tiles
    // group g owns accumulator set g, i.e. every other tile.  Either way two epilogue chains -- row-index load, residual
    // fetch, TMEM reads, conversion, stores -- run side by side: on the fine levels a tile's epilogue (fixed cost per row)
    // is as long as its few-offset main loop, so a single chain was a co-bottleneck of the kernel.
    const int eq = warp & 3, eg = warp >> 2;
    constexpr bool SPLIT_SUB = (T == 2);
    const int sub_lo = SPLIT_SUB ? eg : 0, sub_hi = SPLIT_SUB ? eg + 1 : T;
    const int64_t t_first = SPLIT_SUB ? 0 : eg, t_step = SPLIT_SUB ? 1 : 2;
    if (p.staged == 2) {
      // TMA-tile staging (output rows in natural order, 1x1 layers): a warp's 32 rows x c_out channels sit in smem as
      // c_out/32 boxes of [32 rows][64 B] in the SWIZZLE_64B pattern (16-byte chunk q of row r at q ^ ((r >> 1) & 3):
      // conflict-free 128-bit accesses without padding).  The residual tile arrives by one TMA load per box, the finished
      // tile leaves by one TMA store per box -- c_out/32 copy instructions per warp instead of 32 per-row copies, which
      // the uniform datapath issues one lane at a time.
      const int groups = p.c_out >> 5;
      const uint32_t buf_bytes = (uint32_t)groups * 2048u;
      uint8_t* my_stage = s_stage + (size_t)warp * p.stg_bufs * buf_bytes;
      const uint32_t sw_row = (uint32_t)lane * 64u, sw_x = (uint32_t)(lane >> 1) & 3u;
      const uint32_t row_bytes = (uint32_t)p.c_out * 2;
      uint32_t res_ph = 0;
      int buf = 0;
      int64_t tcount = t_first;
      for (;; tcount += t_step) {
        const int acc = p.n_acc == 2 ? (int)(tcount & 1) : 0;
        const uint32_t acc_ph = (uint32_t)((p.n_acc == 2 ? tcount >> 1 : tcount) & 1);
        mbar_wait(&tstart_bar[acc], acc_ph);              // the MMA warp has published this accumulator's tile
        const int64_t tile = s_acc_tile[acc];
        if (tile < 0) break;
        for (int sub = sub_lo; sub < sub_hi; ++sub) {
          const int64_t row0 = tile * TM + sub * TILE_M + eq * 32;
          const bool any_live = row0 < n_out;
          uint8_t* sbuf = my_stage + (size_t)buf * buf_bytes;
          if (lane == 0) {                                  // lane 0 owns the warp's bulk groups
            if (p.stg_bufs == 2) bulk_wait_read1(); else bulk_wait_read();
          }
          if (p.stg_bufs == 2) buf ^= 1;
          __syncwarp();
          if (p.residual && any_live && lane == 0) {
            mbar_arrive_expect_tx(&res_bar[warp], 32u * row_bytes);      // whole boxes; rows past the end are zero-filled
            for (int g = 0; g < groups; ++g) tma_load_2d(smem_u32(sbuf + g * 2048), &res_map, g * 32, (int)row0, &res_bar[warp]);
          }
          if (sub == sub_lo) {
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
          }
          if (p.residual && any_live) {
            mbar_wait(&res_bar[warp], res_ph);
            res_ph ^= 1;
          }
          for (int c0 = 0; c0 < p.c_out; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)((acc * T + sub) * p.c_out + c0), v);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * s_scale[c0 + j] + s_shift[c0 + j];
            if (p.relu == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            uint8_t* grp = sbuf + (size_t)(c0 >> 5) * 2048 + sw_row;
            if (p.residual) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 rr = *(const uint4*)(grp + (((uint32_t)q ^ sw_x) << 4));
                const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (p.is_bf16) {
                    f[q * 8 + 2 * j] += cvt_in<__nv_bfloat16>((uint16_t)(w[j] & 0xffff));
                    f[q * 8 + 2 * j + 1] += cvt_in<__nv_bfloat16>((uint16_t)(w[j] >> 16));
                  } else {
                    f[q * 8 + 2 * j] += cvt_in<__half>((uint16_t)(w[j] & 0xffff));
                    f[q * 8 + 2 * j + 1] += cvt_in<__half>((uint16_t)(w[j] >> 16));
                  }
                }
              }
            }
            if (p.relu == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack2(f[q * 8], f[q * 8 + 1], p.is_bf16);
              w.y = pack2(f[q * 8 + 2], f[q * 8 + 3], p.is_bf16);
              w.z = pack2(f[q * 8 + 4], f[q * 8 + 5], p.is_bf16);
              w.w = pack2(f[q * 8 + 6], f[q * 8 + 7], p.is_bf16);
              *(uint4*)(grp + (((uint32_t)q ^ sw_x) << 4)) = w;
            }
          }
          fence_proxy_async();                              // generic-proxy smem writes -> visible to the copy engine
          __syncwarp();
          if (lane == 0) {
            if (any_live && !(DBG(p) & 4))
              for (int g = 0; g < groups; ++g) tma_store_2d(&out_map, g * 32, (int)row0, smem_u32(sbuf + g * 2048));
            bulk_commit();
          }
        }   // sub-tiles
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
      }
      if (lane == 0) bulk_wait_all();
    } else if (p.staged) {
      // Shared-memory-staged epilogue: every lane owns one output row.  Its residual row arrives by ONE bulk async copy
      // (completion on the warp's mbarrier), the row is finished in place in smem, and leaves by ONE bulk async store:
      // global traffic is whole contiguous rows moved by the copy engine instead of 16-byte pieces per thread.
      uint8_t* my_stage = s_stage + (size_t)warp * p.stg_bufs * 32 * stg_pitch;
      const size_t buf_bytes = (size_t)32 * stg_pitch;
      int buf = 0;
      const uint32_t row_bytes = (uint32_t)p.c_out * 2;
      uint32_t res_ph = 0;
      int64_t tcount = t_first;
      const bool dbg_me = DBG_ON && threadIdx.x == 0;
      const long long dbg_e0 = dbg_me ? clock64() : 0;
      for (;; tcount += t_step) {
        const int acc = p.n_acc == 2 ? (int)(tcount & 1) : 0;
        const uint32_t acc_ph = (uint32_t)((p.n_acc == 2 ? tcount >> 1 : tcount) & 1);
        const long long dbg_t = dbg_me ? clock64() : 0;
        mbar_wait(&tstart_bar[acc], acc_ph);              // the MMA warp has published this accumulator's tile
        if (dbg_me) DBG_ADD(10, clock64() - dbg_t);
        const int64_t tile = s_acc_tile[acc];
        if (tile < 0) { if (dbg_me) DBG_ADD(9, clock64() - dbg_e0); break; }
        if (DBG(p) & 32) { mbar_wait(&tfull_bar[acc], acc_ph); tc_fence_after(); }
        else
        for (int sub = sub_lo; sub < sub_hi; ++sub) {
          const int64_t o = tile * TM + sub * TILE_M + eq * 32 + lane;
          const bool live = o < n_out;
          const int64_t orow = live ? (p.out_rows ? (int64_t)__ldg(&p.out_rows[o]) : o) : 0;
          uint8_t* my_row = my_stage + buf * buf_bytes + (size_t)lane * stg_pitch;
          if (p.stg_bufs == 2) { bulk_wait_read1(); buf ^= 1; }   // the stores that last used THIS buffer have read it
          else bulk_wait_read();
          __syncwarp();
          if (p.residual) {
            const unsigned live_mask = __ballot_sync(0xffffffffu, live);
            if (lane == 0) mbar_arrive_expect_tx(&res_bar[warp], row_bytes * (uint32_t)__popc(live_mask));
            __syncwarp();
            if (live) bulk_load(smem_u32(my_row), p.residual + orow * p.ld_res * 2, row_bytes, &res_bar[warp]);
          }
          if (sub == sub_lo) {
            const long long dbg_f = dbg_me ? clock64() : 0;
            mbar_wait(&tfull_bar[acc], acc_ph);
            if (dbg_me) DBG_ADD(11, clock64() - dbg_f);
            tc_fence_after();
          }
          if (p.residual) {
            mbar_wait(&res_bar[warp], res_ph);
            res_ph ^= 1;
          }
          for (int c0 = 0; c0 < p.c_out; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)((acc * T + sub) * p.c_out + c0), v);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * s_scale[c0 + j] + s_shift[c0 + j];
            if (p.relu == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            uint4* chunk = (uint4*)(my_row + c0 * 2);
            if (p.residual) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 rr = chunk[q];
                const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (p.is_bf16) {
                    f[q * 8 + 2 * j] += cvt_in<__nv_bfloat16>((uint16_t)(w[j] & 0xffff));
                    f[q * 8 + 2 * j + 1] += cvt_in<__nv_bfloat16>((uint16_t)(w[j] >> 16));
                  } else {
                    f[q * 8 + 2 * j] += cvt_in<__half>((uint16_t)(w[j] & 0xffff));
                    f[q * 8 + 2 * j + 1] += cvt_in<__half>((uint16_t)(w[j] >> 16));
                  }
                }
              }
            }
            if (p.relu == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack2(f[q * 8], f[q * 8 + 1], p.is_bf16);
              w.y = pack2(f[q * 8 + 2], f[q * 8 + 3], p.is_bf16);
              w.z = pack2(f[q * 8 + 4], f[q * 8 + 5], p.is_bf16);
              w.w = pack2(f[q * 8 + 6], f[q * 8 + 7], p.is_bf16);
              chunk[q] = w;
            }
          }
          fence_proxy_async();                            // my generic-proxy smem writes -> visible to the bulk-copy engine
          if (live && !(DBG(p) & 4)) bulk_store(p.out + orow * p.ld_out * 2, smem_u32(my_row), row_bytes);
          bulk_commit();
        }   // sub-tiles
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
      }
      bulk_wait_all();                                    // all rows are in global memory before the CTA retires
    } else {
    int64_t tcount = t_first;
    const int out_es = (p.out_dtype == LB_DT_F32) ? 4 : 2;
    for (;; tcount += t_step) {
      const int acc = p.n_acc == 2 ? (int)(tcount & 1) : 0;
      const uint32_t acc_ph = (uint32_t)((p.n_acc == 2 ? tcount >> 1 : tcount) & 1);
      mbar_wait(&tstart_bar[acc], acc_ph);                // the MMA warp has published this accumulator's tile
      const int64_t tile = s_acc_tile[acc];
      if (tile < 0) break;
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      for (int sub = sub_lo; sub < sub_hi; ++sub) {
      const int64_t o = tile * TM + sub * TILE_M + eq * 32 + lane;
      const bool live = o < n_out;
      const int64_t orow = live ? (p.out_rows ? (int64_t)__ldg(&p.out_rows[o]) : o) : 0;
      char* out_row = p.out + orow * p.ld_out * out_es;
      const char* res_row = p.residual ? p.residual + orow * p.ld_res * 2 : nullptr;
      for (int c0 = 0; c0 < p.c_out; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)((acc * T + sub) * p.c_out + c0), v);
        if (live) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * s_scale[c0 + j] + s_shift[c0 + j];
          if (p.relu == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (res_row) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 rr = __ldg((const uint4*)(res_row + (c0 + q * 8) * 2));
              const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (p.is_bf16) {
                  f[q * 8 + 2 * j] += cvt_in<__nv_bfloat16>((uint16_t)(w[j] & 0xffff));
                  f[q * 8 + 2 * j + 1] += cvt_in<__nv_bfloat16>((uint16_t)(w[j] >> 16));
                } else {
                  f[q * 8 + 2 * j] += cvt_in<__half>((uint16_t)(w[j] & 0xffff));
                  f[q * 8 + 2 * j + 1] += cvt_in<__half>((uint16_t)(w[j] >> 16));
                }
              }
            }
          }
          if (p.relu == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.out_dtype == LB_DT_F32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *(float4*)(out_row + (c0 + q * 4) * 4) = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack2(f[q * 8], f[q * 8 + 1], p.is_bf16);
              w.y = pack2(f[q * 8 + 2], f[q * 8 + 3], p.is_bf16);
              w.z = pack2(f[q * 8 + 4], f[q * 8 + 5], p.is_bf16);
              w.w = pack2(f[q * 8 + 6], f[q * 8 + 7], p.is_bf16);
              *(uint4*)(out_row + (c0 + q * 8) * 2) = w;
            }
          }
        }
      }
      }   // sub-tiles
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  // the last CTA to retire re-arms the scheduler for the next launch on this stream (nobody draws tickets any more)
  if (p.sched && threadIdx.x == 0 && atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
    p.sched[0] = 0;
    p.sched[1] = 0;
    __threadfence();
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

static inline int block_k_for(int c_in) { return (c_in % 64 == 0) ? 64 : ((c_in % 32 == 0) ? 32 : 0); }
int conv_tc_pack8_supported(int k_vol, int c_in, int c_out, int act_dtype) {
  return (act_dtype == LB_DT_BF16 || act_dtype == LB_DT_F16) && c_in == 8 && k_vol >= 1 && k_vol <= MAX_KVOL && c_out % 32 == 0 && c_out >= 32 && c_out <= 256;
}

int conv_tc_supported(int k_vol, int c_in, int c_out, int act_dtype) {
  if (act_dtype != LB_DT_BF16 && act_dtype != LB_DT_F16) return 0;
  if (k_vol < 1 || k_vol > MAX_KVOL) return 0;
  if (block_k_for(c_in) == 0 || c_in > 512) return 0;   // a tile's stage count (<= 27 * c_in / 64) travels in 8 bits
  if (c_out % 32 != 0 || c_out < 32 || c_out > 256) return 0;
  return 1;
}

static size_t idx_bytes(int T, int lean) { return lean ? 0 : (size_t)MAX_KVOL * T * TILE_M * 4; }
static size_t tail_bytes(int T, int lean = 0) {
  // indices (none in lean mode) + scale/shift + ring/accumulator barriers + flags + (tmem ptr, masks, 4 residual barriers), rounded for staging
  return ((idx_bytes(T, lean) + 2 * 256 * 4 + (2 * MAX_STAGES + 4) * 8 + MAX_STAGES * 4 + 256 + 1023) & ~(size_t)1023);
}
static size_t staging_bytes(int c_out, int bufs = 1) { return (size_t)bufs * EPI_WARPS * 32 * (c_out * 2 + 16); }   // one buffer per epilogue warp

int conv_tc_launch(const lb_conv_args& a, cudaStream_t st) {
  const bool pack8 = (a.flags & LB_CONV_PACK8) != 0;
  const int bk = pack8 ? 64 : block_k_for(a.c_in);
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("lb_conv_fwd: cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return LB_ECUDA; }
  if (a.out_dtype != LB_DT_F32 && a.out_dtype != a.act_dtype) {
    set_error("lb_conv_fwd: out_dtype must be F32 or equal act_dtype");
    return LB_EINVAL;
  }
  const int out_es = a.out_dtype == LB_DT_F32 ? 4 : 2;
  if (((uintptr_t)a.out & 15) || (a.ld_out * out_es) % 16 || (a.residual && (((uintptr_t)a.residual & 15) || (a.ld_res * 2) % 16))) {
    set_error("lb_conv_fwd: tensor-core path needs 16-byte aligned out / residual rows");
    return LB_EINVAL;
  }
  // weight viewed as a 2-D tensor [k_vol * c_out rows, c_in cols] (c_in contiguous), box = [c_out rows, bk cols]
  CUtensorMap map;
  // PACK8: weight is [c_out rows, kpad cols] with col = offset * 8 + channel, kpad = k_vol * 8 rounded up to 64
  const int kpad = ((a.k_vol * 8 + 63) / 64) * 64;
  cuuint64_t gdim[2] = {(cuuint64_t)(pack8 ? kpad : a.c_in), (cuuint64_t)(pack8 ? a.c_out : a.k_vol * a.c_out)};
  cuuint64_t gstride[1] = {(cuuint64_t)(pack8 ? kpad : a.c_in) * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)a.c_out};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&map, a.act_dtype == LB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                      const_cast<void*>(a.weight), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("lb_conv_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r); return LB_ECUDA; }

  // Lean producer (no per-tile prologue): gathered layers need the tile-mask table of lb_kmap_tile_masks; identity layers
  // (nbr == NULL) a tensor map over the input so the weight warp can bring the A operand by TMA tile loads.
  // LIDAL_LEAN bit 0: gathered layers, bit 1: identity layers (A/B switch, default both).
  static const int lean_env = getenv("LIDAL_LEAN") ? atoi(getenv("LIDAL_LEAN")) : 2;   // measured (profiles/r02_conv_lean_modes.txt): bit 0 is slower than the prologue path
  static const int lean_gate = (getenv("LIDAL_PROD_MODE") && atoi(getenv("LIDAL_PROD_MODE"))) || (getenv("LIDAL_TMA_GATHER") && atoi(getenv("LIDAL_TMA_GATHER"))) ||
                               (getenv("LIDAL_WEIGHT_WARP") && !atoi(getenv("LIDAL_WEIGHT_WARP"))) || getenv("LIDAL_DBG");
  int lean = 0;
  CUtensorMap in_map = map;                      // placeholder unless a TMA producer is used
  if (!pack8 && !lean_gate && !(a.flags & LB_CONV_NO_LEAN) && !a.n_out_dev && a.n_out < ((int64_t)1 << 30) && a.n_in < ((int64_t)1 << 31)) {
    if (a.nbr && a.tile_masks && (lean_env & 1)) lean = 1;
    else if (!a.nbr && (lean_env & 2) && ((uintptr_t)a.in & 15) == 0 && (a.ld_in * 2) % 16 == 0 && a.n_in > 0) {
      cuuint64_t idim[2] = {(cuuint64_t)a.c_in, (cuuint64_t)a.n_in};
      cuuint64_t istr[1] = {(cuuint64_t)a.ld_in * 2};
      cuuint32_t ibox[2] = {(cuuint32_t)bk, (cuuint32_t)TILE_M};
      if (encode(&in_map, a.act_dtype == LB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                 const_cast<void*>(a.in), idim, istr, ibox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        lean = 2;
    }
  }

  TcParams p;
  p.in = (const char*)a.in; p.n_in = a.n_in; p.ld_in = a.ld_in;
  p.out = (char*)a.out; p.n_out = a.n_out; p.ld_out = a.ld_out;
  p.n_out_dev = a.n_out_dev; p.nbr = a.nbr; p.nbr_ld = a.nbr_ld; p.out_rows = a.out_rows;
  p.k_vol = a.k_vol; p.c_in = a.c_in; p.c_out = a.c_out;
  p.scale = a.scale; p.shift = a.shift; p.residual = (const char*)a.residual; p.ld_res = a.ld_res;
  p.out_dtype = a.out_dtype; p.relu = (a.flags & LB_CONV_RELU) ? ((a.flags & LB_CONV_RELU_FIRST) ? 2 : 1) : 0; p.is_bf16 = a.act_dtype == LB_DT_BF16;
  // 256-row CTA tiles (two accumulators sharing every weight tile) when (a) there are enough rows to keep every SM
  // busy (>= 8 waves of 128-row tiles) and (b) the doubled A operand still leaves a >= 4 stage ring.  The kernel is bound
  // by per-stage issue overhead (producer + MMA warps), so halving the stage count per row wins wherever it fits.
  const int64_t tiles128 = (a.n_out + TILE_M - 1) / TILE_M;
  int T = 1;
  {
    const size_t blk2 = (size_t)2 * TILE_M * bk * 2 + (((size_t)a.c_out * bk * 2 + 1023) & ~(size_t)1023);
    const size_t budget2 = 227 * 1024 - 1024 - tail_bytes(2, lean);
    static const int t2_min_stages = getenv("LIDAL_T2_MIN_STAGES") ? atoi(getenv("LIDAL_T2_MIN_STAGES")) : 4;
    static const int t2_min_waves = getenv("LIDAL_T2_MIN_WAVES") ? atoi(getenv("LIDAL_T2_MIN_WAVES")) : 8;
    if (tiles128 >= (int64_t)t2_min_waves * sm_count() && budget2 / blk2 >= (size_t)t2_min_stages && 2 * a.c_out <= 512) T = 2;
  }
  // a single-block 1x1 layer with a wide output is pure epilogue: 128-row tiles keep two accumulator sets and leave
  // room for double-buffered staging
  if (!pack8 && a.k_vol * (a.c_in / bk) == 1 && a.c_out > 128) T = 1;
  // short K loops (1x1 layers) live and die by the staged epilogue: keep 256-row tiles only if the staging buffer still
  // fits next to the ring (1x1 256->128: 0.27 ms with direct stores at T=2 vs 0.15 ms staged at T=1, measured)
  if (T == 2 && !pack8 && a.k_vol * (a.c_in / bk) < 5 && a.out_dtype != LB_DT_F32 && !(a.flags & LB_CONV_NO_STAGED_EPILOGUE)) {
    const size_t blk2 = (size_t)2 * TILE_M * bk * 2 + (((size_t)a.c_out * bk * 2 + 1023) & ~(size_t)1023);
    const size_t budget2 = 227 * 1024 - 1024 - tail_bytes(2, lean);
    const int bpt = a.k_vol * (a.c_in / bk);
    const size_t want = bpt < 3 ? 3 : bpt;
    if (budget2 <= staging_bytes(a.c_out) || (budget2 - staging_bytes(a.c_out)) / blk2 < want) T = 1;
  }
  if (a.flags & LB_CONV_TILE128) T = 1;
  p.T = T;
  p.n_acc = (2 * T * a.c_out <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.n_acc * T * a.c_out) cols <<= 1;
  p.tmem_cols = cols;
  const int row_bytes = bk * 2;
  const size_t block_bytes = (size_t)T * TILE_M * row_bytes + (((size_t)a.c_out * row_bytes + 1023) & ~(size_t)1023);
  const size_t budget = 227 * 1024 - 1024 - tail_bytes(T, lean);
  // blocks per stage: the loop slot free -> gather issue -> data landed -> MMA -> slot free is a chain of latencies, and on
  // the fine levels the per-stage hand-shake of the gather warps is part of it (profiles/r02_conv_knockouts.txt,
  // r02_conv_lean_modes.txt), so stages carry up to LIDAL_NB_MAX blocks wherever >= 3 such stages still fit.
  static const int nb_max = getenv("LIDAL_NB_MAX") ? atoi(getenv("LIDAL_NB_MAX")) : 2;   // measured: 1 -> 4.97, 2 -> 4.92, 3 -> 4.89 ms of conv per step
  static const int prod_mode_env0 = getenv("LIDAL_PROD_MODE") ? atoi(getenv("LIDAL_PROD_MODE")) : 0;
  static const int wwarp_env0 = getenv("LIDAL_WEIGHT_WARP") ? atoi(getenv("LIDAL_WEIGHT_WARP")) : 1;
  int nb = 1;
  static const int tma_gather_env0 = getenv("LIDAL_TMA_GATHER") ? atoi(getenv("LIDAL_TMA_GATHER")) : 0;
  if (!pack8 && prod_mode_env0 == 0 && wwarp_env0 && !tma_gather_env0 && a.k_vol * (a.c_in / bk) > 1) {
    const size_t stg = (a.out_dtype != LB_DT_F32 && !(a.flags & LB_CONV_NO_STAGED_EPILOGUE)) ? staging_bytes(a.c_out) : 0;
    for (int cand = nb_max; cand > 1; --cand)
      if (budget > stg && (budget - stg) / ((size_t)cand * block_bytes) >= 3) { nb = cand; break; }
  }
  const size_t stage_bytes = nb * block_bytes;
  // smem-staged epilogue (coalesced bulk row copies) when the output is 16-bit and the ring keeps enough stages:
  // all blocks of a tile for short K loops (1x1 layers are pure epilogue), at least 5 stages otherwise
  const int blocks_per_tile = pack8 ? (a.k_vol * 8 + bk - 1) / bk : a.k_vol * (a.c_in / bk);
  int want_stages = blocks_per_tile < 5 ? (blocks_per_tile < 3 ? 3 : blocks_per_tile) : 5;
  if (nb > 1) want_stages = (want_stages + nb - 1) / nb < 3 ? 3 : (want_stages + nb - 1) / nb;   // same number of blocks in flight
  p.staged = 0;
  p.stg_bufs = 1;
  size_t budget_eff = budget;
  // permuted rows leave the staged epilogue as ONE bulk copy per row and lane (32 serialised UBLKCP per warp and sub-tile):
  // for narrow outputs that costs more than the handful of direct 16-byte stores it replaces (A/B: LIDAL_STAGED_MIN_COUT)
  static const int staged_min_cout = getenv("LIDAL_STAGED_MIN_COUT") ? atoi(getenv("LIDAL_STAGED_MIN_COUT")) : 64;
  const bool narrow_permuted = a.out_rows && a.c_out < staged_min_cout;
  if (a.out_dtype != LB_DT_F32 && !(a.flags & LB_CONV_NO_STAGED_EPILOGUE) && !narrow_permuted && budget > staging_bytes(a.c_out) &&
      (budget - staging_bytes(a.c_out)) / stage_bytes >= (size_t)want_stages) {
    p.staged = 1;
    // short K loops (1x1 layers) are pure epilogue: a second staging buffer per warp lets the copy engine drain one
    // sub-tile while the next one is converted (single-buffered, the warp idles for the store's read latency)
    static const bool no_dbuf = getenv("LIDAL_NO_STAGE_DBUF") != nullptr;
    // (with two epilogue groups working on alternate sub-tiles / tiles a second buffer per warp buys nothing: LIDAL_STAGE_DBUF=1 for A/B)
    static const bool want_dbuf = getenv("LIDAL_STAGE_DBUF") != nullptr;
    if (want_dbuf && !no_dbuf && blocks_per_tile < 5 && budget > staging_bytes(a.c_out, 2) &&
        (budget - staging_bytes(a.c_out, 2)) / stage_bytes >= (size_t)want_stages)
      p.stg_bufs = 2;
    budget_eff = budget - staging_bytes(a.c_out, p.stg_bufs);
    // natural row order (no out_rows permutation, host-known row count): whole [32 x 32] boxes through tensor maps
    static const bool no_tile = getenv("LIDAL_NO_TILE_EPILOGUE") != nullptr;
    if (!no_tile && !a.out_rows && !a.n_out_dev && a.n_out < ((int64_t)1 << 31)) p.staged = 2;
  }
  CUtensorMap out_map = map, res_map = map;     // placeholders unless staged == 2
  if (p.staged == 2) {
    const CUtensorMapDataType dt = a.act_dtype == LB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t od[2] = {(cuuint64_t)a.c_out, (cuuint64_t)a.n_out};
    cuuint32_t obox[2] = {32, 32};
    cuuint64_t os[1] = {(cuuint64_t)a.ld_out * 2};
    r = encode(&out_map, dt, 2, a.out, od, os, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && a.residual) {
      cuuint64_t rs[1] = {(cuuint64_t)a.ld_res * 2};
      r = encode(&res_map, dt, 2, const_cast<void*>(a.residual), od, rs, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) p.staged = 1;         // e.g. a stride the tensor map cannot express: per-row copies still work
  }
  int stages = (int)(budget_eff / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { set_error("lb_conv_fwd: shared memory too small for the pipeline"); return LB_ECAP; }
  p.stages = stages;
  p.nb = nb;
  p.pack8 = pack8 ? 1 : 0;
  static const int prod_mode_env = getenv("LIDAL_PROD_MODE") ? atoi(getenv("LIDAL_PROD_MODE")) : 0;   // A/B switch
  static const int tma_gather_env = getenv("LIDAL_TMA_GATHER") ? atoi(getenv("LIDAL_TMA_GATHER")) : 0;   // measured slower in situ (profiles/r02_conv_producer_modes.txt)
  static const int dbg_env = getenv("LIDAL_DBG") ? atoi(getenv("LIDAL_DBG")) : 0;
  p.dbg = dbg_env;
  p.prod_mode = pack8 ? 0 : prod_mode_env;
  p.zero_row = 0; p.zero_mask = 0;
  // TMA gather producer: needs a pool of all-zero rows right behind the input (absent neighbours are fetched from there),
  // a neighbour table (k_vol > 1 or permuted rows) and row indices that fit the tensor map's int32 coordinates
  if (!pack8 && tma_gather_env && a.nbr && a.in_pad_rows >= 16 && (a.in_pad_rows & (a.in_pad_rows - 1)) == 0 &&
      a.n_in + a.in_pad_rows < ((int64_t)1 << 31) && ((uintptr_t)a.in & 15) == 0 && (a.ld_in * 2) % 16 == 0 &&
      (tma_gather_env >= 2 || bk == 64) && LIDAL_PROD_WARPS == 8) {
    cuuint64_t idim[2] = {(cuuint64_t)a.c_in, (cuuint64_t)(a.n_in + a.in_pad_rows)};
    cuuint64_t istr[1] = {(cuuint64_t)a.ld_in * 2};
    cuuint32_t ibox[2] = {(cuuint32_t)bk, 1};
    r = encode(&in_map, a.act_dtype == LB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
               const_cast<void*>(a.in), idim, istr, ibox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      p.prod_mode = 3;
      p.zero_row = (int)a.n_in;
      p.zero_mask = (int)a.in_pad_rows - 1;
    }
  }
  static const int wwarp_env = getenv("LIDAL_WEIGHT_WARP") ? atoi(getenv("LIDAL_WEIGHT_WARP")) : 1;   // A/B switch
  p.wwarp = (!pack8 && p.prod_mode == 0 && wwarp_env) ? 1 : 0;
  p.prod_warps = stages < NUM_PROD_THREADS / 32 ? stages : NUM_PROD_THREADS / 32;
  static const bool static_tiles = getenv("LIDAL_STATIC_TILES") != nullptr;   // A/B switch
  p.sched = static_tiles ? nullptr : (unsigned*)a.sched_ws;   // caller-owned, zeroed once, private to this stream
  p.lean = lean;
  static const int fastpro_env = getenv("LIDAL_FASTPRO") ? atoi(getenv("LIDAL_FASTPRO")) : 1;   // A/B switch
  p.fastpro = (fastpro_env && !lean && !pack8 && a.nbr && a.tile_masks && !a.n_out_dev && !(a.flags & LB_CONV_NO_LEAN) &&
               a.n_out < ((int64_t)1 << 30) && (a.nbr_ld & 3) == 0 && ((uintptr_t)a.nbr & 15) == 0) ? 1 : 0;
  p.tile_masks = a.tile_masks;
  p.idx_bytes = (int)idx_bytes(T, lean);
  static const int lean_arrive_env = getenv("LIDAL_LEAN_ARRIVE") ? atoi(getenv("LIDAL_LEAN_ARRIVE")) : 0;
  static const int lean_lag_env = getenv("LIDAL_LEAN_LAG") ? atoi(getenv("LIDAL_LEAN_LAG")) : 3;
  p.lean_arrive = (lean == 1 && lean_arrive_env && stages >= 3) ? 1 : 0;
  p.lean_lag = stages - 2 < lean_lag_env ? stages - 2 : lean_lag_env;
  if (p.lean_lag < 1) p.lean_lag = 1;
  if (p.lean_lag > 3) p.lean_lag = 3;
  if (lean) p.sched = nullptr;                   // static snake schedule: every role derives the tile sequence itself
  const size_t smem = (size_t)stages * stage_bytes + tail_bytes(T, lean) + (p.staged ? staging_bytes(a.c_out, p.stg_bufs) : 0) + 1024;
  int64_t tiles = (a.n_out + (int64_t)T * TILE_M - 1) / ((int64_t)T * TILE_M);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if (grid < 1) grid = 1;
#define LB_TC_LAUNCH(BKV, TV, KV)                                                                                     \
  do {                                                                                                                \
    {                                                                                                                 \
      static bool opted[64] = {};                /* per device: raise the dynamic shared memory limit once */          \
      int dev_ = 0;                                                                                                   \
      cudaGetDevice(&dev_);                                                                                           \
      if (dev_ < 0 || dev_ >= 64 || !opted[dev_]) {                                                                   \
        LB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BKV, TV, KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        if (dev_ >= 0 && dev_ < 64) opted[dev_] = true;                                                               \
      }                                                                                                               \
    }                                                                                                                 \
    conv_tc_kernel<BKV, TV, KV><<<grid, NUM_THREADS, smem, st>>>(map, out_map, res_map, in_map, p);                              \
    LB_LAUNCHED(1);                                                                                                   \
  } while (0)
#define LB_TC_LAUNCH_K(BKV, TV)                                                                                       \
  do {                                                                                                                \
    if (a.k_vol == 1) LB_TC_LAUNCH(BKV, TV, 1);                                                                       \
    else if (a.k_vol <= 8) LB_TC_LAUNCH(BKV, TV, 8);                                                                  \
    else LB_TC_LAUNCH(BKV, TV, MAX_KVOL);                                                                             \
  } while (0)
  if (bk == 64 && T == 1) LB_TC_LAUNCH_K(64, 1);
  else if (bk == 64) LB_TC_LAUNCH_K(64, 2);
  else if (T == 1) LB_TC_LAUNCH_K(32, 1);
  else LB_TC_LAUNCH_K(32, 2);
#undef LB_TC_LAUNCH_K
#undef LB_TC_LAUNCH
  LB_LAUNCH_CHECK();
  if ((DBG(p) & 128) && p.sched) {
    unsigned long long h[22];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.sched + 16, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemset(p.sched + 16, 0, sizeof(h));
    fprintf(stderr, "[conv dbg] k=%d cin=%d cout=%d n=%lld T=%d stages=%d | mma: total %llu wait_full %llu wait_tempty %llu stages %llu | prod0: total %llu wait_empty %llu barA %llu barB %llu tiles %llu | epi0: total %llu wait_tstart %llu wait_tfull %llu | fill latency (issue->landed, MMA blocked) %llu / %llu | drain latency (commit->free, gather blocked) %llu / %llu | mma segments: fence %llu issue %llu commit %llu | gather segments: ldgsts %llu close %llu\n",
            a.k_vol, a.c_in, a.c_out, (long long)a.n_out, T, stages, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12], h[15], h[13], h[14], h[16], h[17], h[18], h[19], h[20]);
  }
  return LB_OK;
}

}  // namespace lb
