/*
 * lidal_b200 -- C ABI of the B200-native hot path of hzykent/LiDAL.
 *
 * This is the drop-in boundary "B1-native" of SURVEY.md section 8(b): what the reference
 * reaches through `torchsparse.backend.*` (torchsparse==1.4.0, docs/requirements.txt:191 --
 * an un-vendored third-party extension) plus the device side of the scoring chain
 * (score/prob_inference.py:100-113, score/sv_level/LiDAL.py:59-103,230-325).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless marked [host].
 *   - the caller owns all memory including workspaces (`*_bytes` helpers size them);
 *     no device allocation and no host synchronisation inside a call (the exceptions are marked "syncs");
 *     host-side state is limited to per-process caches (SM count, the driver's tensor-map encoder entry point,
 *     the one-time shared-memory opt-in of each kernel) and the thread-local error text, so every call is
 *     stream-capturable.
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *   - return value: LB_OK or a negative LB_E* code; `lb_last_error()` has the text.
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns LB_ECUDA.
 *   - coordinates are int32 [N,4] = (x, y, z, batch), batch LAST (dataset/sk_dataset.py:191,209).
 */
#ifndef LIDAL_B200_H
#define LIDAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_OK 0
#define LB_EINVAL (-1) /* bad argument (null pointer, unsupported shape)          */
#define LB_ECAP (-2)   /* capacity / workspace too small                          */
#define LB_ECUDA (-3)  /* CUDA runtime error or no device                         */

#define LB_ABI_VERSION 3

int lb_abi_version(void);
const char* lb_last_error(void); /* [host] thread-local text of the last failure */
/* number of kernels this library has launched in this process (bench.py's gpu_launches claim) */
uint64_t lb_launch_count(void);
int lb_device_info(int* sm_count, int* cc_major, int* cc_minor); /* [host] outs */

/* ------------------------------------------------------------------ hashing / kernel maps
 * Replaces torchsparse.backend.hash_cuda / kernel_hash_cuda / hash_query_cuda, reached from
 * F.sphash / F.sphashquery at network/utils.py:17,19,42,47-48,70-76 and from every kernel-map
 * build inside spnn.Conv3d (network/utils.py:110-114,129-133,147-155). */

/* 64-bit FNV-1a over (x,y,z,b) as uint32, folded to 60 bits.  out int64 [n]. */
int lb_hash(const int32_t* coords, int64_t n, int64_t* out, void* stream);
/* hash of (xyz + offsets[k], b); offsets int32 [k,3]; out int64 [k,n] (offset-major). */
int lb_kernel_hash(const int32_t* coords, int64_t n, const int32_t* offsets, int k, int64_t* out, void* stream);

/* Open-addressing table key(int64 >= 0) -> smallest row index holding that key. */
size_t lb_hashtable_bytes(int64_t n_keys);
int lb_hashtable_build(const int64_t* keys, int64_t n, void* table, size_t table_bytes, void* stream);
/* out int64 [nq]: row index or -1 (F.sphashquery). */
int lb_hashtable_query(const void* table, size_t table_bytes, const int64_t* queries, int64_t nq, int64_t* out,
                       void* stream);

/* F.spdownsample for stride in {1, kernel_size}: xyz <- floor(xyz / s) * s per axis (s = stride*tensor_stride),
 * rows deduplicated and returned sorted by (b,x,y,z).  coords must lie in [0, 2^16) and batch in [0, 2^15).
 * out_coords int32 [<=n,4]; n_out: device int32[1].  ws from lb_downsample_ws_bytes(n).  No host sync. */
size_t lb_downsample_ws_bytes(int64_t n);
int lb_downsample(const int32_t* coords, int64_t n, const int32_t sample_stride[3] /*[host]*/, int batch_bits,
                  int32_t* out_coords, int32_t* n_out, void* ws, size_t ws_bytes, void* stream);

/* Fused kernel_hash + table lookup: nbr[k][o] = row of (out_coords[o].xyz + offsets[k], b) among the table's
 * keys, or -1.  This is the `results` matrix of the reference's map build.  n_out_dev (device int32[1]) may be
 * NULL, in which case all n_out_cap rows are valid; nbr int32 [k, n_out_cap]. */
int lb_kmap_query(const void* table, size_t table_bytes, const int32_t* out_coords, int64_t n_out_cap,
                  const int32_t* n_out_dev, const int32_t* offsets, int k, int32_t* nbr, void* stream);

/* Submanifold (stride-1) special case of lb_kmap_query: the table was built from `coords` itself and the offsets are
 * point-symmetric (offsets[k-1-j] == -offsets[j], k odd -- true for get_kernel_offsets of an odd kernel).  Then
 * nbr[j][o] = r  <=>  nbr[k-1-j][r] = o, so only (k+1)/2 offsets are probed and the rest is mirrored.  Same result as
 * lb_kmap_query, bit for bit.  nbr int32 [k, nbr_ld]. */
int lb_kmap_query_sym(const void* table, size_t table_bytes, const int32_t* coords, int64_t n, const int32_t* offsets, int k,
                      int32_t* nbr, int64_t nbr_ld, void* stream);

/* Compaction to the reference's (nbmaps, nbsizes): rows (in_idx, out_idx) enumerated k-major then o ascending.
 * nbmaps int32 [k*n_out,2] (worst case), nbsizes int32 [k], total device int32[1]. */
size_t lb_kmap_compact_ws_bytes(int64_t n_out, int k);
int lb_kmap_compact(const int32_t* nbr, int64_t n_out, int k, int32_t* nbmaps, int32_t* nbsizes, int32_t* total,
                    void* ws, size_t ws_bytes, void* stream);

/* Row permutation that groups output rows with equal neighbour masks (bit k = offset k present) so that the
 * 128-row tiles of lb_conv_fwd can skip whole offsets: perm int32 [n_out] (stable sort by the mask with its bits
 * re-ordered so the least frequent offset is most significant; only the LB_MASK_KEY_BITS most significant key bits
 * take part, i.e. for k = 27 the three most frequent offsets do not influence the order) and the permuted
 * table nbr_sorted[k][j] = nbr[k][perm[j]].  Use as  args.nbr = nbr_sorted, args.out_rows = perm. */
#define LB_MASK_KEY_BITS 24
#define LB_MASK_CHUNK_SHIFT 0   /* > 0: group rows into chunks of 2^shift consecutive rows before the mask (L2 locality) */
size_t lb_kmap_sort_ws_bytes(int64_t n_out);
int lb_kmap_sort_by_mask(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm, int32_t* nbr_sorted,
                         void* ws, size_t ws_bytes, void* stream);
/* Same, with an explicit row stride for the permuted table (nbr_sorted[k][j] at k * sorted_ld + j).  A stride that is a
 * multiple of 4 (16-byte aligned rows) lets lb_conv_fwd read the table with 128-bit loads. */
int lb_kmap_sort_by_mask_ld(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm, int32_t* nbr_sorted,
                            int64_t sorted_ld, void* ws, size_t ws_bytes, void* stream);

/* Active-offset mask of every group of 128 consecutive rows of a neighbour table (any row order; normally the mask-sorted
 * one): tile_masks uint32 [ceil(n_out / 128)], bit j of entry g = OR over rows o in [128 g, 128 g + 128) of (nbr[j][o] >= 0).
 * Built once per map and passed to every convolution that uses it (lb_conv_args.tile_masks). */
int lb_kmap_tile_masks(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, uint32_t* tile_masks, void* stream);

/* lb_kmap_sort_by_mask_ld that also emits the tile masks of the SORTED table (tile_masks uint32 [ceil(n_out / 128)], may be
 * NULL) from the per-row masks it already holds -- 4 bytes per row instead of a second pass over the table. */
int lb_kmap_sort_by_mask_tm(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* perm, int32_t* nbr_sorted,
                            int64_t sorted_ld, uint32_t* tile_masks, void* ws, size_t ws_bytes, void* stream);

/* Per-offset inverse of a neighbour table (transposed convolution / dgrad roles):
 * nbr int32 [k, nbr_ld] with values in [0, n_in) or -1  ->  nbr_t int32 [k, n_in], nbr_t[k][nbr[k][o]] = o. */
int lb_kmap_transpose(const int32_t* nbr, int64_t nbr_ld, int64_t n_out, int k, int32_t* nbr_t, int64_t n_in,
                      void* stream);

/* torch.unique(keys) / np.unique(return_index, return_inverse) : uniq int64 [<=n] ascending, n_unique device int32[1],
 * inverse int32 [n] (optional) = position of keys[i] in uniq, first_row int32 [<=n] (optional) = first occurrence of
 * each unique key (network/utils.py:18-19, dataset/sk_dataset.py:167).  key_bits: significant low bits of the keys. */
size_t lb_unique_ws_bytes(int64_t n);
int lb_unique_i64(const int64_t* keys, int64_t n, int key_bits, int64_t* uniq, int32_t* n_unique, int32_t* inverse,
                  int32_t* first_row, void* ws, size_t ws_bytes, void* stream);

/* Engine-side map construction where the row ORDER of coarse levels is free (results are order independent):
 * lb_group_by_key: groups equal int64 keys without sorting; groups are numbered by first occurrence (deterministic).
 *   inverse int32 [n] = group of row i; first_row int32 [<=n] (optional) = first row of each group; n_groups device int32[1].
 * lb_downsample_maps: kernel 2 / stride 2 level transition in one pass: parent voxels (coords // 2ts * 2ts, grouped by
 *   first occurrence) plus both maps: nbr_dn int32 [8, ld_dn] (child row of parent g at offset k, the F.conv3d map of
 *   the strided conv) and nbr_up int32 [8, n] (its per-offset inverse for the transposed conv).  Offset order is
 *   get_kernel_offsets(2): k = 4*ox + 2*oy + oz.  Same voxel SET as lb_downsample; coordinates must be >= 0. */
size_t lb_group_by_key_ws_bytes(int64_t n);
int lb_group_by_key(const int64_t* keys, int64_t n, int32_t* inverse, int32_t* first_row, int32_t* n_groups, void* ws,
                    size_t ws_bytes, void* stream);
size_t lb_downsample_maps_ws_bytes(int64_t n);
int lb_downsample_maps(const int32_t* coords, int64_t n, int tensor_stride, int32_t* out_coords, int32_t* n_out,
                       int32_t* nbr_dn, int64_t ld_dn, int32_t* nbr_up, void* ws, size_t ws_bytes, void* stream);

/* Row counts of all coarser levels in one pass: counts[l-1] = number of distinct parents (coords / 2^l * 2^l, same 60-bit
 * hash identity as lb_downsample_maps) for l = 1..levels (<= 16).  coords int32 [n,4] >= 0, any row multiplicity (points
 * or voxels).  counts may point into pinned host memory: a caller learns the size of the whole pyramid in ONE round trip
 * before building it (lb_downsample_maps reports the same numbers level by level). */
size_t lb_level_counts_ws_bytes(int64_t n, int levels);
int lb_level_counts(const int32_t* coords, int64_t n, int levels, int32_t* counts, void* ws, size_t ws_bytes, void* stream);

/* Stable LSD radix sort of (uint64 key, uint32 value) pairs on bits [0, end_bit). */
size_t lb_sort_pairs_ws_bytes(int64_t n);
int lb_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int end_bit, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ sparse convolution
 * Replaces torchsparse.backend.convolution_forward_cuda (gather -> cuBLAS mm -> scatter per offset), reached from
 * every spnn.Conv3d.forward: network/minkunet.py:23-84, network/spvcnn.py:21-81 via network/utils.py:105-172.
 *
 *   out[o, :] = epilogue( sum_k  in[nbr[k][o], :] @ W[k] )        rows with nbr < 0 contribute nothing
 *   epilogue(v) = relu?( v * scale[c] + shift[c] + residual[o, c] )   (each part optional)
 *
 * Output-stationary implicit GEMM: no atomics, deterministic.  Activations are 16-bit (bf16 or fp16, fp32
 * accumulate in TMEM); `ld_*` are row strides in ELEMENTS so tensors may be column slices of wider buffers
 * (torchsparse.cat without a copy).  out_rows (optional int32 [n_out]) redirects row o to out_rows[o].
 */
#define LB_DT_BF16 0
#define LB_DT_F16 1
#define LB_DT_F32 2

#define LB_CONV_RELU 1      /* apply ReLU last                                   */
#define LB_CONV_FORCE_SIMT 2 /* use the CUDA-core kernel even where tcgen05 applies */
#define LB_CONV_PACK8 8      /* tiny-channel layers (the c_in = 4 stem): `in` rows hold 8 channels (zero padded,
                                c_in == 8) and `weight` is [c_out][k_vol*8 -> padded to 64] with col = offset*8 + ch;
                                8 offsets share one 64-wide K block of the tensor-core kernel               */
#define LB_CONV_TILE128 16   /* force 128-row CTA tiles (default: 256-row tiles, two accumulators per weight tile, on large inputs) */
#define LB_CONV_NO_STAGED_EPILOGUE 32 /* A/B switch: per-thread 16-byte epilogue stores instead of smem-staged bulk rows */
#define LB_CONV_RELU_FIRST 4 /* with LB_CONV_RELU: relu(v*scale+shift) + residual (SPVCNN point branch)  */
#define LB_CONV_NO_LEAN 64   /* A/B switch: keep the per-tile prologue even when tile_masks are given / the rows are identity */

typedef struct lb_conv_args {
  const void* in;          /* [n_in, ld_in] act_dtype                                        */
  int64_t n_in, ld_in;
  void* out;               /* [n_out, ld_out] out_dtype                                      */
  int64_t n_out, ld_out;
  const int32_t* n_out_dev; /* optional device int32[1]: live row count (<= n_out)            */
  const int32_t* nbr;      /* [k_vol, nbr_ld] neighbour table; NULL => identity (k_vol == 1) */
  int64_t nbr_ld;
  const int32_t* out_rows; /* optional row redirect                                          */
  const void* weight;      /* packed by lb_conv_pack_weight: [k_vol][c_out][c_in] act_dtype  */
  int k_vol, c_in, c_out;
  const float* scale;      /* optional [c_out]                                               */
  const float* shift;      /* optional [c_out]                                               */
  const void* residual;    /* optional [n_out, ld_res] act_dtype                             */
  int64_t ld_res;
  int act_dtype;           /* LB_DT_BF16 | LB_DT_F16                                         */
  int out_dtype;           /* LB_DT_BF16 | LB_DT_F16 | LB_DT_F32                             */
  int flags;
  void* sched_ws;          /* optional device scratch of lb_conv_sched_ws_bytes() bytes, zeroed ONCE by the caller and
                              private to one stream (launches on it are ordered; the kernel leaves it zeroed): enables
                              the dynamic tile scheduler.  NULL = static round-robin tiles.                          */
  int64_t in_pad_rows;     /* 0, or a power of two >= 16: rows [n_in, n_in + in_pad_rows) of `in` exist and are all ZERO.
                              Enables the TMA gather producer (tile::gather4 cannot skip rows, so missing neighbours are
                              fetched from this pool); without it the cp.async producer is used.                       */
  const uint32_t* tile_masks; /* optional uint32 [ceil(n_out / 128)] from lb_kmap_tile_masks(nbr, nbr_ld, n_out, k_vol): bit k of
                              entry g is set iff some row of rows [128 g, 128 g + 128) has a neighbour at offset k.  With it the
                              kernel skips its per-tile prologue (no index staging, no offset-mask reduction, static tile order);
                              a bit that is missing where a neighbour exists would drop that contribution, so build it from
                              exactly the table passed as `nbr`.  NULL = the prologue computes the mask per tile.        */
} lb_conv_args;

size_t lb_conv_sched_ws_bytes(void);

/* fp32 [k_vol, c_in, c_out] (the `kernel` parameter layout of spnn.Conv3d) -> 16-bit [k_vol, c_out, c_in]. */
int lb_conv_pack_weight(const float* kernel, int k_vol, int c_in, int c_out, int act_dtype, void* packed,
                        void* stream);
int lb_conv_fwd(const lb_conv_args* args /*[host]*/, void* stream);
/* which kernel lb_conv_fwd would run for this shape: 1 = tcgen05 implicit GEMM, 0 = CUDA-core. */
int lb_conv_uses_tensor_cores(int k_vol, int c_in, int c_out, int act_dtype);

/* Weight gradient (training; replaces the wgrad half of torchsparse.backend.convolution_backward_cuda, reached from
 * loss.backward() at train.py:137):  grad_kernel[k][ci][co] = sum over pairs (i, o) of offset k  x[i][ci] * g[o][co].
 * x [n_x, ld_x], g [n_g, ld_g] 16-bit; pairs device int32 [M,2] = (x_row, g_row) grouped by offset (the compacted
 * kernel map; swap the columns for a transposed conv); pair_begin [host] int32 [k_vol+1] prefix offsets;
 * grad_kernel fp32 [k_vol, c_in, c_out] (the `kernel` parameter layout), zeroed inside.  tcgen05 with MN-major gathered
 * operands, split-K over pair tiles, fp32 atomics (summation order not deterministic).  c_in % 8 == 0, c_out % 32 == 0. */
int lb_conv_wgrad(const void* x, int64_t n_x, int64_t ld_x, const void* g, int64_t n_g, int64_t ld_g, const int32_t* pairs,
                  const int32_t* pair_begin, int k_vol, int c_in, int c_out, int act_dtype, float* grad_kernel, void* stream);

/* fp32 <-> 16-bit row-strided casts used at the fp32 torchsparse boundary. */
int lb_cast(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst, int64_t rows,
            int64_t cols, void* stream);

/* Training backward (fp16 operands): amax = max |src| over a [rows, cols] fp32 matrix (device float[1], zeroed inside;
 * NaN / Inf ignored), then dst = fp16/bf16(src * s) with s = 2^floor(log2(target / amax)) computed on device.
 * scale_out (optional device float[2]) receives {s, 1/s}; inv_vec (optional device float[256]) is filled with 1/s --
 * pass it as lb_conv_args.scale so the dgrad epilogue removes the scale again (exact: powers of two). */
int lb_absmax_f32(const float* src, int64_t ld, int64_t rows, int64_t cols, float* amax, void* stream);
int lb_cast_scaled(const float* src, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst, int64_t rows, int64_t cols,
                   const float* amax, float target, float* scale_out, float* inv_vec, void* stream);

/* ------------------------------------------------------------------ point <-> voxel (SPVCNN)
 * Replaces torchsparse.backend.count_cuda / voxelize_* / devoxelize_* reached from network/utils.py:20-25,49,56,77-95. */
int lb_count(const int32_t* idx, int64_t n, int32_t* counts, int64_t m, void* stream);
/* out[idx[i]] += feats[i] / counts[idx[i]]  (out zeroed inside).  feats f32 [n,c], out f32 [m,c]. */
int lb_voxelize_fwd(const float* feats, const int32_t* idx, const int32_t* counts, int64_t n, int64_t m, int c,
                    float* out, void* stream);
int lb_voxelize_bwd(const float* grad_out, const int32_t* idx, const int32_t* counts, int64_t n, int64_t m, int c,
                    float* grad_feats, void* stream);
/* out[i] = sum_k w[i,k] * feats[idx[i,k]]; idx int32 [n,8], w f32 [n,8]. */
int lb_devoxelize_fwd(const float* feats, const int32_t* idx, const float* w, int64_t n, int64_t m, int c, float* out,
                      void* stream);
int lb_devoxelize_bwd(const float* grad_out, const int32_t* idx, const float* w, int64_t n, int64_t m, int c,
                      float* grad_feats, void* stream);
/* F.calc_ti_weights: coords f32 [n, ld_c>=3], idx int64 [8,n] (corner-major), out f32 [8,n]. */
int lb_ti_weights(const float* coords, int64_t ld_c, const int64_t* idx, int64_t n, float scale, float* out,
                  void* stream);

/* Fused forms used by the inference engine (same arithmetic as the separate calls above):
 * lb_point_cell_query   network/utils.py:42-48 : idx[i] = voxel holding floor(p/stride)*stride, -1 if none.
 * lb_point_corner_query network/utils.py:69-79 : the 8 surrounding voxels (x slowest, z fastest) and their
 *                       F.calc_ti_weights, idx int32 [n,8], w f32 [n,8].   pts f32 [n, ld>=4] = (x,y,z,...,batch LAST).
 * table: lb_hashtable_build over lb_hash(voxel coords).
 * *_ex: 16-bit or fp32 features with a row stride (elements); voxelize accumulates in fp32. */
/* out[i] = src[idx[i]] for 16-byte rows (int32 [.,4] coordinates). */
int lb_gather_rows16(const void* src, const int32_t* idx, int64_t n, void* out, void* stream);
int lb_point_cell_query(const float* pts, int64_t ld, int64_t n, int stride, const void* table, size_t table_bytes,
                        int32_t* idx, void* stream);
int lb_point_corner_query(const float* pts, int64_t ld, int64_t n, int stride, const void* table, size_t table_bytes,
                          int32_t* idx, float* w, void* stream);
int lb_voxelize_fwd_ex(const void* feats, int feats_dtype, int64_t ld_feats, const int32_t* idx, const int32_t* counts,
                       int64_t n, int64_t m, int c, float* out, void* stream);
int lb_devoxelize_fwd_ex(const void* feats, int feats_dtype, int64_t ld_feats, const int32_t* idx, const float* w,
                         int64_t n, int64_t m, int c, void* out, int out_dtype, int64_t ld_out, void* stream);

/* Atomic-free F.spvoxelize for the engine.  lb_segment_order: from the point -> voxel index (idx int32 [n], -1 = none) and
 * its histogram (counts int32 [m], lb_count) build seg_ptr int32 [m+1] (exclusive scan) and order int32 [seg_ptr[m]] = the
 * point ids of every voxel, contiguous per voxel (order inside a voxel unspecified) -- coordinates only, reusable by every
 * voxelize at that level.  lb_voxelize_segments: out[v] = mean of feats[order[seg_ptr[v] .. seg_ptr[v+1])] in fp32,
 * written as 16-bit [m, ld_out]; feats 16-bit [n, ld_feats]; c % 8 == 0. */
size_t lb_segment_order_ws_bytes(int64_t m);
int lb_segment_order(const int32_t* idx, int64_t n, const int32_t* counts, int64_t m, int32_t* seg_ptr, int32_t* order, void* ws,
                     size_t ws_bytes, void* stream);
int lb_voxelize_segments(const void* feats, int dtype, int64_t ld_feats, const int32_t* order, const int32_t* seg_ptr, int64_t m, int c,
                         void* out, int64_t ld_out, void* stream);

/* Score-mode voxelizer (dataset/sk_dataset.py:143-169), two steps around the caller's
 * data-dependent random shift:  lb_tta_transform: coords_f64 = (raw[:, :3] @ trans_m) * scale (float64 [n,3]) and
 * feats f32 [n,4] = (transformed xyz as f32, intensity);  lb_tta_quantize: (coords_f64 + offset).astype(int) ->
 * coords int32 [n,4] = (x, y, z, batch) and keys int64 [n] = x << 2b | y << b | z (b = coord_bits), whose ascending
 * order is np.unique(axis=0)'s row order; err_flag (device int32[1]) is set if a coordinate leaves [0, 2^b). */
int lb_tta_transform(const float* raw, int64_t n, const double* trans_m, double scale, double* coords_f64, float* feats,
                     void* stream);
int lb_tta_quantize(const double* coords_f64, int64_t n, const double* offset, int batch, int coord_bits, int32_t* coords,
                    int64_t* keys, int32_t* err_flag, void* stream);

/* All `reps` TTA views of ONE scan at once (the batch score/prob_inference.py:91-97 consumes): lb_tta_transform +
 * the min/max of dataset/sk_dataset.py:154-155 + the random shift of :156 + lb_tta_quantize for every view, two launches,
 * no host round trip.  The caller draws the random numbers in the reference's order (per view: randn(3,3), randint,
 * rand, rand(3), rand(3)) and passes [host] trans_m f64 [reps,9] (after the flip / yaw product), r1, r2 f64 [reps,3].
 * Outputs are view-major: feats_p f32 [reps*n,4], coords_p int32 [reps*n,4] (batch column = view), keys int64 [reps*n]
 * = view << 3b | x << 2b | y << b | z  (b = coord_bits): ONE ascending unique over all keys (lb_unique_i64 with
 * key_bits = 3b + 4) yields the collated voxel order, first-point rows and the already offset inverse indices of
 * dataset/sk_dataset.py:188-242.  err_flag: device int32[1], set if a coordinate leaves [0, full_scale).  reps <= 16. */
size_t lb_tta_views_ws_bytes(int64_t n, int reps);
int lb_tta_views(const float* raw, int64_t n, int reps, const double* trans_m, const double* r1, const double* r2,
                 double scale, double full_scale, int coord_bits, float* feats_p, int32_t* coords_p, int64_t* keys,
                 int32_t* err_flag, void* ws, size_t ws_bytes, void* stream);

/* Pose registration (dataset/prepare_kdtree_sk.py:76-80):
 * xyz[i] = (hstack(raw[i,:3], 1) as float32 -> float64 products with pose.T, summed in column order)[:3].
 * raw f32 [n, ld_raw >= 3]; pose [host] f64 [16] row-major 4x4 sensor->world; xyz f64 [n,3].  Bit-equal to numpy. */
int lb_register_points(const float* raw, int64_t ld_raw, int64_t n, const double* pose, double* xyz, void* stream);

/* ------------------------------------------------------------------ prob_inference tail
 * score/prob_inference.py:100-113: gather logits by inverse index, softmax, mean over views, argmax.
 * logits f32 [n_vox, n_cls]; inverse int64 [reps*n_pts]; prob f32 [n_pts, n_cls]; pred int64 [n_pts]. */
int lb_tta_softmax_mean_argmax(const float* logits, int64_t n_vox, int n_cls, const int64_t* inverse, int reps,
                               int64_t n_pts, float* prob, int64_t* pred, void* stream);

/* The out_feat branch of the same tail (score/prob_inference.py:103-105,116-118): out[p] = mean over views of
 * feat[inverse[v*n_pts + p]] in float32, views added in order.  feat [n_vox, ld] of feat_dtype (LB_DT_*), c channels. */
int lb_tta_feat_mean(const void* feat, int feat_dtype, int64_t ld, int64_t n_vox, const int64_t* inverse, int reps,
                     int64_t n_pts, int c, float* out, void* stream);

/* ------------------------------------------------------------------ inter-frame scoring
 * score/sv_level/LiDAL.py:59-103.  One uniform grid per frame replaces the pickled sklearn KD-tree; the match rule
 * is unchanged: exact nearest neighbour in float64, accepted iff sqrt(d2) <= dis_thresh. */
size_t lb_frame_grid_bytes(int64_t n_pts);
/* xyz f64 [n,3] registered coordinates (dataset/prepare_kdtree_sk.py:77-80); `cell` must exceed dis_thresh. */
size_t lb_frame_grid_ws_bytes(int64_t n_pts);
int lb_frame_grid_build(const double* xyz, int64_t n, double cell, void* grid, size_t grid_bytes, void* ws,
                        size_t ws_bytes, void* stream);

typedef struct lb_frame_ref { /* one neighbouring frame, all device pointers */
  const void* grid;    /* built by lb_frame_grid_build over `xyz` (holds a cell-sorted copy of the coordinates) */
  const double* xyz;   /* [n,3] original order (kept for callers; the kernel reads the grid's copy)              */
  const float* prob;   /* [n,n_cls] mean softmax (prob_inference)   */
  int64_t n;
} lb_frame_ref;

/* worker_func's point loop (LiDAL.py:59-81) for ONE query frame against its neighbour window, fused:
 * for each neighbour frame in the given order (nei_ids order): exact 1-NN, match iff dist <= dis_thresh,
 * sum_prob += P_n[nn]; interd += sum_c kl_div(P_q+1e-5, P_n[nn]+1e-5); count += 1; then
 * intere = entropy(sum_prob / count) and interd /= (count-1) where > 0.
 * q_grid: the QUERY frame's own grid (its points are walked in cell-sorted order); q_prob in original row order.
 * nbrs: [host] array of n_nbr (<= 32) frames.  Outputs: interd f64 [nq], intere f32 [nq]; optional count int32 [nq]
 * (= matches) and nn int32 [nq, n_nbr] (matched neighbour row or -1 per neighbour frame).  ws: scratch of
 * lb_interframe_score_ws_bytes(nq, n_nbr) bytes -- phase 1 (one thread per (point, frame) pair: float32 screening on
 * cell-relative coordinates, survivors in exact float64) stores its result there frame-major, phase 2 (one warp per
 * point, lane = class) consumes it. */
size_t lb_interframe_score_ws_bytes(int64_t nq, int n_nbr);
int lb_interframe_score(const void* q_grid, const float* q_prob, int64_t nq, int n_cls, const lb_frame_ref* nbrs,
                        int n_nbr, double dis_thresh, double cell, double* interd, float* intere, int32_t* count,
                        int32_t* nn, void* ws, size_t ws_bytes, void* stream);

/* LiDAL.py:87-98: per-region means over the ragged `sv2point` lists given in CSR form:
 * region_ptr int32 [r+1], region_pts int32 [region_ptr[r]].  Outs: d,e f32 [r]; optional pnums i64 [r],
 * centers f32 [r,3].  One block per region; the two score means follow numpy's pairwise summation order (float64 for
 * interd, float32 for intere) so they equal the reference's `.mean()` bit for bit; deterministic. */
int lb_region_reduce(const double* interd, const float* intere, const double* xyz, const int32_t* region_ptr,
                     const int32_t* region_pts, int n_regions, float* sv_interds, float* sv_interes,
                     int64_t* sv_pnums, float* sv_centers, void* stream);

/* Ascending stable argsort of f32 keys (device radix sort; LiDAL.py:235,283); order int32 [n]. */
size_t lb_argsort_ws_bytes(int64_t n);
int lb_argsort_f32(const float* keys, int64_t n, int32_t* order, void* ws, size_t ws_bytes, void* stream);

/* Regions whose centres lie within `radius` of each other, in the float32 arithmetic of LiDAL.py:252.
 * Pass 1 (nbr_idx == NULL): row[i] = number of in-range regions of i.  Pass 2: row = exclusive offsets
 * (caller scans), nbr_idx receives the lists (order within a row unspecified). */
size_t lb_region_pairs_ws_bytes(int64_t n);
int lb_region_pairs(const float* centers, int64_t n, float radius, int32_t* row, int32_t* nbr_idx, void* ws,
                    size_t ws_bytes, void* stream);

/* HOST function (no device work): the greedy walk of LiDAL.py:242-270 (labelled pass) / :293-325 (pseudo-label pass)
 * over the regions in `visit` order (int64 region ids, already sorted by the caller: descending divergence for the first
 * pass, ascending for the second).  interds / interes double [n_regions] (the float32 region means widened exactly),
 * pnums int64 [n_regions]; row_ptr int64 [n_regions+1] + nbr_idx int32: the in-range lists of lb_region_pairs (host
 * copies); flags int64 [n_regions] in/out.  A candidate with an accepted region within range swaps with it when its
 * entropy is better (prefer_higher_entropy: LiDAL.py:254, otherwise :306); when several accepted regions are in range the
 * reference takes the first one in the iteration order of a CPython set -- replayed slot-exactly (set_last_dummy selects
 * the dummy-reuse rule of the interpreter version; the Python host verifies the model against its own `set` first).
 * point_limit = round(0.01 * train_point_num) (LiDAL.py:240); skip_zero: LiDAL.py:296.  n_added_out (optional) = regions
 * accepted at the end. */
int lb_select_walk(const int64_t* visit, int64_t n_visit, const double* interds, const double* interes,
                   const int64_t* pnums, const int64_t* row_ptr, const int32_t* nbr_idx, int64_t n_regions,
                   int64_t* flags, int64_t flag_value, int64_t point_limit, int prefer_higher_entropy, int skip_zero,
                   int set_last_dummy, int64_t* n_added_out);

/* Frame-level baselines on the same prob maps (score/frame_level/softmax_entropy.py:34,
 * margin_sampling.py:33-34, least_confidence_sampling.py): out3 (device double[3]) = mean entropy(prob), mean
 * (top1 - top2), mean top1 over the n points of one frame. */
size_t lb_frame_level_ws_bytes(void);
int lb_frame_level_scores(const float* prob, int64_t n, int n_cls, double* out3, void* ws, size_t ws_bytes, void* stream);

/* score/frame_level/segment_entropy.py:41-48: frame_sege = sum over regions of (sum_c -q_c log2(q_c + 1e-12)) * len / n,
 * q_c = share of class c among the region's predictions (float64, class order, region order).  pred int64 [n];
 * regions in CSR form; out: device double[1]; ws >= 8 * n_regions bytes.  n_cls <= 64. */
int lb_segment_entropy(const int64_t* pred, int64_t n, int n_cls, const int32_t* region_ptr, const int32_t* region_pts,
                       int n_regions, double* out, void* ws, size_t ws_bytes, void* stream);

/* ReDAL region scoring (score/sv_level/ReDAL.py:63-79):
 * lb_redal_point_scores: point_score = alpha * mean_c(-p log2(p + 1e-12)) + gamma * curvature, float32 throughout
 *   (numpy pairwise order over classes); curvature f32 [n] may be NULL (treated as 0).
 * lb_region_mean_f32:   out[r] = vals[region r].mean() in numpy's float32 pairwise order (sv_scores, ReDAL.py:78).
 * lb_region_feat_mean:  out[r,:] = feat[region r, :].mean(0), float32, rows added in order (sv_feats, ReDAL.py:79);
 *   feat f32 [n, ld], c channels (the 96-dim out_feat of lb_tta_feat_mean). */
int lb_redal_point_scores(const float* prob, int64_t n, int n_cls, const float* curvature, float alpha, float gamma,
                          float* point_score, void* stream);
int lb_region_mean_f32(const float* vals, const int32_t* region_ptr, const int32_t* region_pts, int n_regions, float* out,
                       void* stream);
int lb_region_feat_mean(const float* feat, int64_t ld, int c, const int32_t* region_ptr, const int32_t* region_pts,
                        int n_regions, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDAL_B200_H */
