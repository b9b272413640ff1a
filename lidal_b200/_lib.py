"""ctypes binding of the C-ABI library (include/lidal_b200.h).  PyTorch supplies device memory and
streams only; every signature below is plain pointers and sizes.

There is deliberately no fallback: if ``liblidal_b200.so`` cannot be loaded the import of any
compute path raises, and every wrapper refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIDAL_LIB") or os.path.join(HERE, "liblidal_b200.so")   # LIDAL_LIB: A/B kernel variants

LB_DT_BF16, LB_DT_F16, LB_DT_F32 = 0, 1, 2
LB_CONV_RELU, LB_CONV_FORCE_SIMT, LB_CONV_RELU_FIRST, LB_CONV_PACK8, LB_CONV_TILE128, LB_CONV_NO_STAGED = 1, 2, 4, 8, 16, 32
LB_CONV_NO_LEAN = 64
DT_OF = {torch.bfloat16: LB_DT_BF16, torch.float16: LB_DT_F16, torch.float32: LB_DT_F32}

vp, i64, i32, sz, dbl, flt = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_double, C.c_float


class ConvArgs(C.Structure):
    _fields_ = [("inp", vp), ("n_in", i64), ("ld_in", i64), ("out", vp), ("n_out", i64), ("ld_out", i64),
                ("n_out_dev", vp), ("nbr", vp), ("nbr_ld", i64), ("out_rows", vp), ("weight", vp),
                ("k_vol", i32), ("c_in", i32), ("c_out", i32), ("scale", vp), ("shift", vp), ("residual", vp),
                ("ld_res", i64), ("act_dtype", i32), ("out_dtype", i32), ("flags", i32), ("sched_ws", vp), ("in_pad_rows", i64), ("tile_masks", vp)]


class FrameRef(C.Structure):
    _fields_ = [("grid", vp), ("xyz", vp), ("prob", vp), ("n", i64)]


# name -> (restype, argtypes); must list every symbol include/lidal_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "lb_abi_version": (i32, []),
    "lb_last_error": (C.c_char_p, []),
    "lb_launch_count": (C.c_uint64, []),
    "lb_device_info": (i32, [C.POINTER(i32)] * 3),
    "lb_hash": (i32, [vp, i64, vp, vp]),
    "lb_kernel_hash": (i32, [vp, i64, vp, i32, vp, vp]),
    "lb_hashtable_bytes": (sz, [i64]),
    "lb_hashtable_build": (i32, [vp, i64, vp, sz, vp]),
    "lb_hashtable_query": (i32, [vp, sz, vp, i64, vp, vp]),
    "lb_downsample_ws_bytes": (sz, [i64]),
    "lb_downsample": (i32, [vp, i64, C.POINTER(i32), i32, vp, vp, vp, sz, vp]),
    "lb_kmap_query": (i32, [vp, sz, vp, i64, vp, vp, i32, vp, vp]),
    "lb_kmap_query_sym": (i32, [vp, sz, vp, i64, vp, i32, vp, i64, vp]),
    "lb_kmap_compact_ws_bytes": (sz, [i64, i32]),
    "lb_kmap_compact": (i32, [vp, i64, i32, vp, vp, vp, vp, sz, vp]),
    "lb_kmap_sort_ws_bytes": (sz, [i64]),
    "lb_kmap_sort_by_mask": (i32, [vp, i64, i64, i32, vp, vp, vp, sz, vp]),
    "lb_kmap_sort_by_mask_ld": (i32, [vp, i64, i64, i32, vp, vp, i64, vp, sz, vp]),
    "lb_kmap_sort_by_mask_tm": (i32, [vp, i64, i64, i32, vp, vp, i64, vp, vp, sz, vp]),
    "lb_kmap_transpose": (i32, [vp, i64, i64, i32, vp, i64, vp]),
    "lb_kmap_tile_masks": (i32, [vp, i64, i64, i32, vp, vp]),
    "lb_unique_ws_bytes": (sz, [i64]),
    "lb_unique_i64": (i32, [vp, i64, i32, vp, vp, vp, vp, vp, sz, vp]),
    "lb_group_by_key_ws_bytes": (sz, [i64]),
    "lb_group_by_key": (i32, [vp, i64, vp, vp, vp, vp, sz, vp]),
    "lb_downsample_maps_ws_bytes": (sz, [i64]),
    "lb_downsample_maps": (i32, [vp, i64, i32, vp, vp, vp, i64, vp, vp, sz, vp]),
    "lb_level_counts_ws_bytes": (sz, [i64, i32]),
    "lb_level_counts": (i32, [vp, i64, i32, vp, vp, sz, vp]),
    "lb_sort_pairs_ws_bytes": (sz, [i64]),
    "lb_sort_pairs": (i32, [vp, vp, i64, i32, vp, sz, vp]),
    "lb_conv_pack_weight": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "lb_conv_fwd": (i32, [C.POINTER(ConvArgs), vp]),
    "lb_conv_wgrad": (i32, [vp, i64, i64, vp, i64, i64, vp, C.POINTER(i32), i32, i32, i32, i32, vp, vp]),
    "lb_conv_uses_tensor_cores": (i32, [i32, i32, i32, i32]),
    "lb_conv_sched_ws_bytes": (sz, []),
    "lb_cast": (i32, [vp, i32, i64, vp, i32, i64, i64, i64, vp]),
    "lb_absmax_f32": (i32, [vp, i64, i64, i64, vp, vp]),
    "lb_cast_scaled": (i32, [vp, i64, vp, i32, i64, i64, i64, vp, flt, vp, vp, vp]),
    "lb_count": (i32, [vp, i64, vp, i64, vp]),
    "lb_voxelize_fwd": (i32, [vp, vp, vp, i64, i64, i32, vp, vp]),
    "lb_voxelize_bwd": (i32, [vp, vp, vp, i64, i64, i32, vp, vp]),
    "lb_devoxelize_fwd": (i32, [vp, vp, vp, i64, i64, i32, vp, vp]),
    "lb_devoxelize_bwd": (i32, [vp, vp, vp, i64, i64, i32, vp, vp]),
    "lb_ti_weights": (i32, [vp, i64, vp, i64, flt, vp, vp]),
    "lb_gather_rows16": (i32, [vp, vp, i64, vp, vp]),
    "lb_point_cell_query": (i32, [vp, i64, i64, i32, vp, sz, vp, vp]),
    "lb_point_corner_query": (i32, [vp, i64, i64, i32, vp, sz, vp, vp, vp]),
    "lb_voxelize_fwd_ex": (i32, [vp, i32, i64, vp, vp, i64, i64, i32, vp, vp]),
    "lb_devoxelize_fwd_ex": (i32, [vp, i32, i64, vp, vp, i64, i64, i32, vp, i32, i64, vp]),
    "lb_segment_order_ws_bytes": (sz, [i64]),
    "lb_segment_order": (i32, [vp, i64, vp, i64, vp, vp, vp, sz, vp]),
    "lb_voxelize_segments": (i32, [vp, i32, i64, vp, vp, i64, i32, vp, i64, vp]),
    "lb_tta_transform": (i32, [vp, i64, C.POINTER(dbl), dbl, vp, vp, vp]),
    "lb_tta_quantize": (i32, [vp, i64, C.POINTER(dbl), i32, i32, vp, vp, vp, vp]),
    "lb_tta_views_ws_bytes": (sz, [i64, i32]),
    "lb_tta_views": (i32, [vp, i64, i32, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), dbl, dbl, i32, vp, vp, vp, vp, vp, sz, vp]),
    "lb_register_points": (i32, [vp, i64, i64, C.POINTER(dbl), vp, vp]),
    "lb_tta_feat_mean": (i32, [vp, i32, i64, i64, vp, i32, i64, i32, vp, vp]),
    "lb_segment_entropy": (i32, [vp, i64, i32, vp, vp, i32, vp, vp, sz, vp]),
    "lb_redal_point_scores": (i32, [vp, i64, i32, vp, flt, flt, vp, vp]),
    "lb_region_mean_f32": (i32, [vp, vp, vp, i32, vp, vp]),
    "lb_region_feat_mean": (i32, [vp, i64, i32, vp, vp, i32, vp, vp]),
    "lb_tta_softmax_mean_argmax": (i32, [vp, i64, i32, vp, i32, i64, vp, vp, vp]),
    "lb_frame_grid_bytes": (sz, [i64]),
    "lb_frame_grid_ws_bytes": (sz, [i64]),
    "lb_frame_grid_build": (i32, [vp, i64, dbl, vp, sz, vp, sz, vp]),
    "lb_interframe_score_ws_bytes": (sz, [i64, i32]),
    "lb_interframe_score": (i32, [vp, vp, i64, i32, C.POINTER(FrameRef), i32, dbl, dbl, vp, vp, vp, vp, vp, sz, vp]),
    "lb_frame_level_ws_bytes": (sz, []),
    "lb_frame_level_scores": (i32, [vp, i64, i32, vp, vp, sz, vp]),
    "lb_region_reduce": (i32, [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    "lb_argsort_ws_bytes": (sz, [i64]),
    "lb_argsort_f32": (i32, [vp, i64, vp, vp, sz, vp]),
    "lb_region_pairs_ws_bytes": (sz, [i64]),
    "lb_region_pairs": (i32, [vp, i64, flt, vp, vp, vp, sz, vp]),
    "lb_select_walk": (i32, [vp, i64, vp, vp, vp, vp, vp, i64, vp, i64, i64, i32, i32, i32, vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) the in-tree CUDA library; raise loudly if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m lidal_b200.build` (or __graft_entry__.build()). "
                "lidal_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class LidalError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise LidalError(f"lidal_b200 error {rc}: {lib().lb_last_error().decode()}")


def ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def stream_handle() -> int:
    """cudaStream_t of the current device's current stream.  Called once per launch (~240 times per step), so it takes
    torch's C entry points directly: ``torch.cuda.current_stream()`` builds a Stream object through several Python layers
    (~10 us per call, 2 ms of host time per step)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def stream():
    return C.c_void_p(stream_handle())


_SCHED_CELLS: dict = {}


def conv_sched_ws():
    """The tile-scheduler scratch of lb_conv_fwd for the current (device, stream): allocated and zeroed once, then reused
    (the kernel leaves it zeroed; launches on one stream are ordered).  Returns a raw device pointer."""
    dev = _raw_device() if _raw_device is not None else torch.cuda.current_device()
    key = (dev, stream_handle())
    cell = _SCHED_CELLS.get(key)
    if cell is None:
        cell = torch.zeros(lib().lb_conv_sched_ws_bytes() // 4, dtype=torch.int32, device=f"cuda:{dev}")
        torch.cuda.current_stream().synchronize()         # zeroing is ordered on this stream anyway; once per stream
        _SCHED_CELLS[key] = cell
    return cell.data_ptr()


def require_cuda(*tensors):
    """Every wrapper launches on the CURRENT device's current stream: refuse CPU tensors (no fallback) and tensors that
    live on another GPU (the kernel would run on the wrong device -- wrap the call in ``torch.cuda.device(t.device)``)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise LidalError("lidal_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise LidalError(f"tensor on {t.device} but the current CUDA device is cuda:{cur}: lidal_b200 launches on the "
                             "current device's stream (use torch.cuda.set_device / torch.cuda.device)")
