"""ctypes binding of libl2d_b200.so (include/l2d_b200.h).

There is no fallback: if the shared library has not been built (`python live2diff_b200/csrc/build.py`
or `__graft_entry__.build()`), importing the ops raises.  Every call checks the return code and raises
RuntimeError(l2d_last_error()), mirroring the reference, which surfaces failures as Python exceptions.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# L2D_LIB_OVERRIDE (developer A/B runs, profiles/ab_build.py): load another build of the library -- its source hash is then
# not compared with the tree, and a line on stderr says so
_OVERRIDE = os.environ.get("L2D_LIB_OVERRIDE")
LIB_PATH = _OVERRIDE or os.path.join(_HERE, "libl2d_b200.so")

vp, i64, i32, f32, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64


class L2DTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", vp), ("ndim", C.c_int32), ("shape", C.c_int64 * 5)]


class L2DUnetConfig(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("block_out_channels", C.c_int32 * 8), ("layers_per_block", C.c_int32),
                ("heads", C.c_int32), ("cross_attention_dim", C.c_int32), ("ctx_len", C.c_int32),
                ("groups", C.c_int32), ("window", C.c_int32), ("n_rows", C.c_int32), ("latent_h", C.c_int32),
                ("latent_w", C.c_int32), ("mapping_channels", C.c_int32 * 8), ("n_mapping", C.c_int32),
                ("down_has_attn", C.c_int32 * 8), ("up_has_attn", C.c_int32 * 8), ("norm_eps", C.c_float),
                ("use_cuda_graph", C.c_int32), ("warmup_frames", C.c_int32)]


class L2DUnetStepArgs(C.Structure):
    _fields_ = [("sample", vp), ("timestep", vp), ("encoder_hidden_states", vp), ("temporal_attention_mask", vp),
                ("depth_sample", vp), ("kv_cache", C.POINTER(vp)), ("n_kv", C.c_int32), ("pe_idx", vp),
                ("update_idx", vp), ("out_sample", vp), ("reuse_constants", C.c_int32)]


# name -> (restype, argtypes); the symbol list is also what tests/test_cabi.py checks against the header
SIGNATURES = {
    "l2d_abi_version": (i32, []),
    "l2d_build_hash": (C.c_char_p, []),
    "l2d_last_error": (C.c_char_p, []),
    "l2d_launch_count": (i64, []),
    "l2d_kv_attn": (i32, [vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "l2d_warmup_attn": (i32, [vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp]),
    "l2d_layernorm": (i32, [vp, vp, vp, vp, i32, i32, f32, vp]),
    "l2d_groupnorm_workspace_bytes": (i64, [i32, i32]),
    "l2d_groupnorm": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, i32, i32, vp]),
    "l2d_im2col3x3": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "l2d_im2col3x3_nchw4": (i32, [vp, vp, i32, i32, i32, vp]),
    "l2d_nchw_to_nhwc": (i32, [vp, vp, i32, i32, i32, vp]),
    "l2d_nhwc_to_nchw": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "l2d_gemm": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, vp, vp, i32, vp, i64, i32, vp]),
    "l2d_conv3x3": (i32, [vp, i32, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, i64, i32, vp]),
    "l2d_gemm_tile_n": (i32, [i32, i32, i32]),
    "l2d_geglu_interleave": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "l2d_small_linear": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "l2d_timestep_embedding": (i32, [vp, vp, i32, i32, vp]),
    "l2d_attention": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp]),
    "l2d_lcm_step": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "l2d_tt_create": (i32, [C.POINTER(vp), C.POINTER(L2DTensor), i32, i32, i32, i32, i32, i32, i32, i32]),
    "l2d_tt_forward": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "l2d_tt_destroy": (None, [vp]),
    "l2d_unet_create": (i32, [C.POINTER(vp), C.POINTER(L2DUnetConfig), C.POINTER(L2DTensor), i32]),
    "l2d_unet_create_shared": (i32, [C.POINTER(vp), C.POINTER(L2DUnetConfig), vp]),
    "l2d_unet_step": (i32, [vp, C.POINTER(L2DUnetStepArgs), vp]),
    "l2d_unet_profile_step": (i32, [vp, C.POINTER(L2DUnetStepArgs), vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "l2d_unet_constants_epoch": (i64, [vp]),
    "l2d_unet_device_bytes": (i64, [vp]),
    "l2d_unet_launches_per_step": (i64, [vp]),
    "l2d_unet_destroy": (None, [vp]),
    "l2d_taesd_create": (i32, [C.POINTER(vp), C.POINTER(L2DTensor), i32, i32, i32, i32]),
    "l2d_taesd_encode": (i32, [vp, vp, vp, i32, vp]),
    "l2d_taesd_decode": (i32, [vp, vp, vp, i32, i32, vp]),
    "l2d_taesd_device_bytes": (i64, [vp]),
    "l2d_taesd_destroy": (None, [vp]),
    "l2d_image_u8_to_f16": (i32, [vp, vp, i32, i32, i32, vp]),
    "l2d_image_f16_to_u8": (i32, [vp, vp, i32, i32, i32, vp]),
    "l2d_stream_create": (i32, [C.POINTER(vp), vp, C.POINTER(i64), C.POINTER(f32), i32, u64, i32, i32]),
    "l2d_stream_destroy": (None, [vp]),
    "l2d_stream_reset": (i32, [vp, vp]),
    "l2d_stream_set_prompt": (i32, [vp, vp, i32, vp]),
    "l2d_stream_set_cache": (i32, [vp, C.POINTER(vp), i32]),
    "l2d_stream_frame": (i32, [vp, vp, vp, vp, vp, vp]),
    "l2d_stream_launches_per_frame": (i64, [vp]),
    "l2d_stream_get_schedule": (i32, [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.POINTER(u64)]),
    "l2d_stream_state_bytes": (i64, [vp]),
    "l2d_stream_save_state": (i32, [vp, vp, i64]),
    "l2d_stream_load_state": (i32, [vp, vp, i64]),
    "l2d_ring_schedule_host": (i32, [C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), i32, i32, i32, i32, i32]),
    "l2d_stream_randn_host": (i32, [u64, u64, C.c_uint32, C.POINTER(f32), i32]),
    "l2d_philox4x32_10_host": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}

# developer hooks (include/l2d_b200_debug.h): used by profiles/*.py only, not part of the drop-in ABI
DEBUG_SIGNATURES = {
    "l2d_kv_attn_set_debug": (None, [vp]),
    "l2d_gemm_set_debug": (None, [vp]),
    "l2d_unet_set_ablation": (None, [vp, i32]),
    "l2d_stream_invalidate_graph": (None, [vp]),
    "l2d_debug_flash_ctas_per_sm": (i32, [i32]),
    "l2d_flash_set_debug": (None, [vp]),
}

ABI_VERSION = 3
_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA library first (python live2diff_b200/csrc/build.py). "
                "live2diff_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in {**SIGNATURES, **DEBUG_SIGNATURES}.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.l2d_abi_version() != ABI_VERSION:
            raise RuntimeError("libl2d_b200.so ABI version mismatch: rebuild (python live2diff_b200/csrc/build.py)")
        # a library compiled from other sources than the ones lying next to it is refused (stale kernels would make
        # every parity / bench number meaningless); installs that ship no sources skip the check
        try:
            from .csrc import build as _build

            want = _build.source_hash()
        except Exception:
            want = None
        got = handle.l2d_build_hash().decode()
        if _OVERRIDE:
            import sys

            print(f"[live2diff_b200] L2D_LIB_OVERRIDE: using {LIB_PATH} (build {got[:12]}); source-hash check skipped", file=sys.stderr)
            want = None
        if want is not None and got != want:
            raise RuntimeError(f"libl2d_b200.so was built from other sources (hash {got[:12]} != tree {want[:12]}): "
                               "rebuild with python live2diff_b200/csrc/build.py")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().l2d_last_error()
        raise RuntimeError(f"libl2d_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def make_tensor_table(named):
    """{name: torch fp16 cuda contiguous tensor} -> (ctypes array of l2d_tensor, keep-alive list)."""
    arr = (L2DTensor * len(named))()
    keep = []
    for i, (name, t) in enumerate(named.items()):
        b = name.encode()
        keep.append(b)
        arr[i].name = b
        arr[i].data = t.data_ptr()
        arr[i].ndim = t.dim()
        if t.dim() > 5:
            raise ValueError(f"{name}: rank > 5")
        for d in range(t.dim()):
            arr[i].shape[d] = t.shape[d]
    return arr, keep
