"""ctypes binding of ``libsynt_isic_b200.so`` (the C ABI declared in include/synt_isic.h).

There is NO fallback: if the library is missing or a call fails, a ``RuntimeError`` is
raised.  Nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsynt_isic_b200.so"

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int)
c_i64p = C.POINTER(C.c_longlong)
c_u8p = C.POINTER(C.c_ubyte)
vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/synt_isic.h
SIGNATURES = {
    "synt_version": (C.c_char_p, []),
    "synt_last_error": (C.c_char_p, []),
    "synt_unet_num_params": (C.c_int, []),
    "synt_unet_param_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, c_i64p, c_i64p]),
    "synt_unet_total_param_count": (C.c_longlong, []),
    "synt_unet_create": (C.c_int, [vp, C.c_longlong, C.c_int, C.POINTER(vp)]),
    "synt_unet_destroy": (C.c_int, [vp]),
    "synt_unet_forward": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "synt_unet_debug_forward": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_char_p, vp, C.c_longlong, c_i32p, c_i32p, c_i32p, vp]),
    "synt_unet_set_schedule": (C.c_int, [vp, C.c_int, c_i32p, c_f32p]),
    "synt_unet_sample": (C.c_int, [vp, vp, C.c_int, vp, C.c_ulonglong, C.c_longlong, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "synt_unet_generate_host": (C.c_int, [vp, vp, C.c_int, C.c_ulonglong, C.c_longlong, C.c_int, vp, vp]),
    "synt_unet_profile_step": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), c_i32p, vp]),
    "synt_unet_profile_records": (C.c_int, [vp, C.POINTER(C.c_double), C.c_int]),
    "synt_unet_workspace_bytes": (C.c_longlong, [vp]),
    "synt_unet_launch_count": (C.c_longlong, [vp]),
    "synt_ddpm_tables": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, c_i64p, c_f32p, c_f32p]),
    "synt_ddpm_step": (C.c_int, [vp, vp, vp, vp, C.c_longlong, c_f32p, vp]),
    "synt_to_uint8": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "synt_resnet18_num_params": (C.c_int, []),
    "synt_resnet18_param_info": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_int, c_i64p, c_i64p]),
    "synt_resnet18_create": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_int, C.POINTER(vp)]),
    "synt_resnet18_destroy": (C.c_int, [vp]),
    "synt_resnet18_logits": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "synt_resnet18_logits_host": (C.c_int, [vp, vp, C.c_int, vp]),
    "synt_resnet18_debug": (C.c_int, [vp, vp, C.c_int, C.c_char_p, vp, C.c_longlong, c_i32p, c_i32p, c_i32p, vp]),
    "synt_resnet18_launch_count": (C.c_longlong, [vp]),
    "synt_resnet18_profile": (C.c_int, [vp, vp, C.c_int, vp, C.POINTER(C.c_double), vp]),
    "synt_resnet18_score_grad": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "synt_resnet18_grad_debug": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_char_p, vp, C.c_longlong, c_i32p, c_i32p, c_i32p, vp]),
    "synt_select_regions": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                      vp, vp, vp]),
    "synt_ig_interpolate": (C.c_int, [vp, vp, C.c_int, C.c_longlong, vp, vp]),
    "synt_ig_reduce": (C.c_int, [vp, vp, vp, C.c_int, C.c_longlong, vp, vp]),
    "synt_intervene_blend": (C.c_int, [vp, vp, vp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "synt_intervene_blend_ex": (C.c_int, [vp, vp, vp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "synt_stat_bootstrap_mean_diff": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, C.c_ulonglong, C.c_int, vp, vp]),
    "synt_stat_permutation_mean_diff": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_ulonglong, C.c_int, vp, vp]),
    "synt_debug_conv": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_int, vp]),
    "synt_debug_conv_gn": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int,
                                     vp, vp, vp, vp, C.c_int, vp, c_i32p, vp]),
    "synt_debug_conv_up2x": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, c_i32p, vp]),
    "synt_unet_set_step_mask": (C.c_int, [vp, vp, C.c_int]),
    "synt_unet_set_image_keys": (C.c_int, [vp, vp]),
    "synt_debug_set_conv_pair": (C.c_int, [C.c_int]),
    "synt_debug_pack_upsample_phases": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "synt_debug_attention": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "synt_patch_mask_apply": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the library once; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m synt_isic_b200.build` "
                "(nvcc, sm_100a). synt_isic_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if a declared symbol is absent
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().synt_last_error()
        raise RuntimeError(f"synt_isic_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def unet_manifest():
    """[(name, numel, offset)] in the library's packing order (diffusers key names)."""
    L = lib()
    out = []
    buf = C.create_string_buffer(256)
    numel, off = C.c_longlong(), C.c_longlong()
    for i in range(L.synt_unet_num_params()):
        check(L.synt_unet_param_info(i, buf, 256, C.byref(numel), C.byref(off)), "unet_param_info")
        out.append((buf.value.decode(), numel.value, off.value))
    return out


def resnet18_manifest(num_classes: int = 7):
    L = lib()
    out = []
    buf = C.create_string_buffer(256)
    numel, off = C.c_longlong(), C.c_longlong()
    for i in range(L.synt_resnet18_num_params()):
        check(L.synt_resnet18_param_info(num_classes, i, buf, 256, C.byref(numel), C.byref(off)), "resnet18_param_info")
        out.append((buf.value.decode(), numel.value, off.value))
    return out
