"""ctypes binding of libsaro_gs_b200.so (the C ABI in include/saro_gs_b200.h).

There is NO CPU or PyTorch fallback: if the shared library is missing or does not export
the full ABI this module raises, loudly, at first use.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsaro_gs_b200.so")

SGS_FLAG_KEEP_FOR_BACKWARD = 1
SGS_FLAG_NO_TILE_CULL = 2

RESIZE_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_i64 = ctypes.c_int64

# symbol -> (restype, argtypes); must list every symbol include/saro_gs_b200.h declares
ABI = {
    "sgs_abi_version": (_i, []),
    "sgs_last_error": (ctypes.c_char_p, []),
    "sgs_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "sgs_forward": (_i64, [RESIZE_FN, _vp, RESIZE_FN, _vp, RESIZE_FN, _vp,   # resize callbacks
                           _i, _i, _i,                                        # P, D, M
                           _vp, _i, _i,                                       # background, width, height
                           _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp,             # means3D .. cov3D_precomp
                           _vp, _vp, _vp, _f, _f, _i,                         # view, proj, campos, tanfov, prefiltered
                           _vp, _vp, _vp, _i, _vp]),                          # out_color, out_depth, radii, flags, stream
    "sgs_backward": (_i, [_i, _i, _i, _i64, _vp, _i, _i,                      # P, D, M, R, bg, W, H
                          _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp,
                          _vp, _vp, _vp,                                      # state buffers
                          _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_debug_export": (_i, [_i, _i, _i, _i64, _vp, _vp, _vp] + [_vp] * 10 + [_vp]),
    "sgs_debug_kept": (_i64, [_vp, _vp]),
    "sgs_debug_set_capacity": (None, [_i64]),
    "sgs_debug_set_binning_mode": (None, [_i]),
    "sgs_last_forward_counts": (_i, [_vp]),
    "sgs_debug_binning_profile": (_i, [_i, _vp]),
    "sgs_loss_workspace_floats": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "sgs_l1_dssim_forward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_l1_dssim_backward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_l1_dssim_loss_forward": (_i, [_i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "sgs_l1_dssim_loss_backward": (_i, [_i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "sgs_deform_packed_bytes": (ctypes.c_size_t, []),
    "sgs_deform_workspace_bytes": (ctypes.c_size_t, [_i]),
    "sgs_deform_pack_mlp": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_deform_eval": (_i64, [_i, _i, _f] + [_vp] * 10 + [_vp, ctypes.c_size_t] + [_vp] * 5 + [_vp]),
    "sgs_deform_image_bytes": (ctypes.c_size_t, []),
    "sgs_deform_pack_general": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_deform_train_forward": (_i, [_i, _i, _f, _vp, _vp, _i, _vp, _vp]),
    "sgs_deform_train_backward": (_i, [_i, _i, _i, _vp, _vp]),
    "sgs_deform_wgrad_max_ctas": (_i, []),
    "sgs_deform_wgrad_partial_floats": (ctypes.c_size_t, []),
    "sgs_deform_wgrad": (_i, [_i, _i, _vp, _vp, _vp]),
    "sgs_deform_planes_bytes": (ctypes.c_size_t, [_i, _i]),
    "sgs_deform_train_epilogue_forward": (_i, [_i, _f, _f] + [_vp] * 15 + [_vp]),
    "sgs_deform_train_epilogue_backward": (_i, [_i, _f, _f] + [_vp] * 16 + [_vp]),
    "sgs_densify_add_view": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_densify_attach": (_i, [_i, _vp, _vp, _vp]),
    "sgs_densify_commit": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_plane_levels": (_i, [_i, _i, _i]),
    "sgs_plane_pyramid_floats": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "sgs_plane_build": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "sgs_plane_sample_forward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "sgs_plane_sample_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "sgs_plane_fold": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "sgs_profile_enable": (None, [_i]),
    "sgs_profile_read": (_i, [_vp, _vp, _vp]),
}

STAGES = ["preprocess_fwd", "depth_sort_scan", "duplicate", "tile_sort", "tile_ranges", "render_fwd", "bwd_zero",
          "render_bwd", "preprocess_bwd"]



class PlaneDesc(ctypes.Structure):
    """sgs_plane_t of include/saro_gs_b200.h"""
    _fields_ = [("pyramid", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int), ("dim_u", ctypes.c_int),
                ("dim_v", ctypes.c_int), ("max_mip_level", ctypes.c_int)]


class MLPJob(ctypes.Structure):
    """sgs_mlp_job_t of include/saro_gs_b200.h"""
    _fields_ = [("packed", ctypes.c_void_p), ("inp", ctypes.c_void_p), ("out", ctypes.c_void_p), ("save_a", ctypes.c_void_p),
                ("save_b", ctypes.c_void_p), ("save_in", ctypes.c_void_p), ("mask_a", ctypes.c_void_p), ("mask_b", ctypes.c_void_p), ("n_io", ctypes.c_int),
                ("zero_time", ctypes.c_int)]


class WgradTask(ctypes.Structure):
    """sgs_wgrad_task_t of include/saro_gs_b200.h"""
    _fields_ = [("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("groups_b", ctypes.c_int), ("dW", ctypes.c_void_p), ("ldw", ctypes.c_int), ("rows", ctypes.c_int), ("cols", ctypes.c_int),
                ("transposed", ctypes.c_int), ("db", ctypes.c_void_p), ("accumulate", ctypes.c_int)]


_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load (once) and type the native library. Raises NativeLibraryMissing if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found. Build it with `python -m saro_gs_b200.build` "
            "(or __graft_entry__.build()). There is no CPU/PyTorch fallback for the rasterizer.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in ABI.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryMissing(f"{LIB_PATH} does not export {name}: stale build?") from e
        fn.restype = res
        fn.argtypes = args
    if lib.sgs_abi_version() != 1:
        raise NativeLibraryMissing(f"ABI version mismatch: library reports {lib.sgs_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def last_error():
    msg = load().sgs_last_error()
    return msg.decode() if msg else ""
