"""Host-side binding layer: the three functions the reference exposes as its pybind module
``_C`` (``$R/ext.cpp:15-19``, implemented in ``$R/rasterize_points.cu``), with identical
names, argument order, return tuples and error behaviour — implemented on top of the C ABI
(include/saro_gs_b200.h) instead of libtorch C++.  PyTorch is used only for device memory
and the current stream.

    rasterize_gaussians            <- RasterizeGaussiansCUDA          $R/rasterize_points.cu:35-115
    rasterize_gaussians_backward   <- RasterizeGaussiansBackwardCUDA  $R/rasterize_points.cu:117-194
    mark_visible                   <- markVisible                     $R/rasterize_points.cu:196-215
"""
import ctypes
import threading

import torch

from . import _lib


def _ptr(t):
    """Device pointer of a tensor, NULL for empty tensors (the reference's convention for
    'argument not provided': empty tensor -> null data pointer, $R/cuda_rasterizer/forward.cu:205,241)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, dev):
    """contiguous float32 view on `dev` (the reference calls .contiguous().data<float>())."""
    if t is None or t.numel() == 0:
        return t
    if t.device != dev:
        t = t.to(dev)
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {t.dtype}")
    return t.contiguous()


# One persistent ctypes callback (creating CFUNCTYPE objects per call is slow and a bound-method
# callback forms a reference cycle that keeps 100s of MB of state buffers alive until Python's
# cyclic GC runs).  `user` carries the buffer slot; per-thread state makes it re-entrant.
_tls = threading.local()


def _resize_dispatch(user, nbytes):
    """Python side of sgs_resize_fn: like resizeFunctional ($R/rasterize_points.cu:27-33), but a
    fresh uint8 CUDA tensor per call (the state must outlive the call for backward)."""
    try:
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=_tls.device)
        _tls.bufs[int(user or 0)] = t
        return t.data_ptr()
    except Exception:  # e.g. out of memory: report allocation failure through the ABI
        return 0


_RESIZE_CB = _lib.RESIZE_FN(_resize_dispatch)
_GEOM, _BINNING, _IMAGE = 1, 2, 3

# Which binning buffers were filled by an inference-mode forward (no packed per-tile records): keyed by the buffer's
# device address, not by a Python attribute on the tensor object — autograd's saved_tensors may hand back a different
# tensor object for the same storage.  A later forward that reuses the address overwrites the entry, so the table
# stays as small as the set of live addresses (bounded anyway).
_inference_only = {}


def _note_binning_buffer(buf, kept_for_backward):
    if len(_inference_only) > 8192:
        _inference_only.clear()
    _inference_only[(buf.device.index, buf.data_ptr())] = not kept_for_backward


def _check(code, what):
    if code < 0:
        raise RuntimeError(f"{what} failed ({code}): {_lib.last_error()}")
    return code


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            f"saro_gs_b200: {name} must be a CUDA tensor — the rasterizer has no CPU path "
            "(and deliberately no fallback).")


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, keep_for_backward=True, _no_tile_cull=False):
    """-> (num_rendered:int, out_color[3,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer, out_depth[1,H,W])

    `keep_for_backward` is the only addition to the reference signature (default keeps the
    reference behaviour: state buffers usable by rasterize_gaussians_backward).
    """
    lib = _lib.load()
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    _require_cuda(means3D, "means3D")
    dev = means3D.device
    P = int(means3D.shape[0])
    H, W = int(image_height), int(image_width)

    if P == 0:
        # $R/rasterize_points.cu:67-69,80: zero-filled outputs, empty state buffers, nothing launched
        out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((0,), dtype=torch.int32, device=dev)
        out_depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
        e = lambda: torch.empty(0, dtype=torch.uint8, device=dev)
        return 0, out_color, radii, e(), e(), e(), out_depth

    means3D = _f32c(means3D, dev)
    colors = _f32c(colors, dev)
    opacity = _f32c(opacity, dev)
    scales = _f32c(scales, dev)
    rotations = _f32c(rotations, dev)
    cov3D_precomp = _f32c(cov3D_precomp, dev)
    sh = _f32c(sh, dev)
    background = _f32c(background, dev)
    viewmatrix = _f32c(viewmatrix, dev)
    projmatrix = _f32c(projmatrix, dev)
    campos = _f32c(campos, dev)

    M = int(sh.shape[1]) if (sh is not None and sh.numel() != 0) else 0

    # every element is written by the kernels: no zero-fill traffic
    out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    out_depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)

    flags = (_lib.SGS_FLAG_KEEP_FOR_BACKWARD if keep_for_backward else 0) | \
            (_lib.SGS_FLAG_NO_TILE_CULL if _no_tile_cull else 0)
    _tls.device = dev
    _tls.bufs = bufs = {}
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rendered = lib.sgs_forward(
            _RESIZE_CB, _GEOM, _RESIZE_CB, _BINNING, _RESIZE_CB, _IMAGE,
            P, int(degree), M,
            _ptr(background), W, H,
            _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(scales), float(scale_modifier),
            _ptr(rotations), _ptr(cov3D_precomp),
            _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
            1 if prefiltered else 0,
            _ptr(out_color), _ptr(out_depth), _ptr(radii), flags, ctypes.c_void_p(stream))
    _tls.bufs = None
    _check(rendered, "sgs_forward")
    # remembered per binning-buffer address so that a backward call on inference-only state fails loudly
    _note_binning_buffer(bufs[_BINNING], bool(keep_for_backward))
    return int(rendered), out_color, radii, bufs[_GEOM], bufs[_BINNING], bufs[_IMAGE], out_depth


# Densification statistics armed for the next backward call (densify.BatchDensifyStats.attach_next_backward).  A plain
# module slot, not a thread-local: autograd runs backward on its own thread, the arming call comes from the caller's.
_pending_densify = None
_pending_lock = threading.Lock()


def arm_densify_sink(stats):
    global _pending_densify
    with _pending_lock:
        _pending_densify = stats


def _take_densify_sink():
    global _pending_densify
    with _pending_lock:
        stats, _pending_densify = _pending_densify, None
    return stats


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh,
                                 degree, campos, geomBuffer, R, binningBuffer, imageBuffer):
    """-> (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6],
           dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4])"""
    lib = _lib.load()
    _require_cuda(means3D, "means3D")
    dev = means3D.device
    P = int(means3D.shape[0])
    H, W = int(dL_dout_color.shape[1]), int(dL_dout_color.shape[2])
    M = int(sh.shape[1]) if (sh is not None and sh.numel() != 0) else 0
    opts = dict(dtype=torch.float32, device=dev)

    if P == 0:
        z = lambda *s: torch.zeros(s, **opts)
        return z(0, 3), z(0, 3), z(0, 1), z(0, 3), z(0, 6), z(0, M, 3), z(0, 3), z(0, 4)
    if _inference_only.get((binningBuffer.device.index, binningBuffer.data_ptr()), False):
        raise RuntimeError("rasterize_gaussians_backward: the forward call ran with keep_for_backward=False "
                           "(inference mode) — its state buffers do not hold the per-tile lists backward needs")

    means3D = _f32c(means3D, dev)
    colors = _f32c(colors, dev)
    scales = _f32c(scales, dev)
    rotations = _f32c(rotations, dev)
    cov3D_precomp = _f32c(cov3D_precomp, dev)
    sh = _f32c(sh, dev)
    background = _f32c(background, dev)
    viewmatrix = _f32c(viewmatrix, dev)
    projmatrix = _f32c(projmatrix, dev)
    campos = _f32c(campos, dev)
    dL_dout_color = _f32c(dL_dout_color, dev)
    radii = radii.contiguous()

    # all outputs are fully written by the fused backward-preprocess kernel
    dL_dmeans3D = torch.empty((P, 3), **opts)
    dL_dmeans2D = torch.empty((P, 3), **opts)
    dL_dcolors = torch.empty((P, 3), **opts)
    dL_dopacity = torch.empty((P, 1), **opts)
    dL_dcov3D = torch.empty((P, 6), **opts)
    dL_dsh = torch.empty((P, M, 3), **opts)
    dL_dscales = torch.empty((P, 3), **opts)
    dL_drotations = torch.empty((P, 4), **opts)

    sink = _take_densify_sink()
    if sink is not None and (sink.P != P or sink.device != dev):
        raise RuntimeError(f"densification statistics armed for {sink.P} Gaussians on {sink.device}, backward runs "
                           f"{P} on {dev}")
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        if sink is not None:    # same thread as sgs_backward below: the C side keeps the sink per thread, one-shot
            _check(lib.sgs_densify_attach(P, sink.grad_sum.data_ptr(), sink.vis_count.data_ptr(),
                                          sink.radii_max.data_ptr()), "sgs_densify_attach")
        rc = lib.sgs_backward(
            P, int(degree), M, int(R), _ptr(background), W, H,
            _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales), float(scale_modifier), _ptr(rotations),
            _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
            _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
            _ptr(dL_dout_color), _ptr(dL_dmeans2D), None, _ptr(dL_dopacity), _ptr(dL_dcolors),
            _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations),
            ctypes.c_void_p(stream))
    _check(rc, "sgs_backward")
    if sink is not None:
        sink.views += 1
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D, viewmatrix, projmatrix):
    """bool[P]: z_view > 0.2 ($R/cuda_rasterizer/auxiliary.h:139-164 via checkFrustum)."""
    lib = _lib.load()
    _require_cuda(means3D, "means3D")
    dev = means3D.device
    P = int(means3D.shape[0])
    present = torch.zeros((P,), dtype=torch.bool, device=dev)
    if P != 0:
        means3D = _f32c(means3D, dev)
        viewmatrix = _f32c(viewmatrix, dev)
        projmatrix = _f32c(projmatrix, dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _check(lib.sgs_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(present),
                                        ctypes.c_void_p(stream)), "sgs_mark_visible")
    return present


def debug_export(P, W, H, R, geomBuffer, binningBuffer, imageBuffer):
    """Parity-test introspection: dict of internal state tensors (see sgs_debug_export)."""
    lib = _lib.load()
    dev = geomBuffer.device
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    u32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    out = dict(
        tiles_touched=torch.zeros(P, **u32), ranges=torch.zeros((tiles, 2), **u32),
        n_contrib=torch.zeros(H * W, **u32), final_T=torch.zeros(H * W, **f32),
        means2D=torch.zeros((P, 2), **f32), conic_opacity=torch.zeros((P, 4), **f32),
        rgbd=torch.zeros((P, 4), **f32), cov3D=torch.zeros((P, 6), **f32),
        tile_count=torch.zeros(tiles, **u32), point_list=torch.zeros(max(int(R), 1), **u32))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(lib.sgs_debug_export(P, W, H, int(R), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                                    _ptr(out["tiles_touched"]), _ptr(out["ranges"]), _ptr(out["n_contrib"]),
                                    _ptr(out["final_T"]), _ptr(out["means2D"]), _ptr(out["conic_opacity"]),
                                    _ptr(out["rgbd"]), _ptr(out["cov3D"]), _ptr(out["tile_count"]),
                                    _ptr(out["point_list"]), ctypes.c_void_p(stream)), "sgs_debug_export")
        kept = _check(lib.sgs_debug_kept(_ptr(binningBuffer), ctypes.c_void_p(stream)), "sgs_debug_kept")
    out["point_list"] = out["point_list"][:int(kept)]
    out["kept"] = int(kept)
    return out
