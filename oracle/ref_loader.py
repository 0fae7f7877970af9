"""TEST INFRASTRUCTURE — loads the compiled, unmodified reference rasterizer (oracle/_ref/_C*.so,
built by oracle/build_ref.py from $R = /root/reference/submodules/gaussian_rasterization_ch3)
and exposes it through the same host layer as the product (saro_gs_b200.rasterizer.make_api),
plus parsers for the reference's opaque state buffers so integer state (tiles_touched, ranges,
n_contrib, point_list) can be compared bit-for-bit.

Buffer layouts restate the carving order of
  GeometryState::fromChunk   $R/cuda_rasterizer/rasterizer_impl.cu:155-171
  ImageState::fromChunk      $R/cuda_rasterizer/rasterizer_impl.cu:173-180
  BinningState::fromChunk    $R/cuda_rasterizer/rasterizer_impl.cu:182-196
(each array aligned to 128 bytes, $R/cuda_rasterizer/rasterizer_impl.h:21-27).
"""
import importlib.util
import os

import torch

from .build_ref import ref_so_path

_mod = None


def available():
    return os.path.exists(ref_so_path())


def load_ref_C():
    """The reference's pybind module `_C` (needs CUDA at call time, not at import time)."""
    global _mod
    if _mod is None:
        path = ref_so_path()
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        spec = importlib.util.spec_from_file_location("_C", path)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def ref_api():
    """(rasterize_gaussians, GaussianRasterizer, Function) driven by the reference `_C`."""
    from saro_gs_b200.rasterizer import make_api
    return make_api(load_ref_C(), supports_keep_flag=False)


def _carve(buf, offset, dtype, count):
    base = buf.data_ptr()
    start = ((base + offset + 127) & ~127) - base
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    view = buf[start:start + nbytes].view(dtype)
    return view, start + nbytes


def parse_ref_geom(buf, P):
    off = 0
    out = {}
    out["depths"], off = _carve(buf, off, torch.float32, P)
    out["clamped"], off = _carve(buf, off, torch.uint8, 3 * P)
    out["internal_radii"], off = _carve(buf, off, torch.int32, P)
    m2d, off = _carve(buf, off, torch.float32, 2 * P)
    out["means2D"] = m2d.view(P, 2)
    c3, off = _carve(buf, off, torch.float32, 6 * P)
    out["cov3D"] = c3.view(P, 6)
    co, off = _carve(buf, off, torch.float32, 4 * P)
    out["conic_opacity"] = co.view(P, 4)
    rgb, off = _carve(buf, off, torch.float32, 3 * P)
    out["rgb"] = rgb.view(P, 3)
    out["tiles_touched"], off = _carve(buf, off, torch.int32, P)
    return out


def parse_ref_img(buf, N):
    off = 0
    out = {}
    out["accum_alpha"], off = _carve(buf, off, torch.float32, N)
    out["n_contrib"], off = _carve(buf, off, torch.int32, N)
    r, off = _carve(buf, off, torch.int32, 2 * N)
    out["ranges"] = r.view(N, 2)
    return out


def parse_ref_binning(buf, R):
    off = 0
    out = {}
    out["point_list"], off = _carve(buf, off, torch.int32, R)
    return out
