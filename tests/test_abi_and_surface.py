"""CPU tests: the C-ABI library loads and exports every declared symbol; the Python surface
mirrors the reference ($R/diff_gaussian_rasterization_ch3/__init__.py); no silent CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "saro_gs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgs_[a-z0-9_]+)\s*\(", text)) - {"sgs_resize_fn"})


def test_header_symbols_exported(native_lib):
    syms = declared_symbols()
    assert {"sgs_forward", "sgs_backward", "sgs_mark_visible", "sgs_abi_version", "sgs_last_error"} <= set(syms)
    for s in syms:
        assert hasattr(native_lib, s), f"{s} declared in include/saro_gs_b200.h but not exported"


def test_binding_table_matches_header(native_lib):
    from saro_gs_b200 import _lib
    assert sorted(_lib.ABI) == declared_symbols()
    assert native_lib.sgs_abi_version() == 1


def test_invalid_arguments_return_error_codes(native_lib):
    # negative sizes are rejected before any CUDA call: safe on a CPU-only box
    from saro_gs_b200 import _lib
    null_cb = ctypes.cast(None, _lib.RESIZE_FN)
    rc = native_lib.sgs_forward(null_cb, None, null_cb, None, null_cb, None, -1, 0, 0, None, 16, 16,
                                None, None, None, None, None, 1.0, None, None, None, None, None, 1.0, 1.0, 0,
                                None, None, None, 0, None)
    assert rc == -1
    assert b"bad sizes" in native_lib.sgs_last_error()
    # P == 0 short-circuits like the reference binding ($R/rasterize_points.cu:80)
    rc = native_lib.sgs_forward(null_cb, None, null_cb, None, null_cb, None, 0, 0, 0, None, 16, 16,
                                None, None, None, None, None, 1.0, None, None, None, None, None, 1.0, 1.0, 0,
                                None, None, None, 0, None)
    assert rc == 0


def test_settings_namedtuple_matches_reference_order():
    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered")


def _settings():
    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings
    return GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                         torch.zeros(3), False)


def test_argument_exclusivity_messages():
    from diff_gaussian_rasterization_ch3 import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.zeros(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_bad_means_shape_raises_like_reference(native_lib):
    import saro_gs_b200 as sgs
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        sgs._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), torch.Tensor([]), torch.zeros(4, 1),
                                   torch.zeros(4, 3), torch.zeros(4, 4), 1.0, torch.Tensor([]), torch.eye(4),
                                   torch.eye(4), 1.0, 1.0, 16, 16, torch.zeros(4, 16, 3), 3, torch.zeros(3), False)


def test_no_cpu_fallback(native_lib):
    """CPU tensors must fail loudly, not fall back to some PyTorch path."""
    from diff_gaussian_rasterization_ch3 import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.markVisible(m)


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from saro_gs_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NativeLibraryMissing):
        _lib.load()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under saro_gs_b200/ may reference it."""
    pkg = os.path.join(ROOT, "saro_gs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "splat_oracle" not in text, f
