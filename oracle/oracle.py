"""TEST INFRASTRUCTURE — Python (ctypes + numpy) wrapper of the C oracle (splat_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import
this module.  The product path (saro_gs_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_libs = {}


def build(force=False):
    """Compile the oracle with gcc (both precisions). Falls back to a serial build if OpenMP is unavailable."""
    os.makedirs(_BUILD, exist_ok=True)
    src = os.path.join(_HERE, "splat_oracle.c")
    for prec, real in (("f64", "double"), ("f32", "float")):
        out = os.path.join(_BUILD, f"liboracle_{prec}.so")
        if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(src):
            continue
        base = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", f"-DORACLE_REAL={real}",
                "-o", out, src, "-lm"]
        ok = False
        for cc in ("/usr/bin/gcc", "gcc", "cc"):
            for extra in (["-fopenmp"], []):
                r = subprocess.run([cc] + extra + base, capture_output=True, text=True)
                if r.returncode == 0:
                    ok = True
                    break
            if ok:
                break
        if not ok:
            raise RuntimeError(f"could not build the CPU oracle: {r.stderr}")


def _load(prec):
    if prec in _libs:
        return _libs[prec]
    path = os.path.join(_BUILD, f"liboracle_{prec}.so")
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.oracle_forward.restype = vp
    lib.oracle_forward.argtypes = [i, i, i, vp, i, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f]
    lib.oracle_backward.restype = None
    lib.oracle_backward.argtypes = [vp] * 14
    lib.oracle_free.restype = None
    lib.oracle_free.argtypes = [vp]
    lib.oracle_num_rendered.restype = ctypes.c_int64
    lib.oracle_num_rendered.argtypes = [vp]
    for name in ("radii", "tiles_touched", "point_list", "ranges", "n_contrib", "final_T", "color", "depth_img",
                 "means2D", "conic_opacity", "rgb", "cov3D"):
        fn = getattr(lib, "oracle_" + name)
        fn.restype = vp
        fn.argtypes = [vp]
    lib.oracle_real_bytes.restype = i
    _libs[prec] = lib
    return lib


def _f32(a):
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if a.size else None


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _view(ptr, dtype, shape):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class OracleResult:
    """Forward outputs + internal integer state; `.backward(dL_dcolor)` gives the gradients."""

    def __init__(self, lib, ctx, real, keep, dims):
        self._lib, self._ctx, self._real, self._keep = lib, ctx, real, keep
        P, W, H, M = dims
        self.P, self.W, self.H, self.M = P, W, H, M
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        self.num_rendered = int(lib.oracle_num_rendered(ctx))
        R = self.num_rendered
        self.radii = _view(lib.oracle_radii(ctx), np.int32, (P,))
        self.tiles_touched = _view(lib.oracle_tiles_touched(ctx), np.uint32, (P,))
        self.point_list = _view(lib.oracle_point_list(ctx), np.uint32, (R,))
        self.ranges = _view(lib.oracle_ranges(ctx), np.uint32, (tiles, 2))
        self.n_contrib = _view(lib.oracle_n_contrib(ctx), np.uint32, (H, W))
        self.final_T = _view(lib.oracle_final_T(ctx), real, (H, W))
        self.color = _view(lib.oracle_color(ctx), real, (3, H, W))
        self.depth = _view(lib.oracle_depth_img(ctx), real, (1, H, W))
        self.means2D = _view(lib.oracle_means2D(ctx), real, (P, 2))
        self.conic_opacity = _view(lib.oracle_conic_opacity(ctx), real, (P, 4))
        self.rgb = _view(lib.oracle_rgb(ctx), real, (P, 3))
        self.cov3D = _view(lib.oracle_cov3D(ctx), real, (P, 6))

    def backward(self, dL_dcolor):
        lib, P, M, real = self._lib, self.P, self.M, self._real
        g = _f32(dL_dcolor)
        out = {
            "means2D": np.zeros((P, 3), real), "colors": np.zeros((P, 3), real), "opacities": np.zeros((P, 1), real),
            "means3D": np.zeros((P, 3), real), "cov3D": np.zeros((P, 6), real), "shs": np.zeros((P, M, 3), real),
            "scales": np.zeros((P, 3), real), "rotations": np.zeros((P, 4), real)}
        k = self._keep
        lib.oracle_backward(self._ctx, _p(g), _p(k["means3D"]), _p(k["shs"]), _p(k["scales"]), _p(k["rotations"]),
                            _p(out["means2D"]), _p(out["colors"]), _p(out["opacities"]), _p(out["means3D"]),
                            _p(out["cov3D"]), _p(out["shs"]), _p(out["scales"]), _p(out["rotations"]))
        return out

    def __del__(self):
        try:
            self._lib.oracle_free(self._ctx)
        except Exception:
            pass


def forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, width, height, tanfovx, tanfovy, sh_degree=0,
            shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
            precision="f64"):
    """CPU oracle forward.  Arguments follow GaussianRasterizer.forward + GaussianRasterizationSettings."""
    lib = _load(precision)
    real = np.float64 if precision == "f64" else np.float32
    keep = dict(means3D=_f32(means3D), shs=_f32(shs), colors=_f32(colors_precomp), opac=_f32(opacities),
                scales=_f32(scales), rotations=_f32(rotations), cov=_f32(cov3D_precomp), view=_f32(viewmatrix),
                proj=_f32(projmatrix), campos=_f32(campos), bg=_f32(bg))
    P = 0 if keep["means3D"] is None else keep["means3D"].shape[0]
    M = 0 if keep["shs"] is None else keep["shs"].shape[1]
    ctx = lib.oracle_forward(P, int(sh_degree), M, _p(keep["bg"]), int(width), int(height), _p(keep["means3D"]),
                             _p(keep["shs"]), _p(keep["colors"]), _p(keep["opac"]), _p(keep["scales"]),
                             float(scale_modifier), _p(keep["rotations"]), _p(keep["cov"]), _p(keep["view"]),
                             _p(keep["proj"]), _p(keep["campos"]), float(tanfovx), float(tanfovy))
    return OracleResult(lib, ctx, real, keep, (P, int(width), int(height), M))


def forward_scene(scene, cam, bg, precision="f64", **kw):
    """Convenience for saro_gs_b200.synthetic Scene/Camera tuples."""
    return forward(scene.means3D, scene.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, bg, cam.width,
                   cam.height, cam.tanfovx, cam.tanfovy, sh_degree=scene.sh_degree, shs=scene.shs,
                   scales=scene.scales, rotations=scene.rotations, precision=precision, **kw)
