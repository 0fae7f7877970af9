"""Host-side mirror of the reference's scale-aware residual field — same class name, constructor, buffers, parameter
names (state_dict compatible: ``grids.<level>.<plane>``) and methods as ``scene/hexplane.py:ScaleAwareResField`` —
with ``forward`` running on the sm_100a plane-sampler kernels behind the C ABI (sgs_plane_* in
include/saro_gs_b200.h) instead of ``nvdiffrast.torch.texture`` + ~60 PyTorch ops per call.

    from saro_gs_b200.hexplane import ScaleAwareResField      # drop-in for scene.hexplane.ScaleAwareResField

What is mirrored (reference file:line):
    __init__ / init_grid_param (planes zero-initialised, [1, C, reso[b], reso[a]] for the pair (a, b))   hexplane.py:62-89, 160-200
    set_aabb (buffers aabb, duration, max_level, base_scale)                                           hexplane.py:205-229
    get_level                                                                                          hexplane.py:231-242
    forward == get_density                                                                             hexplane.py:247-274
    get_grid_parameters, planetv, timesmooth                                                           hexplane.py:276-325

Only the planes receive gradients: the reference always samples at detached positions / times / scales
(scene/saro_gaussian.py:765,780,865); a position or scale that requires a gradient raises instead of silently
returning none.  There is no PyTorch fallback: without the native library the module raises.
"""
import ctypes
import itertools

import torch
import torch.nn as nn

from . import _lib


class UnsupportedPlaneConfig(RuntimeError):
    pass


def init_grid_param(grid_nd, in_dim, out_dim, reso):
    """hexplane.py:62-89 — one zero-initialised plane per coordinate pair, shape [1, out_dim, reso[b], reso[a]]."""
    assert in_dim == len(reso), "Resolution must have same number of elements as input-dimension"
    assert grid_nd <= in_dim
    coo_combs = list(itertools.combinations(range(in_dim), grid_nd))
    grid_coefs = nn.ParameterList()
    for coo_comb in coo_combs:
        grid_coefs.append(nn.Parameter(torch.zeros([1, out_dim] + [reso[cc] for cc in coo_comb[::-1]])))
    return grid_coefs


def compute_plane_smoothness(t):
    """hexplane.py:139-146"""
    h = t.shape[2]
    first_difference = t[..., 1:, :] - t[..., :h - 1, :]
    second_difference = first_difference[..., 1:, :] - first_difference[..., :h - 2, :]
    return torch.square(second_difference).mean()


def compute_plane_tv(t):
    """hexplane.py:148-155"""
    batch_size, c, h, w = t.shape
    count_h = batch_size * c * (h - 1) * w
    count_w = batch_size * c * h * (w - 1)
    h_tv = torch.square(t[..., 1:, :] - t[..., :h - 1, :]).sum()
    w_tv = torch.square(t[..., :, 1:] - t[..., :, :w - 1]).sum()
    return 2 * (h_tv / count_h + w_tv / count_w)


def _check(code, what):
    if code < 0:
        raise RuntimeError(f"{what} failed ({code}): {_lib.last_error()}")
    return code


class _Pyramids:
    """Channels-last mip pyramids of one resolution level's six planes, rebuilt when a plane's version changes."""

    def __init__(self):
        self.buf = {}        # plane index -> tensor
        self.version = {}    # plane index -> (data_ptr, _version)

    def get(self, lib, ci, plane, max_mip):
        _, C, H, W = plane.shape
        key = (plane.data_ptr(), plane._version, plane.device)
        if self.version.get(ci) != key:
            n = lib.sgs_plane_pyramid_floats(C, H, W, max_mip)
            if n == 0:
                raise UnsupportedPlaneConfig(f"plane [{C}, {H}, {W}]: extents must be even at every mip level that is built")
            buf = self.buf.get(ci)
            if buf is None or buf.numel() != n or buf.device != plane.device:
                buf = torch.empty(n, dtype=torch.float32, device=plane.device)
            stream = torch.cuda.current_stream(plane.device).cuda_stream
            _check(lib.sgs_plane_build(C, H, W, max_mip, plane.data_ptr(), buf.data_ptr(), ctypes.c_void_p(stream)),
                   "sgs_plane_build")
            self.buf[ci] = buf
            self.version[ci] = key
        return self.buf[ci]


COO_COMBS = list(itertools.combinations(range(4), 2))


def _descs(pyramids, shapes):
    arr = (_lib.PlaneDesc * 6)()
    for ci, comb in enumerate(COO_COMBS):
        _, _, H, W = shapes[ci]
        arr[ci] = _lib.PlaneDesc(pyramids[ci].data_ptr(), H, W, comb[0], comb[1], 0 if 3 in comb else 7)
    return arr


class _PlaneField(torch.autograd.Function):
    @staticmethod
    def forward(ctx, field, pts, timestamps, scales, *planes):
        lib = _lib.load()
        dev = pts.device
        N = int(pts.shape[0])
        n_levels = len(field.grids)
        C = int(planes[0].shape[1])
        out = torch.empty((N, C * n_levels), dtype=torch.float32, device=dev)
        reso0 = (ctypes.c_int * 3)(*[int(r) for r in field.reso_list[0][:3]])
        time_scale = float(field._duration_py) / (float(field._duration_py) - 1.0)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for li in range(n_levels):
                level_planes = planes[6 * li:6 * li + 6]
                pyr = [field._pyramids[li].get(lib, ci, p.detach(), 0 if 3 in COO_COMBS[ci] else 7)
                       for ci, p in enumerate(level_planes)]
                descs = _descs(pyr, [p.shape for p in level_planes])
                _check(lib.sgs_plane_sample_forward(N, C, pts.data_ptr(), timestamps.data_ptr(), scales.data_ptr(),
                                                    field.aabb.data_ptr(), field.base_scale.data_ptr(), time_scale,
                                                    reso0, 6, descs, C * n_levels, C * li, out.data_ptr(), stream),
                       "sgs_plane_sample_forward")
        ctx.field = field
        ctx.shapes = [tuple(p.shape) for p in planes]
        ctx.save_for_backward(pts, timestamps, scales)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        field = ctx.field
        pts, timestamps, scales = ctx.saved_tensors
        dev = pts.device
        N = int(pts.shape[0])
        n_levels = len(field.grids)
        C = int(ctx.shapes[0][1])
        dout = dout.contiguous()
        reso0 = (ctypes.c_int * 3)(*[int(r) for r in field.reso_list[0][:3]])
        time_scale = float(field._duration_py) / (float(field._duration_py) - 1.0)
        grads = []
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for li in range(n_levels):
                shapes = ctx.shapes[6 * li:6 * li + 6]
                sizes = [lib.sgs_plane_pyramid_floats(C, s[2], s[3], 0 if 3 in COO_COMBS[ci] else 7)
                         for ci, s in enumerate(shapes)]
                flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)   # one memset for the six pyramids
                gp, off = [], 0
                for n in sizes:
                    gp.append(flat[off:off + n])
                    off += n
                descs = _descs(gp, shapes)
                _check(lib.sgs_plane_sample_backward(N, C, pts.data_ptr(), timestamps.data_ptr(), scales.data_ptr(),
                                                     field.aabb.data_ptr(), field.base_scale.data_ptr(), time_scale,
                                                     reso0, 6, descs, C * n_levels, C * li, dout.data_ptr(), stream),
                       "sgs_plane_sample_backward")
                for ci, s in enumerate(shapes):
                    d = torch.empty(s, dtype=torch.float32, device=dev)
                    _check(lib.sgs_plane_fold(C, s[2], s[3], 0 if 3 in COO_COMBS[ci] else 7, gp[ci].data_ptr(),
                                              d.data_ptr(), stream), "sgs_plane_fold")
                    grads.append(d)
        return (None, None, None, None) + tuple(grads)


class ScaleAwareResField(nn.Module):
    def __init__(self, planeconfig, multires) -> None:
        super().__init__()
        self.grid_config = [planeconfig]
        self.multiscale_res_multipliers = multires
        self.concat_features = True
        self.concat_plane = False
        if planeconfig["grid_dimensions"] != 2 or planeconfig["input_coordinate_dim"] != 4:
            raise UnsupportedPlaneConfig("the plane sampler implements the reference's configuration: 2-D planes of a 4-D field")

        self.grids = nn.ModuleList()
        self.feat_dim = 0
        self.reso_list = []
        for res in self.multiscale_res_multipliers:
            config = self.grid_config[0].copy()
            config["resolution"] = [r * res for r in config["resolution"][:3]] + config["resolution"][3:]
            self.reso_list.append(config["resolution"])
            gp = init_grid_param(grid_nd=config["grid_dimensions"], in_dim=config["input_coordinate_dim"],
                                 out_dim=config["output_coordinate_dim"], reso=config["resolution"])
            self.feat_dim += gp[-1].shape[1]
            self.grids.append(gp)
        self._pyramids = [_Pyramids() for _ in self.grids]
        self._duration_py = None
        print("feature_dim:", self.feat_dim)

    @property
    def get_aabb(self):
        return self.aabb[0], self.aabb[1]

    def set_aabb(self, xyz_max, xyz_min, duration):
        dev = self.grids[0][0].device     # the reference hard-codes "cuda" (one GPU per process, cuda:0)
        aabb = torch.tensor([xyz_max, xyz_min], dtype=torch.float32, device=dev)
        self.register_buffer("aabb", aabb)
        self.register_buffer("duration", torch.tensor([duration]))
        self._duration_py = float(duration)
        self.per_grid_size = []
        for res in self.reso_list:
            self.per_grid_size.append([(xyz_max[i] - xyz_min[i]) / res[i] for i in range(3)])
        max_level = torch.log(torch.tensor(self.reso_list[0]))
        base_scale = torch.tensor(self.per_grid_size[0], dtype=torch.float32, device=dev)
        self.register_buffer("max_level", max_level)
        self.register_buffer("base_scale", base_scale)

    def get_level(self, scales: torch.Tensor):
        min_scale = self.base_scale / 2
        max_scale = min_scale * torch.tensor(self.reso_list[0][:3]).to(min_scale)
        scales = torch.clamp(scales, min_scale, max_scale)
        level = torch.log2(2 * scales / self.base_scale.unsqueeze(0))
        level = torch.cat((level, torch.zeros((level.shape[0], 1)).to(level)), dim=-1)
        level[:, 3] = 0.0
        return level

    def set_base_scale(self, scale):
        pass

    def get_density(self, pts, timestamps=None, scales=None):
        if self._duration_py is None:
            if not hasattr(self, "duration"):
                raise RuntimeError("ScaleAwareResField: call set_aabb(xyz_max, xyz_min, duration) first")
            self._duration_py = float(self.duration.item())     # buffers restored by load_state_dict
        if not pts.is_cuda:
            raise RuntimeError("saro_gs_b200.hexplane: inputs must be CUDA tensors — the plane sampler has no CPU path")
        for name, t in (("pts", pts), ("timestamps", timestamps), ("scales", scales)):
            if t.requires_grad:
                raise UnsupportedPlaneConfig(
                    f"{name} requires a gradient: the plane sampler differentiates the planes only (the reference "
                    "samples at detached positions / times / scales, scene/saro_gaussian.py:765,780,865)")
        pts = pts.reshape(-1, pts.shape[-1]).contiguous().float()
        N = pts.shape[0]
        if N < 1:
            return torch.zeros((0, 1), device=pts.device)
        timestamps = timestamps.reshape(-1).contiguous().float()
        scales = scales.reshape(-1, 3).contiguous().float()
        if timestamps.shape[0] != N or scales.shape[0] != N:
            raise RuntimeError("pts, timestamps and scales must describe the same number of points")
        planes = [p for level in self.grids for p in level]
        C = planes[0].shape[1]
        for p in planes:
            if p.shape[1] != C or p.dtype != torch.float32 or not p.is_contiguous():
                raise UnsupportedPlaneConfig("planes must be contiguous float32 with one feature width")
        return _PlaneField.apply(self, pts, timestamps, scales, *planes)

    def forward(self, pts, timestamps=None, scales=None):
        return self.get_density(pts, timestamps, scales)

    @property
    def get_grid_parameters(self):
        return self.grids.parameters()

    def planetv(self):
        total = 0
        for grids in self.grids:
            for spatial_plane in [0, 1, 3]:
                total += compute_plane_tv(grids[spatial_plane])
        return total

    def timesmooth(self):
        total = 0
        for grids in self.grids:
            for time_idx in [1, 4, 5]:
                total += compute_plane_smoothness(grids[time_idx])
        return total
