"""Per-frame deformation -> rasterizer hand-off.

Host-side mirror of the reference's ``GaussianModel.get_deformation_eval`` (scene/saro_gaussian.py:871-921): same
name, same argument, same return order ``(means3D, rotations, scales, opacity, shs)`` — so a SaRO-GS maintainer can do

    from saro_gs_b200.deformation import get_deformation_eval
    GaussianModel.get_deformation_eval = get_deformation_eval

and ``renderer/__init__.py:190`` runs on the fused tcgen05 kernel unchanged.  All arithmetic is in
``csrc/sgs_deform.cu`` behind the C ABI (``sgs_deform_pack_mlp`` / ``sgs_deform_eval``); torch is used for device
memory and the current stream only.  There is no PyTorch fallback: CPU tensors, other dtypes or an unsupported
configuration raise.
"""
import torch

from . import _lib

TIME_DIMS = 9          # get_embedder(4): x, sin/cos of 1x, 2x, 4x, 8x  (saro_gaussian.py:94, :922-969)
HIDDEN = 128           # args.deform_hidden_dim                       (arguments/__init__.py:65)
_OUT_DIMS = (3, 7, 48)


class UnsupportedDeformationConfig(NotImplementedError):
    pass


def _ptr(t):
    return t.data_ptr()


def _check(t, name, shape_tail, n=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if n is not None and (t.shape[0] != n or tuple(t.shape[1:]) not in shape_tail):
        raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected ({n}, {shape_tail[0]})")
    return t.detach().contiguous()


def _linears(mlp):
    """(W1, b1, W2, b2, W3, b3) of nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) or of a 6-tuple of tensors."""
    if isinstance(mlp, (tuple, list)):
        if len(mlp) != 6:
            raise UnsupportedDeformationConfig("an MLP is (W1, b1, W2, b2, W3, b3)")
        return tuple(mlp)
    layers = [m for m in mlp if hasattr(m, "weight")]
    if len(layers) != 3:
        raise UnsupportedDeformationConfig(f"expected a 3-layer MLP, found {len(layers)} linear layers")
    out = []
    for layer in layers:
        out += [layer.weight, layer.bias]
    return tuple(out)


class PackedMLPs:
    """The three deformation MLPs in the tensor-core layout the kernel keeps resident (one device buffer)."""

    def __init__(self, motion_mlp, rot_mlp, shs_mlp):
        lib = _lib.load()
        params = [_linears(m) for m in (motion_mlp, rot_mlp, shs_mlp)]
        dev = params[0][0].device
        in_dim = params[0][0].shape[1]
        for m, (ps, out_dim) in enumerate(zip(params, _OUT_DIMS)):
            W1, b1, W2, b2, W3, b3 = ps
            want = [(HIDDEN, in_dim), (HIDDEN,), (HIDDEN, HIDDEN), (HIDDEN,), (out_dim, HIDDEN), (out_dim,)]
            got = [tuple(t.shape) for t in ps]
            if got != want:
                raise UnsupportedDeformationConfig(f"MLP {m}: parameter shapes {got}, supported {want}")
        if in_dim - TIME_DIMS not in (8, 16, 24, 32):
            raise UnsupportedDeformationConfig(
                f"MLP input width {in_dim} = plane feature width {in_dim - TIME_DIMS} + {TIME_DIMS}: the kernel supports "
                "plane feature widths 8, 16, 24 and 32 (every shipped config uses 16 or 32)")
        self.in_dim = in_dim
        self.feat_dim = in_dim - TIME_DIMS
        self.buffer = torch.empty(lib.sgs_deform_packed_bytes(), dtype=torch.uint8, device=dev)
        self.versions = None
        # modules are re-read on every refresh (they may swap their parameters); plain tuples are fixed
        self._layers = [None if isinstance(m, (tuple, list)) else [l for l in m if hasattr(l, "weight")]
                        for m in (motion_mlp, rot_mlp, shs_mlp)]
        self._sources = params
        self.refresh()

    def _current_versions(self):
        key = []
        for i, layers in enumerate(self._layers):
            if layers is not None:
                self._sources[i] = tuple(t for l in layers for t in (l.weight, l.bias))
            for t in self._sources[i]:
                key.append((t.data_ptr(), t._version))
        return key

    def refresh(self):
        """Re-pack if any weight tensor was written since the last call (optimizer step, load_state_dict)."""
        cur = self._current_versions()
        if cur == self.versions:
            return
        lib = _lib.load()
        stream = torch.cuda.current_stream(self.buffer.device).cuda_stream
        with torch.cuda.device(self.buffer.device):
            for m, ps in enumerate(self._sources):
                ts = [_check(t, f"mlp{m} parameter", None) for t in ps]
                rc = lib.sgs_deform_pack_mlp(m, self.in_dim, *[_ptr(t) for t in ts], _ptr(self.buffer), stream)
                if rc != 0:
                    raise RuntimeError(f"sgs_deform_pack_mlp failed ({rc}): {_lib.last_error()}")
        self.versions = cur


def _validate_inputs(packed, xyz, rotation, scaling, opacity, features_dc, features_rest, temporal_pos, lifespan,
                     hexplane_feature):
    n = xyz.shape[0]
    return n, (_check(xyz, "xyz", [(3,)], n), _check(rotation, "rotation", [(4,)], n), _check(scaling, "scaling", [(3,)], n),
               _check(opacity, "opacity", [(1,), ()], n), _check(features_dc, "features_dc", [(1, 3), (3,)], n),
               _check(features_rest, "features_rest", [(15, 3), (45,)], n),
               _check(temporal_pos, "temporal_pos", [(1,), ()], n), _check(lifespan, "lifespan", [(1,), ()], n),
               _check(hexplane_feature, "hexplane_feature", [(packed.feat_dim,)], n))


def _run(timestamp, n, inputs, packed, workspace):
    lib = _lib.load()
    packed.refresh()
    dev = inputs[0].device
    opts = dict(dtype=torch.float32, device=dev)
    means3D = torch.empty((n, 3), **opts)
    rot = torch.empty((n, 4), **opts)
    scale = torch.empty((n, 3), **opts)
    opa = torch.empty((n, 1), **opts)
    shs = torch.empty((n, 16, 3), **opts)
    if n == 0:
        return means3D, rot, scale, opa, shs
    ws_bytes = lib.sgs_deform_workspace_bytes(n)
    if workspace is None or workspace.numel() < ws_bytes or workspace.device != dev:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        sel = lib.sgs_deform_eval(n, packed.feat_dim, float(timestamp), *[t.data_ptr() for t in inputs], _ptr(packed.buffer),
                                  _ptr(workspace), workspace.numel(), _ptr(means3D), _ptr(rot), _ptr(scale), _ptr(opa),
                                  _ptr(shs), stream)
    if sel < 0:
        raise RuntimeError(f"sgs_deform_eval failed ({sel}): {_lib.last_error()}")
    # `workspace` may be released by the caller right away: the caching allocator is stream-ordered, like any torch op
    return means3D[:sel], rot[:sel], scale[:sel], opa[:sel], shs[:sel]


def deformation_eval(timestamp, xyz, rotation, scaling, opacity, features_dc, features_rest, temporal_pos, lifespan,
                     hexplane_feature, packed, workspace=None):
    """Explicit-tensor form.  Returns (means3D [S,3], rotations [S,4], scales [S,3], opacity [S,1], shs [S,16,3])
    for the S Gaussians whose survival state exceeds 0.001 at `timestamp`, in source order."""
    n, inputs = _validate_inputs(packed, xyz, rotation, scaling, opacity, features_dc, features_rest, temporal_pos, lifespan,
                                 hexplane_feature)
    return _run(timestamp, n, inputs, packed, workspace)


def get_deformation_eval(self, timestamp, rays=None):
    """Drop-in for GaussianModel.get_deformation_eval (scene/saro_gaussian.py:871-921); `self` is the Gaussian model
    (after get_deformfeature(), :863-869, which caches `hexplane_feature` and `_lifespan`)."""
    args = self.args
    if not (args.dx and args.drot and args.dopacity and args.dsh):
        raise UnsupportedDeformationConfig(
            "the fused hand-off implements the configuration all shipped configs use (dx, drot, dopacity, dsh all on); "
            f"got dx={args.dx} drot={args.drot} dopacity={args.dopacity} dsh={args.dsh}")
    cache = getattr(self, "_sgs_deform_cache", None)
    key = (id(self.motion_mlp), id(self.rot_mlp), id(self.shs_mlp))
    if cache is None or cache["key"] != key:
        cache = {"key": key, "packed": PackedMLPs(self.motion_mlp, self.rot_mlp, self.shs_mlp), "workspace": None}
        self._sgs_deform_cache = cache
    # the model's tensors are validated once and re-used while the model keeps the same tensor objects and storage
    # (densification, load_state_dict and .to() all replace them)
    raw = (self._xyz, self._rotation, self._scaling, self._opacity, self._features_dc, self._features_rest,
           self.get_temporalpos, self._lifespan, self.hexplane_feature)
    known = cache.get("raw")
    if known is None or any(a is not b for a, b in zip(raw, known)) or \
            any(a.data_ptr() != b.data_ptr() for a, b in zip(raw, cache["inputs"])):
        cache["n"], cache["inputs"] = _validate_inputs(cache["packed"], *raw)
        cache["raw"] = raw
    n = cache["n"]
    need = _lib.load().sgs_deform_workspace_bytes(n)
    ws = cache["workspace"]
    if ws is None or ws.numel() < need or ws.device != self._xyz.device:
        ws = cache["workspace"] = torch.empty(need, dtype=torch.uint8, device=self._xyz.device)
    if torch.is_tensor(timestamp):
        timestamp = float(timestamp)
    return _run(timestamp, n, cache["inputs"], cache["packed"], ws)
