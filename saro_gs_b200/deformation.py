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
        self.buffer = torch.zeros(lib.sgs_deform_packed_bytes(), dtype=torch.uint8, device=dev)   # the kernel copies whole images: no stale bytes
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


# ======================================================================================================================
# Training path: GaussianModel.get_deformation (scene/saro_gaussian.py:779-847)
# ======================================================================================================================
_TRAIN_MLPS = ("motion_mlp", "rot_mlp", "shs_mlp", "opacity_mlp")


class TrainImages:
    """Forward and data-gradient tensor-core images of the four MLPs get_deformation evaluates (motion, rot, shs and
    opacity_mlp, scene/saro_gaussian.py:102-108); re-packed when a weight tensor was written (optimizer step)."""

    def __init__(self, motion_mlp, rot_mlp, shs_mlp, opacity_mlp):
        lib = _lib.load()
        self._modules = (motion_mlp, rot_mlp, shs_mlp, opacity_mlp)
        self.shapes = []
        params = [_linears(m) for m in self._modules]
        dev = params[0][0].device
        in_w = params[0][0].shape[1]
        self.feat_dim = in_w - TIME_DIMS
        if self.feat_dim not in (8, 16, 24, 32):
            raise UnsupportedDeformationConfig(f"plane feature width {self.feat_dim}: the kernels support 8, 16, 24 and 32")
        for m, ps in enumerate(params):
            W1, b1, W2, b2, W3, b3 = ps
            w_in, hid2, n_out = W1.shape[1], W2.shape[0], W3.shape[0]
            ok = (W1.shape[0] == HIDDEN and W2.shape[1] == HIDDEN and W3.shape[1] == hid2 and hid2 <= HIDDEN and
                  w_in == (self.feat_dim if m == 3 else in_w) and (n_out <= 8 or n_out == 48) and
                  tuple(b1.shape) == (HIDDEN,) and tuple(b2.shape) == (hid2,) and tuple(b3.shape) == (n_out,))
            if not ok:
                raise UnsupportedDeformationConfig(f"{_TRAIN_MLPS[m]}: parameter shapes {[tuple(t.shape) for t in ps]} are not "
                                                   f"a {w_in}-128-(<=128)-(<=8 | 48) MLP")
            self.shapes.append((w_in, hid2, n_out))
        nbytes = lib.sgs_deform_image_bytes()
        self.stride = (nbytes + 255) // 256 * 256
        self.buffer = torch.zeros(8 * self.stride, dtype=torch.uint8, device=dev)      # [mlp][forward | backward]; zeroed: the kernels copy whole images
        self.versions = None

    def params(self):
        """The 24 parameter tensors as they are now (modules may swap them)."""
        return [t for m in self._modules for t in _linears(m)]

    def image(self, mlp, backward):
        return self.buffer.data_ptr() + (2 * mlp + int(backward)) * self.stride

    def refresh(self, params):
        cur = [(t.data_ptr(), t._version) for t in params]
        if cur == self.versions:
            return
        lib = _lib.load()
        stream = torch.cuda.current_stream(self.buffer.device).cuda_stream
        with torch.cuda.device(self.buffer.device):
            for m in range(4):
                ts = [_check(t, f"{_TRAIN_MLPS[m]} parameter", None) for t in params[6 * m:6 * m + 6]]
                w_in, hid2, n_out = self.shapes[m]
                for backward in (0, 1):
                    rc = lib.sgs_deform_pack_general(backward, w_in, hid2, n_out, self.feat_dim, *[_ptr(t) for t in ts],
                                                     self.image(m, backward), stream)
                    if rc != 0:
                        raise RuntimeError(f"sgs_deform_pack_general failed ({rc}): {_lib.last_error()}")
        self.versions = cur


_phase_marks = None      # bench.py sets this to a list to collect (name, CUDA event) marks of the backward's phases


def _mark(name):
    if _phase_marks is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        _phase_marks.append((name, e))


class _TrainMLPs(torch.autograd.Function):
    """All MLP evaluations of one get_deformation call: one tcgen05 launch forward, one for the data-gradient chains.
    jobs: tuple of (mlp index, zero_time, differentiable).  Returns one raw output [N, n_out] per job."""

    @staticmethod
    def forward(ctx, feat, tpos, timestamp, jobs, images, *params):
        lib = _lib.load()
        images.refresh(params)
        n, F = feat.shape[0], images.feat_dim
        dev = feat.device
        feat_c = _check(feat, "hexplane feature", [(F,)], n)
        tpos_c = _check(tpos, "temporal_pos", [(1,), ()], n)
        want = any(ctx.needs_input_grad)          # all False under no_grad: nothing is saved then
        outs, keep = [], []
        arr = (_lib.MLPJob * len(jobs))()
        f32 = dict(dtype=torch.float32, device=dev)
        planes = lambda groups: torch.empty(lib.sgs_deform_planes_bytes(n, groups), dtype=torch.uint8, device=dev)
        x_planes = {}                                        # the MLP input is the same for every job of a time mode
        for k, (m, zero_time, diff) in enumerate(jobs):
            n_out = images.shapes[m][2]
            out = torch.empty((n, n_out), **f32)
            outs.append(out)
            save = want and diff and n > 0
            h1 = planes(16) if save else None
            h2 = planes(16) if save else None
            m1 = torch.empty((n, 4), dtype=torch.int32, device=dev) if save else None
            m2 = torch.empty((n, 4), dtype=torch.int32, device=dev) if save else None
            xin = None
            if save and bool(zero_time) not in x_planes:
                xin = x_planes[bool(zero_time)] = planes(6)
            keep.append((h1, h2, m1, m2, bool(zero_time)))
            p0 = lambda t: t.data_ptr() if t is not None else None
            arr[k] = _lib.MLPJob(images.image(m, 0), None, out.data_ptr(), p0(h1), p0(h2), p0(xin), p0(m1), p0(m2), n_out,
                                 int(zero_time))
        if n > 0:
            with torch.cuda.device(dev):
                rc = lib.sgs_deform_train_forward(n, F, float(timestamp), tpos_c.data_ptr(), feat_c.data_ptr(), len(jobs), arr,
                                                  torch.cuda.current_stream(dev).cuda_stream)
            if rc != 0:
                raise RuntimeError(f"sgs_deform_train_forward failed ({rc}): {_lib.last_error()}")
        ctx.jobs, ctx.images, ctx.keep, ctx.x_planes = jobs, images, keep, x_planes
        ctx.save_for_backward(feat_c, tpos_c, *params)
        ctx.mark_non_differentiable(*[o for o, (_, _, diff) in zip(outs, jobs) if not diff])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        lib = _lib.load()
        feat, tpos, *params = ctx.saved_tensors
        jobs, images, keep = ctx.jobs, ctx.images, ctx.keep
        n, F = feat.shape
        dev = feat.device
        f32 = dict(dtype=torch.float32, device=dev)
        live = [k for k, (m, z, diff) in enumerate(jobs) if diff and gouts[k] is not None and keep[k][0] is not None]
        grads = [None] * len(params)
        if not live or n == 0:
            return (torch.zeros_like(feat) if ctx.needs_input_grad[0] else None, None, None, None, None, *grads)
        _mark("bwd_begin")
        planes = lambda groups: torch.empty(lib.sgs_deform_planes_bytes(n, groups), dtype=torch.uint8, device=dev)
        arr = (_lib.MLPJob * len(live))()
        slabs = torch.empty((len(live), n, F), **f32)
        work = []
        for i, k in enumerate(live):
            m, zero_time, _ = jobs[k]
            h1, h2, m1, m2, zt = keep[k]
            dy = gouts[k].contiguous()
            dh2, dh1 = planes(16), planes(16)
            dyp = planes(6 if dy.shape[1] > 8 else 2)
            work.append((m, zt, dy, dyp, h1, h2, dh1, dh2))
            arr[i] = _lib.MLPJob(images.image(m, 1), dy.data_ptr(), slabs[i].data_ptr(), dh2.data_ptr(), dh1.data_ptr(), dyp.data_ptr(),
                                 m2.data_ptr(), m1.data_ptr(), dy.shape[1], 0)
        with torch.cuda.device(dev):
            rc = lib.sgs_deform_train_backward(n, F, len(live), arr, torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"sgs_deform_train_backward failed ({rc}): {_lib.last_error()}")
        _mark("data_gradient_kernel")
        dfeat = slabs.sum(dim=0) if len(live) > 1 else slabs[0]
        # weight gradients: one TMA + tcgen05 launch for all (job, layer) GEMMs dW = G^T X over the rows — both operands
        # of every GEMM are operand planes the forward / data-gradient kernels emitted — then a deterministic reduction
        # of the per-CTA partials straight into parameter-shaped tensors.  A second evaluation of the same MLP (base
        # feature) accumulates.
        out = {}
        tasks = []
        for m, zt, dy, dyp, h1, h2, dh1, dh2 in work:
            w_in, hid2, n_out = images.shapes[m]
            acc = int(m in out)
            if not acc:
                out[m] = [torch.empty((HIDDEN, w_in), **f32), torch.empty(HIDDEN, **f32), torch.empty((hid2, HIDDEN), **f32),
                          torch.empty(hid2, **f32), torch.empty((n_out, hid2), **f32), None]
            dW1, db1, dW2, db2, dW3, _ = out[m]
            tasks += [_lib.WgradTask(dh1.data_ptr(), ctx.x_planes[zt].data_ptr(), 6, dW1.data_ptr(), w_in, HIDDEN, w_in, 0, db1.data_ptr(), acc),
                      _lib.WgradTask(dh2.data_ptr(), h1.data_ptr(), 16, dW2.data_ptr(), HIDDEN, hid2, HIDDEN, 0, db2.data_ptr(), acc),
                      _lib.WgradTask(h2.data_ptr(), dyp.data_ptr(), 6 if n_out > 8 else 2, dW3.data_ptr(), hid2, hid2, n_out, 1, None, acc)]
            db3 = dy.sum(0)
            out[m][5] = db3 if out[m][5] is None else out[m][5] + db3
        arr_w = (_lib.WgradTask * len(tasks))(*tasks)
        partials = torch.empty(lib.sgs_deform_wgrad_max_ctas() * lib.sgs_deform_wgrad_partial_floats(), **f32)
        with torch.cuda.device(dev):
            rc = lib.sgs_deform_wgrad(n, len(tasks), arr_w, partials.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"sgs_deform_wgrad failed ({rc}): {_lib.last_error()}")
        for m, gs in out.items():
            for q, g in enumerate(gs):
                if ctx.needs_input_grad[5 + 6 * m + q]:
                    grads[6 * m + q] = g
        _mark("weight_gradients")
        ctx.keep = None
        return (dfeat if ctx.needs_input_grad[0] else None, None, None, None, None, *grads)


class _TrainEpilogue(torch.autograd.Function):
    """The elementwise statements between the MLP outputs and the rasterizer's inputs (scene/saro_gaussian.py:782-831)
    as one kernel forward and one backward.  Returns (means3D, rotations, scales, opacity, lifespan, real_xyz)."""

    @staticmethod
    def forward(ctx, timestamp, min_scale, life_raw, motion_raw, rot_raw, motion_base_raw, xyz, rotation, scaling, opacity, tpos):
        lib = _lib.load()
        n = xyz.shape[0]
        dev = xyz.device
        c = lambda t, name, tail: _check(t, name, tail, n)
        ins = (c(life_raw, "lifespan MLP output", [(1,)]), c(motion_raw, "motion MLP output", [(3,)]), c(rot_raw, "rot MLP output", [(7,)]),
               c(motion_base_raw, "motion MLP output (base)", [(3,)]), c(xyz, "xyz", [(3,)]), c(rotation, "rotation", [(4,)]),
               c(scaling, "scaling", [(3,)]), c(opacity, "opacity", [(1,), ()]), c(tpos, "temporal_pos", [(1,), ()]))
        f32 = dict(dtype=torch.float32, device=dev)
        outs = (torch.empty((n, 3), **f32), torch.empty((n, 4), **f32), torch.empty((n, 3), **f32), torch.empty((n, 1), **f32),
                torch.empty((n, 1), **f32), torch.empty((n, 3), **f32))
        if n > 0:
            with torch.cuda.device(dev):
                rc = lib.sgs_deform_train_epilogue_forward(n, float(timestamp), float(min_scale), *[t.data_ptr() for t in ins],
                                                           *[t.data_ptr() for t in outs], torch.cuda.current_stream(dev).cuda_stream)
            if rc != 0:
                raise RuntimeError(f"sgs_deform_train_epilogue_forward failed ({rc}): {_lib.last_error()}")
        ctx.consts = (float(timestamp), float(min_scale), tuple(opacity.shape), tuple(tpos.shape))
        ctx.save_for_backward(ins[0], ins[2], ins[5], ins[6], ins[7], ins[8])
        ctx.mark_non_differentiable(outs[5])
        return outs

    @staticmethod
    def backward(ctx, g_motion, g_rot, g_scale, g_op, g_life, _g_real):
        lib = _lib.load()
        life_raw, rot_raw, rotation, scaling, opacity, tpos = ctx.saved_tensors
        timestamp, min_scale, op_shape, tpos_shape = ctx.consts
        n = rotation.shape[0]
        dev = rotation.device
        f32 = dict(dtype=torch.float32, device=dev)
        d_life, d_rot_raw = torch.empty((n, 1), **f32), torch.empty((n, 7), **f32)
        d_rotation, d_scaling = torch.empty((n, 4), **f32), torch.empty((n, 3), **f32)
        d_opacity, d_tpos = torch.empty(op_shape, **f32), torch.empty(tpos_shape, **f32)
        gs = [None if g is None else g.contiguous() for g in (g_rot, g_scale, g_op, g_life)]
        if n > 0:
            with torch.cuda.device(dev):
                rc = lib.sgs_deform_train_epilogue_backward(
                    n, timestamp, min_scale, life_raw.data_ptr(), rot_raw.data_ptr(), rotation.data_ptr(), scaling.data_ptr(),
                    opacity.data_ptr(), tpos.data_ptr(), *[None if g is None else g.data_ptr() for g in gs], d_life.data_ptr(),
                    d_rot_raw.data_ptr(), d_rotation.data_ptr(), d_scaling.data_ptr(), d_opacity.data_ptr(), d_tpos.data_ptr(),
                    torch.cuda.current_stream(dev).cuda_stream)
            if rc != 0:
                raise RuntimeError(f"sgs_deform_train_epilogue_backward failed ({rc}): {_lib.last_error()}")
        # means3D = xyz + motion_raw: both gradients are the incoming one
        return (None, None, d_life, g_motion, d_rot_raw, None, g_motion, d_rotation, d_scaling, d_opacity, d_tpos)


def get_deformation(self, timestamp, rays=None):
    """Drop-in for GaussianModel.get_deformation (scene/saro_gaussian.py:779-847): same side effects (`_lifespan`,
    `scale_residual`, `shs_residual`, `motion_residual`, `real_xyz`), same return order.  The plane field is whatever
    module the model carries (`saro_gs_b200.hexplane.ScaleAwareResField` for the native sampler); the seven MLP
    evaluations and their backward run in the tcgen05 kernels of csrc/sgs_deform.cu, the elementwise statements between
    them and the returned tensors in one kernel forward and one backward (`_TrainEpilogue`); only the SH residual add
    (`cat` + add) is left to PyTorch."""
    args = self.args
    if not (args.dx and args.drot and args.dopacity and args.dsh):
        raise UnsupportedDeformationConfig(
            "the fused training path implements the configuration all shipped configs use (dx, drot, dopacity, dsh all on); "
            f"got dx={args.dx} drot={args.drot} dopacity={args.dopacity} dsh={args.dsh}")
    if not self._xyz.is_cuda:
        raise RuntimeError("get_deformation: the model must live on a CUDA device (there is no CPU path)")
    cache = getattr(self, "_sgs_train_cache", None)
    key = tuple(id(getattr(self, k)) for k in _TRAIN_MLPS)
    if cache is None or cache["key"] != key:
        cache = {"key": key, "images": TrainImages(*[getattr(self, k) for k in _TRAIN_MLPS])}
        self._sgs_train_cache = cache
    images = cache["images"]

    hexplane_feature = self.hexplane(self._xyz.detach(), self.get_temporalpos.detach(), self.get_scaling.detach())   # :780
    jobs = [(3, True, True), (0, False, True), (1, False, True), (2, False, True)]       # lifespan, motion, rot, shs at t
    names = ["life", "motion", "rot", "shs"]
    if args.scale_reg:                                                                    # :796-797
        jobs.append((1, True, True)); names.append("rot_base")
    if args.shs_reg:                                                                      # :798-799
        jobs.append((2, True, True)); names.append("shs_base")
    jobs.append((0, True, bool(args.motion_reg))); names.append("motion_base")            # :800-804
    outs = dict(zip(names, _TrainMLPs.apply(hexplane_feature, self.get_temporalpos.detach(), timestamp, tuple(jobs), images,
                                            *images.params())))

    if self.rotation_activation is not torch.nn.functional.normalize or self.scaling_activation is not torch.exp or \
            self.opacity_activation is not torch.sigmoid:
        raise UnsupportedDeformationConfig("the fused epilogue implements the reference's activations (normalize, exp, sigmoid; "
                                           "scene/saro_gaussian.py:39-47)")
    min_scale = self.args.min_interval / (self.duration)                                  # :783
    motion, rot, scale, opacity, lifespan, real_xyz = _TrainEpilogue.apply(
        timestamp, min_scale, outs["life"], outs["motion"], outs["rot"], outs["motion_base"], self._xyz, self._rotation,
        self._scaling, self._opacity, self.get_temporalpos)                               # :782-831 in one kernel
    self._lifespan = lifespan
    if args.scale_reg:
        self.scale_residual = outs["rot_base"][:, 4:]
    if args.shs_reg:
        self.shs_residual = outs["shs_base"].reshape(-1, 16, 3)
    if args.motion_reg:
        self.motion_residual = outs["motion_base"]
    self.real_xyz = real_xyz                                                              # :803-804 (no gradient)
    shs = torch.cat((self._features_dc, self._features_rest), dim=1) + outs["shs"].reshape(-1, 16, 3)   # :837-841
    return motion, rot, scale, opacity, shs


def _lifespan_native(self, hexplane_feature):
    """lifespan = (1 - min_scale) * (1 - opacity_mlp(feature)) + min_scale  (scene/saro_gaussian.py:782-784) with the MLP as
    one job of the tcgen05 forward kernel; differentiable (planes, opacity_mlp) when gradients are enabled."""
    cache = getattr(self, "_sgs_train_cache", None)
    key = tuple(id(getattr(self, k)) for k in _TRAIN_MLPS)
    if cache is None or cache["key"] != key:
        cache = {"key": key, "images": TrainImages(*[getattr(self, k) for k in _TRAIN_MLPS])}
        self._sgs_train_cache = cache
    images = cache["images"]
    (raw,) = _TrainMLPs.apply(hexplane_feature, self.get_temporalpos.detach(), 0.0, ((3, True, True),), images, *images.params())
    min_scale = self.args.min_interval / (self.duration)
    return (1 - min_scale) * (1 - torch.sigmoid(raw)) + min_scale


def get_deformfeature(self):
    """Drop-in for GaussianModel.get_deformfeature (scene/saro_gaussian.py:863-869), called once before the test-time
    render loop (test.py:199): caches `hexplane_feature` and `_lifespan` for get_deformation_eval."""
    if not self._xyz.is_cuda:
        raise RuntimeError("get_deformfeature: the model must live on a CUDA device (there is no CPU path)")
    self.hexplane_feature = self.hexplane(self._xyz.detach(), self.get_temporalpos.detach(), self.get_scaling.detach())
    self._lifespan = _lifespan_native(self, self.hexplane_feature)


def get_intergral(self, start=0.0, end=1.0):
    """Drop-in for GaussianModel.get_intergral (scene/saro_gaussian.py:761-777, Eq. 22 of the paper; used by the
    densification masks at :349 and :720): the plane sample and the lifespan MLP run natively under no_grad, the closed
    form Q is the reference's own statement."""
    import numpy as np
    if not self._xyz.is_cuda:
        raise RuntimeError("get_intergral: the model must live on a CUDA device (there is no CPU path)")
    with torch.no_grad():
        feature = self.hexplane(self._xyz.detach(), self.get_temporalpos.detach(), self.get_scaling.detach())
        lifespan = _lifespan_native(self, feature)

    def Q(x):
        a1 = torch.tensor([0.070565902], device=x.device)
        a2 = torch.tensor([1.5976], device=x.device)
        return 1 - 1 / (1 + torch.exp(a1 * x ** 3 + a2 * x))

    p1 = Q(2 * np.sqrt(2) * (end - self.get_temporalpos) / lifespan)
    p2 = Q(2 * np.sqrt(2) * (start - self.get_temporalpos) / lifespan)
    return lifespan * np.sqrt(np.pi) / 2 * (p1 - p2)
