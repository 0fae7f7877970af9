"""Clone / split / prune with optimizer-state surgery (the part that changes the
number of Gaussians): host-side mirror of the reference's GaussianModel methods

    _prune_optimizer, prune_points              scene/saro_gaussian.py:555-597
    cat_tensors_to_optimizer, densification_postfix   :600-645
    densify_and_splitv2, densify_and_clone      :650-700
    densify_pruneclone                          :704-739
    replace_tensor_to_optimizer, reset_opacity  :451-454, 540-553

with the same names, arguments and side effects (attributes `_xyz`, `_features_dc`, `_features_rest`, `_opacity`,
`_scaling`, `_rotation`, `_temporal_pos`, the statistics buffers, and an Adam optimizer whose per-Gaussian parameter
groups are named "xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "temporal_pos").  `install(cls)` attaches
them to a GaussianModel class, so `helper_train.controlgaussians` (helper_train.py:103-174) runs unchanged.

Design difference.  The reference performs densify_pruneclone as three rounds of boolean-mask / torch.cat surgery
(clone -> cat; split -> cat -> prune the parents; prune), each round re-allocating the seven parameter tensors and
their two Adam moments (21 tensors, ~0.7 GB moved per round at 300 k Gaussians).  Here the three rounds are PLANNED
first — one source-row index per surviving row — and every tensor is gathered ONCE.  The plan reproduces the
reference's row order ([unsplit originals | clones | split children], then the final prune) and its random stream
(one torch.normal call with the reference's shapes), so the result is bit-identical (tests/test_surgery.py, against the
reference's own statements executed via `ast`).

Data-parallel note (one view per rank, sharding.py): every rank must take the same decisions.  The statistics are
all-reduced before this runs (densify.BatchDensifyStats.all_reduce); the split samples come from the device RNG, so
either seed all ranks identically and keep their RNG consumption in lockstep, or pass `broadcast_from=0` to
densify_pruneclone, which broadcasts the new positions of the split children from that rank.
"""
import numpy as np
import torch
from torch import nn

PER_GAUSSIAN_GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "temporal_pos")
_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
         "scaling": "_scaling", "rotation": "_rotation", "temporal_pos": "_temporal_pos"}


def build_rotation(r):
    """utils/general_utils.py:127-148 (rotation matrix of an un-normalised quaternion), on the tensor's own device."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device, dtype=r.dtype)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r_ * z)
    R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y)
    R[:, 2, 1] = 2 * (y * z + r_ * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def _groups(self):
    """the optimizer's per-Gaussian groups, as the reference selects them (:561)"""
    for group in self.optimizer.param_groups:
        if len(group["params"]) == 1 and "mlp" not in group["name"] and group["name"] != "hexplane":
            yield group


def _regather(self, rows, fresh=None, overrides=None):
    """Rebuild every per-Gaussian parameter (and its Adam moments) as `tensor[rows]`.
    fresh: bool [len(rows)] — rows that are NEW Gaussians (clones / split children): their Adam moments are zero,
    as torch.zeros_like(extension_tensor) gives them in cat_tensors_to_optimizer.
    overrides: {group name: (row mask, values)} — rows whose value is not a copy of the source row."""
    out = {}
    for group in _groups(self):
        name = group["name"]
        if name not in _ATTR:
            continue
        old = group["params"][0]
        new = old.detach()[rows]
        if overrides and name in overrides:
            mask, values = overrides[name]
            new[mask] = values
        state = self.optimizer.state.get(old, None)
        if state is not None:
            for key in ("exp_avg", "exp_avg_sq"):
                m = state[key][rows]
                if fresh is not None:
                    m[fresh] = 0
                state[key] = m
            del self.optimizer.state[old]
        group["params"][0] = nn.Parameter(new.requires_grad_(True))
        if state is not None:
            self.optimizer.state[group["params"][0]] = state
        out[name] = group["params"][0]
    for name, p in out.items():
        setattr(self, _ATTR[name], p)
    return out


def prune_points(self, mask):
    """:586-597 — drop the rows where `mask` is True from every per-Gaussian tensor, Adam moment and statistic."""
    valid = ~mask
    rows = torch.nonzero(valid).squeeze(1)
    _regather(self, rows)
    self.xyz_gradient_accum = self.xyz_gradient_accum[valid]
    self.t_gradient_accum = self.t_gradient_accum[valid]
    self.denom = self.denom[valid]
    self.max_radii2D = self.max_radii2D[valid]


def _prune_optimizer(self, mask):
    """:555-584 (mask = rows to KEEP)"""
    return _regather(self, torch.nonzero(mask).squeeze(1))


def cat_tensors_to_optimizer(self, tensors_dict):
    """:600-620"""
    optimizable = {}
    for group in self.optimizer.param_groups:
        if len(group["params"]) == 1 and group["name"] in tensors_dict:
            ext = tensors_dict[group["name"]]
            old = group["params"][0]
            state = self.optimizer.state.get(old, None)
            if state is not None:
                state["exp_avg"] = torch.cat((state["exp_avg"], torch.zeros_like(ext)), dim=0)
                state["exp_avg_sq"] = torch.cat((state["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
                del self.optimizer.state[old]
            group["params"][0] = nn.Parameter(torch.cat((old.detach(), ext), dim=0).requires_grad_(True))
            if state is not None:
                self.optimizer.state[group["params"][0]] = state
            optimizable[group["name"]] = group["params"][0]
    return optimizable


def _reset_statistics(self, n):
    dev = self._xyz.device
    self.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
    self.denom = torch.zeros((n, 1), device=dev)
    self.max_radii2D = torch.zeros((n,), device=dev)
    self.t_gradient_accum = torch.zeros((n, 1), device=dev)


def densification_postfix(self, new_xyz, new_features_dc, new_feature_rest, new_opacities, new_scaling, new_rotation,
                          new_temporal_pos, dummy=None):
    """:622-645"""
    d = {"xyz": new_xyz, "f_dc": new_features_dc, "f_rest": new_feature_rest, "opacity": new_opacities,
         "scaling": new_scaling, "rotation": new_rotation, "temporal_pos": new_temporal_pos}
    opt = cat_tensors_to_optimizer(self, d)
    for name, p in opt.items():
        setattr(self, _ATTR[name], p)
    _reset_statistics(self, self._xyz.shape[0])


def _clone_mask(self, grads, grad_threshold, scene_extent):
    sel = torch.where(torch.norm(grads, dim=-1) >= grad_threshold, True, False)
    return torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values <= self.percent_dense * scene_extent)


def _split_mask(self, grads, grad_threshold, scene_extent, n_now):
    padded = torch.zeros((n_now,), device=self._xyz.device)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = torch.where(padded >= grad_threshold, True, False)
    return torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values > self.percent_dense * scene_extent)


def _split_children(self, sel, N):
    """positions / log-scales of the N children of every selected Gaussian (:660-666): one torch.normal call with the
    reference's shapes, so the device RNG stream is consumed exactly as the reference consumes it."""
    stds = self.get_scaling[sel].repeat(N, 1)
    means = torch.zeros((stds.size(0), 3), device=stds.device)
    samples = torch.normal(mean=means, std=stds)
    rots = build_rotation(self._rotation[sel]).repeat(N, 1, 1)
    new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.get_xyz[sel].repeat(N, 1)
    new_scaling = self.scaling_inverse_activation(self.get_scaling[sel].repeat(N, 1) / (0.8 * N))
    return new_xyz, new_scaling


def densify_and_clone(self, grads, grad_threshold, scene_extent, t_grads=None):
    """:686-700"""
    sel = _clone_mask(self, grads, grad_threshold, scene_extent)
    P = self._xyz.shape[0]
    rows = torch.cat((torch.arange(P, device=sel.device), torch.nonzero(sel).squeeze(1)))
    fresh = torch.zeros(rows.shape[0], dtype=torch.bool, device=sel.device)
    fresh[P:] = True
    _regather(self, rows, fresh)
    _reset_statistics(self, rows.shape[0])


def densify_and_splitv2(self, grads, grad_threshold, scene_extent, N=2, t_grads=None):
    """:650-683"""
    with torch.no_grad():
        n0 = self._xyz.shape[0]
        sel = _split_mask(self, grads, grad_threshold, scene_extent, n0)
        new_xyz, new_scaling = _split_children(self, sel, N)
        parents = torch.nonzero(sel).squeeze(1)
        keep = torch.nonzero(~sel).squeeze(1)
        rows = torch.cat((keep, parents.repeat(N)))
        fresh = torch.zeros(rows.shape[0], dtype=torch.bool, device=sel.device)
        fresh[keep.shape[0]:] = True
        _regather(self, rows, fresh, {"xyz": (fresh, new_xyz), "scaling": (fresh, new_scaling)})
        _reset_statistics(self, rows.shape[0])


def densify_pruneclone(self, max_grad, min_opacity, extent, max_screen_size, splitN=1, broadcast_from=None):
    """:704-739 — clone, split (N = 2), prune, planned as ONE gather per tensor."""
    with torch.no_grad():
        dev = self._xyz.device
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        grads = grads * self.inv_intergral_fordensify
        P = self._xyz.shape[0]
        N = 2

        # ---- plan: rows of the reference's intermediate set [unsplit originals | clones | split children]
        clone_sel = _clone_mask(self, grads, max_grad, extent)
        split_sel = _split_mask(self, grads, max_grad, extent, P)    # clones carry a zero gradient: never selected
        new_xyz, new_scaling = _split_children(self, split_sel, N)
        if broadcast_from is not None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.broadcast(new_xyz, src=broadcast_from)
        keep = torch.nonzero(~split_sel).squeeze(1)
        clones = torch.nonzero(clone_sel).squeeze(1)
        parents = torch.nonzero(split_sel).squeeze(1)
        rows = torch.cat((keep, clones, parents.repeat(N)))
        n_keep, n_clone = keep.shape[0], clones.shape[0]
        fresh = torch.zeros(rows.shape[0], dtype=torch.bool, device=dev)
        fresh[n_keep:] = True
        child = torch.zeros_like(fresh)
        child[n_keep + n_clone:] = True

        # ---- the prune decision needs the intermediate values of four tensors only
        xyz_mid = self._xyz.detach()[rows]
        xyz_mid[child] = new_xyz
        scaling_mid = self._scaling.detach()[rows]
        scaling_mid[child] = new_scaling
        mid = _MidView(self, xyz_mid, scaling_mid, self._opacity.detach()[rows], self._temporal_pos.detach()[rows])
        prune_mask = (mid.get_opacity < min_opacity).squeeze()
        intergral_mask = (mid.get_intergral() < self.min_intergral).squeeze()
        prune_mask = torch.logical_or(prune_mask, intergral_mask)
        if self.args.loader == "colmap":
            prune_mask = torch.logical_or(prune_mask, (xyz_mid[:, 2] < 4.5).squeeze())
        if max_screen_size:
            # densification_postfix has just zeroed max_radii2D (:641), so the screen-size test never fires here;
            # kept for fidelity
            big_points_vs = torch.zeros(rows.shape[0], device=dev) > max_screen_size
            big_points_ws = mid.get_scaling.max(dim=1).values > 0.1 * extent
            if self.args.pw:
                prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_points_vs), big_points_ws)
            else:
                prune_mask = torch.logical_or(prune_mask, big_points_vs)

        # ---- one gather per tensor
        survive = ~prune_mask
        _regather(self, rows[survive], fresh[survive],
                  {"xyz": (child[survive], new_xyz[survive[n_keep + n_clone:]]),
                   "scaling": (child[survive], new_scaling[survive[n_keep + n_clone:]])})
        _reset_statistics(self, int(survive.sum()))
    if self._xyz.is_cuda:
        torch.cuda.empty_cache()


class _MidView:
    """The model as the reference sees it between the split and the final prune, for the four tensors the prune
    decision reads (get_opacity, get_scaling, get_temporalpos via get_intergral, _xyz)."""

    def __init__(self, model, xyz, scaling, opacity, temporal_pos):
        self._m = model
        self._xyz, self._scaling, self._opacity, self._temporal_pos = xyz, scaling, opacity, temporal_pos

    def __getattr__(self, name):
        return getattr(self._m, name)

    @property
    def get_opacity(self):
        return self._m.opacity_activation(self._opacity)

    @property
    def get_scaling(self):
        return self._m.scaling_activation(self._scaling)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_temporalpos(self):
        return torch.sigmoid(self._temporal_pos) if self._m.args.sigmoid_tcenter else self._temporal_pos

    def get_intergral(self, start=0.0, end=1.0):
        return type(self._m).get_intergral(self, start, end)


def replace_tensor_to_optimizer(self, tensor, name):
    """:540-553"""
    optimizable = {}
    for group in self.optimizer.param_groups:
        if group["name"] == name:
            old = group["params"][0]
            state = self.optimizer.state.get(old, None)
            state["exp_avg"] = torch.zeros_like(tensor)
            state["exp_avg_sq"] = torch.zeros_like(tensor)
            del self.optimizer.state[old]
            group["params"][0] = nn.Parameter(tensor.requires_grad_(True))
            self.optimizer.state[group["params"][0]] = state
            optimizable[group["name"]] = group["params"][0]
    return optimizable


def reset_opacity(self):
    """:451-454"""
    opacities_new = inverse_sigmoid(torch.min(self.get_opacity, torch.ones_like(self.get_opacity) * 0.01))
    self._opacity = replace_tensor_to_optimizer(self, opacities_new, "opacity")["opacity"]


def get_intergral(self, start=0.0, end=1.0):
    """:761-777 (Eq. 22 of the paper) on the tensors' own device."""
    with torch.no_grad():
        hexplane_feature = self.hexplane(self._xyz.detach(), self.get_temporalpos.detach(), self.get_scaling.detach())
        lifespan = 1 - self.opacity_mlp(hexplane_feature.clone())
        min_scale = self.args.min_interval / (self.duration)
        lifespan = (1 - min_scale) * lifespan + min_scale
    dev = self._xyz.device

    def Q(x):
        a1 = torch.tensor([0.070565902], device=dev)
        a2 = torch.tensor([1.5976], device=dev)
        return 1 - 1 / (1 + torch.exp(a1 * x ** 3 + a2 * x))
    p1 = Q(2 * np.sqrt(2) * (end - self.get_temporalpos) / lifespan)
    p2 = Q(2 * np.sqrt(2) * (start - self.get_temporalpos) / lifespan)
    return lifespan * np.sqrt(np.pi) / 2 * (p1 - p2)


METHODS = {f.__name__: f for f in (prune_points, _prune_optimizer, cat_tensors_to_optimizer, densification_postfix,
                                   densify_and_clone, densify_and_splitv2, densify_pruneclone,
                                   replace_tensor_to_optimizer, reset_opacity)}


def install(cls, with_intergral=False):
    """Attach the surgery methods to a GaussianModel class (the reference's scene.saro_gaussian.GaussianModel or a
    stand-in with the same attributes)."""
    for name, f in METHODS.items():
        setattr(cls, name, f)
    if with_intergral:
        cls.get_intergral = get_intergral
    return cls
