"""Densification statistics of one training iteration: the consumer of the rasterizer's
`radii` and means2D gradient in the reference's training loop.

The reference keeps three Python lists per iteration (train.py:192-194), appends one tensor per view to each
(:211-215) and reduces them with stack/sum/max/boolean-index ops (:281-292).  `BatchDensifyStats` keeps three running
device buffers instead and does the same arithmetic in two streaming kernels behind the C ABI
(`sgs_densify_add_view`, `sgs_densify_commit`); in data-parallel training (one view per rank) the
buffers are all-reduced between the two.  No CPU / PyTorch fallback: CPU tensors raise.

    stats = BatchDensifyStats(P, device)                      # replaces the three lists
    for cam in batch:
        ... render, loss.backward() ...
        stats.add_view(viewspace_point_tensor.grad, radii)    # replaces train.py:211-215
        # (or, fused: stats.attach_next_backward() BEFORE loss.backward() — the rasterizer's backward does it)
    stats.all_reduce()                                        # data parallel only
    stats.commit(gaussians)                                   # replaces train.py:281-292
"""
import torch

from . import _lib


def _dev_f32(t, name, n, tail=()):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous() or t.shape[0] != n or tuple(t.shape[1:]) not in tail:
        raise RuntimeError(f"{name}: expected contiguous float32 of shape ({n}, {tail[0]}), got {t.dtype} {tuple(t.shape)}")
    return t


class BatchDensifyStats:
    def __init__(self, num_points, device):
        self.P = int(num_points)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchDensifyStats needs a CUDA device (there is no CPU path)")
        # one allocation, three [P] views: reset() is a single memset
        self._buffers = torch.zeros(3, self.P, dtype=torch.int32, device=self.device)
        self.grad_sum = self._buffers[0].view(torch.float32)
        self.vis_count = self._buffers[1]
        self.radii_max = self._buffers[2]
        self.views = 0

    def reset(self):
        self._buffers.zero_()
        self.views = 0

    def add_view(self, viewspace_point_grad, radii):
        """viewspace_point_grad: the [P, 3] gradient of the rasterizer's means2D input; radii: its int32 [P] output."""
        g = _dev_f32(viewspace_point_grad, "viewspace_point_grad", self.P, [(3,)])
        if not radii.is_cuda or radii.dtype != torch.int32 or radii.shape != (self.P,) or not radii.is_contiguous():
            raise RuntimeError(f"radii: expected contiguous CUDA int32 of shape ({self.P},), got {radii.dtype} {tuple(radii.shape)}")
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.sgs_densify_add_view(self.P, g.data_ptr(), radii.data_ptr(), self.grad_sum.data_ptr(),
                                          self.vis_count.data_ptr(), self.radii_max.data_ptr(),
                                          torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"sgs_densify_add_view failed ({rc}): {_lib.last_error()}")
        self.views += 1

    def attach_next_backward(self):
        """Fused form of add_view: the NEXT rasterizer backward (of self.P Gaussians on self.device) applies this view's
        update in the epilogue of its last kernel — no extra launch, the gradient and the radii are still in registers.
        Call it before loss.backward(); do not also call add_view for that view."""
        from . import backend
        backend.arm_densify_sink(self)

    def all_reduce(self, group=None):
        """Data-parallel training: combine the views of all ranks (NCCL over NVLink; three small collectives)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        reduce_running_buffers(self.grad_sum, self.vis_count, self.radii_max, group)

    def commit(self, gaussians):
        """Updates gaussians.max_radii2D [P], .xyz_gradient_accum [P,1], .denom [P,1] in place (train.py:290-291)."""
        mr = _dev_f32(gaussians.max_radii2D, "max_radii2D", self.P, [()])
        acc = _dev_f32(gaussians.xyz_gradient_accum, "xyz_gradient_accum", self.P, [(1,), ()])
        den = _dev_f32(gaussians.denom, "denom", self.P, [(1,), ()])
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.sgs_densify_commit(self.P, self.grad_sum.data_ptr(), self.vis_count.data_ptr(), self.radii_max.data_ptr(),
                                        mr.data_ptr(), acc.data_ptr(), den.data_ptr(),
                                        torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"sgs_densify_commit failed ({rc}): {_lib.last_error()}")


def reduce_running_buffers(grad_sum, vis_count, radii_max, group=None):
    """The exchange step of data-parallel densification statistics: SUM, SUM, MAX over ranks (any backend)."""
    import torch.distributed as dist
    dist.all_reduce(grad_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(vis_count, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(radii_max, op=dist.ReduceOp.MAX, group=group)
