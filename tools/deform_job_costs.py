"""Developer tool (GPU box): time the training-deformation forward / data-gradient kernels with ALL CTAs on one job
kind at a time — the per-tile cost ratios that drive the CTA split between jobs (csrc/sgs_deform.cu: job_cost)."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import deform_torch
from saro_gs_b200 import _lib, deformation as D

dev = torch.device("cuda:0")
lib = _lib.load()
N, F = 300_000, 32
g = torch.Generator().manual_seed(0)
feat = (torch.randn(N, F, generator=g) * 0.5).to(dev)
tpos = torch.rand(N, 1, generator=g).to(dev)
mlps = deform_torch.make_train_mlps(F, device=dev, seed=1)
images = D.TrainImages(mlps["motion"], mlps["rot"], mlps["shs"], mlps["opacity"])
images.refresh(images.params())
stream = torch.cuda.current_stream(dev).cuda_stream
planes = lambda G: torch.empty(lib.sgs_deform_planes_bytes(N, G), dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for m, name in ((3, "life"), (0, "motion"), (1, "rot"), (2, "shs")):
    n_out = images.shapes[m][2]
    out = torch.empty((N, n_out), device=dev)
    h1, h2, x, m1, m2 = planes(16), planes(16), planes(6), torch.empty((N, 4), dtype=torch.int32, device=dev), torch.empty((N, 4), dtype=torch.int32, device=dev)
    for save in (False, True):
        p0 = lambda t: t.data_ptr() if save else None
        job = (_lib.MLPJob * 1)(_lib.MLPJob(images.image(m, 0), None, out.data_ptr(), p0(h1), p0(h2), p0(x), p0(m1), p0(m2), n_out, 0))
        us = timed(lambda: lib.sgs_deform_train_forward(N, F, 0.4, tpos.data_ptr(), feat.data_ptr(), 1, job, stream))
        print(f"forward  {name:6s} save={int(save)}: {us:7.1f} us")
    dy = torch.randn(N, n_out, device=dev)
    slab = torch.empty((N, F), device=dev)
    dh2, dh1, dyp = planes(16), planes(16), planes(6 if n_out > 8 else 2)
    job = (_lib.MLPJob * 1)(_lib.MLPJob(images.image(m, 1), dy.data_ptr(), slab.data_ptr(), dh2.data_ptr(), dh1.data_ptr(), dyp.data_ptr(),
                                         m2.data_ptr(), m1.data_ptr(), n_out, 0))
    us = timed(lambda: lib.sgs_deform_train_backward(N, F, 1, job, stream))
    print(f"backward {name:6s}        : {us:7.1f} us")
