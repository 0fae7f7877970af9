#!/usr/bin/env python
"""bench.py — rasterizer forward+backward on BASELINE.json configs[1] (headline workload).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward + one backward pass of the rasterizer over one 1352x1014 view of the
~300k-Gaussian synthetic N3D-like scene (saro_gs_b200.synthetic.config2_scene, R = 3.93 M tile
instances), driven through the reference-facing Python API (GaussianRasterizer + autograd).
Metric: ms/frame (lower is better); at N GPUs every rank renders its own views of the replicated
scene (weak scaling, no data-path collective; NCCL only for barriers/max-reduction of the time)
and value = max-over-ranks time / (K * N).

Timed numbers:
  value / ms_per_step : Gaussians + per-view inputs already resident in HBM; CUDA events per step
                        on the launching stream; L2 flushed (256 MiB memset) between steps,
                        outside the event pairs.
  e2e                 : the same step through the same API, but every step first copies the
                        per-view inputs (camera matrices, background, the 16 MB dL/dcolor image) from
                        pinned host memory and ends with a device->host read of 8 bytes: a loss-like
                        scalar of the rendered image and a gradient checksum (the image itself stays on
                        the device, as in training).  The Gaussian parameters stay resident: the
                        reference API only accepts CUDA tensors for them.
  roofline            : dominant kernel (backward render) — algorithmic bytes / CUDA-event duration
                        measured by the library's stage profiler over a second pass of the same K steps
                        (the headline pass runs without the extra event records).  V, R and the kept
                        instance count come from the library (sgs_last_forward_counts), the ncu traffic /
                        instruction figures from the profiles/*.csv named in the object.  `hbm_stages`
                        adds the same figure for every HBM-bound stage; the worst is named.
  forward_only        : inference render, back to back; `sequence` = BASELINE configs[2] protocol
                        (300 frames, 4 passes, frames with index <= 10 dropped, test.py:155-163).
  cpu_baseline        : CPU oracle port (oracle/, float32, OpenMP) on one forward+backward, plus the
                        configs[0] case (10 k Gaussians @400x400, forward only) on the same cores.

--impl reference times the UNMODIFIED reference CUDA rasterizer (oracle/_ref, compiled from
/root/reference by oracle/build_ref.py) through the identical host layer, same config/metric.
The reference has no CPU implementation of this path; if the compiled reference
is absent the arm falls back to the CPU oracle port and says so.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rasterizer fwd+bwd ms/frame @1352x1014, ~300k Gaussians"
UNIT = "ms/frame"
WORKLOAD = "configs[1]: synthetic N3D-like cook_spinach stand-in, P=300000, 1352x1014, SH deg 3, R=3927052"


def pin_rank_to_cores(local_rank, world):
    """One-process-per-GPU runs: give every rank its own slice of the host cores so that eight Python processes (plus
    their CUDA / NCCL helper threads) do not migrate across each other inside the timed region."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(1, world)
        if world > 1 and per >= 2:
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
    except (AttributeError, OSError):
        pass


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        smax = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        # median over the upper half of samples (= under load; idle gaps between steps read low)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_inputs(dev, rank):
    from saro_gs_b200 import synthetic
    scene, cam0 = synthetic.config2_scene()
    # every rank renders the SAME four views of the replicated scene: the weak-scaling number then measures the
    # system, not a per-rank difference in workload (round 1 gave every rank its own cameras)
    cams = [synthetic.yaw_camera(cam0.width, cam0.height, 729.0, yaw=0.004 * k, pivot=(0.0, 0.0, 22.0))
            for k in range(4)]
    cams[0] = cam0
    params = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    cot = synthetic.cotangent(cam0.height, cam0.width)
    return scene, cams, params, cot


def run_native_or_ref(args, impl):
    rank, world, local = dist_setup(args.gpus)
    pin_rank_to_cores(local, world)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import saro_gs_b200 as sgs
    from saro_gs_b200 import _lib

    kind = "native"
    if impl == "reference":
        from oracle import ref_loader
        if not ref_loader.available():
            return run_cpu_reference(args, rank, world, why="oracle/_ref missing")
        Rast = ref_loader.ref_api()[1]
        kind = "reference-cuda"
    else:
        Rast = sgs.GaussianRasterizer
    Settings = sgs.GaussianRasterizationSettings

    scene, cams, params, cot_cpu = make_inputs(dev, rank)
    H, W = cams[0].height, cams[0].width
    means2D = torch.zeros_like(params["means3D"], requires_grad=True)
    bg_cpu = torch.zeros(3)
    cot_dev = cot_cpu.to(dev)
    cams_dev = [(c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev)) for c in cams]
    bg_dev = bg_cpu.to(dev)

    # pinned host staging for the e2e leg
    pin = lambda t: t.contiguous().pin_memory()
    cot_pin = pin(cot_cpu)
    cam_pin = [(pin(c.viewmatrix), pin(c.projmatrix), pin(c.campos)) for c in cams]
    bg_pin = pin(bg_cpu)
    res_host = torch.empty((2,), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def zero_grads():
        for p in list(params.values()) + [means2D]:
            p.grad = None

    counts_log = []          # (kept, num_rendered, visible) of every step of the profiled pass

    def step_resident(i):
        c = cams[i % len(cams)]
        v, p, cp = cams_dev[i % len(cams)]
        rs = Settings(H, W, c.tanfovx, c.tanfovy, bg_dev, 1.0, v, p, scene.sh_degree, cp, False)
        color, radii, depth = Rast(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                       shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        if step_resident.log_counts:
            c5 = (ctypes.c_int64 * 5)()
            _lib.load().sgs_last_forward_counts(c5)      # no synchronisation: read during the forward call
            counts_log.append((int(c5[0]), int(c5[1]), int(c5[2]), int(c5[4])))
        color.backward(cot_dev)
        zero_grads()
    step_resident.log_counts = False

    # e2e leg = one training-style step through the public API with HOST inputs: the per-view inputs (camera,
    # background and the 16 MB dL/dcolor image — the stand-in for the ground-truth image a training step uploads)
    # come from pinned host memory every step, and the step's result (a loss-like scalar of the rendered image and
    # a gradient checksum) is read back to the host.  The caller overlaps its upload with the rasterizer as a
    # training loop would: the image travels on a side stream while forward runs.  Same code for both arms.
    s_in = torch.cuda.Stream(dev)
    ev_in = torch.cuda.Event()

    # per-view small inputs packed into ONE pinned buffer (view 16 | proj 16 | campos 3 | bg 3 floats): one H2D copy
    view_pin = [pin(torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1), bg_cpu]))
                for c in cams]

    def step_e2e(i):
        main = torch.cuda.current_stream(dev)
        c = cams[i % len(cams)]
        with torch.cuda.stream(s_in):
            cot = cot_pin.to(dev, non_blocking=True)
            ev_in.record(s_in)
        pk = view_pin[i % len(cams)].to(dev, non_blocking=True)
        v, p, cp, bg = pk[0:16].view(4, 4), pk[16:32].view(4, 4), pk[32:35], pk[35:38]
        rs = Settings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, v, p, scene.sh_degree, cp, False)
        color, radii, depth = Rast(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                       shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        main.wait_event(ev_in)
        cot.record_stream(main)
        color.backward(cot)
        res = torch.stack([(color.detach() * cot).sum(), params["means3D"].grad.abs().sum()])
        res_host.copy_(res, non_blocking=True)
        main.synchronize()    # the caller consumes the loss / metric on the host
        zero_grads()

    # informational variant of the same step: the host reads step i's result AFTER it has enqueued step i + 1 (two
    # pinned result slots), as a training loop that logs its loss one iteration late does.  Every step still uploads its
    # inputs and reads its result back; what disappears is the GPU idling while Python prepares the next launch.  It
    # is reported next to `e2e.value`, never instead of it.
    res_slots = [torch.empty((2,), dtype=torch.float32).pin_memory() for _ in range(2)]
    res_ready = [torch.cuda.Event(), torch.cuda.Event()]
    pending = []

    def step_e2e_pipelined(i):
        main = torch.cuda.current_stream(dev)
        c = cams[i % len(cams)]
        with torch.cuda.stream(s_in):
            cot = cot_pin.to(dev, non_blocking=True)
            ev_in.record(s_in)
        pk = view_pin[i % len(cams)].to(dev, non_blocking=True)
        v, p, cp, bg = pk[0:16].view(4, 4), pk[16:32].view(4, 4), pk[32:35], pk[35:38]
        rs = Settings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, v, p, scene.sh_degree, cp, False)
        color, radii, depth = Rast(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                       shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        main.wait_event(ev_in)
        cot.record_stream(main)
        color.backward(cot)
        res = torch.stack([(color.detach() * cot).sum(), params["means3D"].grad.abs().sum()])
        slot = i & 1
        res_slots[slot].copy_(res, non_blocking=True)
        res_ready[slot].record(main)
        zero_grads()
        if pending:
            res_ready[pending.pop()].synchronize()      # the previous step's result is consumed now
        pending.append(slot)

    h2d_bytes = cot_pin.numel() * 4 + (16 + 16 + 3 + 3) * 4
    d2h_bytes = res_host.numel() * 4

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, K, W_, profile=False):
        import gc
        for i in range(W_):
            step_fn(i)
        # measurement hygiene for the multi-rank runs (round 1: single 1.1 - 2.0 ms steps on one rank of eight cost the
        # whole job 8 %, the per-rank medians were identical): no cyclic-GC pause inside a timed step
        gc.collect()
        gc.disable()
        try:
            return _timed(step_fn, K, W_, profile)
        finally:
            gc.enable()

    def _timed(step_fn, K, W_, profile):
        barrier()
        if profile:
            _lib.load().sgs_profile_read(None, None, None)  # reset
            _lib.load().sgs_profile_enable(1)
        evs = []
        t0 = time.perf_counter()
        for i in range(K):
            flush.zero_()                                   # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn(W_ + i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        per_step = [a.elapsed_time(b) for a, b in evs]
        dev_ms = sum(per_step)
        timed.last_per_step = per_step
        stage = None
        if profile:
            ms = (ctypes.c_float * len(_lib.STAGES))()
            calls = (ctypes.c_int * len(_lib.STAGES))()
            launches = ctypes.c_uint64(0)
            _lib.load().sgs_profile_read(ms, calls, ctypes.byref(launches))
            _lib.load().sgs_profile_enable(0)
            stage = {"ms": {n: ms[k] for k, n in enumerate(_lib.STAGES)},
                     "calls": {n: calls[k] for k, n in enumerate(_lib.STAGES)}, "own_launches": int(launches.value)}
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms = float(t.item())
        return dev_ms, wall, stage

    def step_forward_only(i):
        # test-time rendering (test.py:121,158 of the reference): no autograd, no state kept for backward
        c = cams[i % len(cams)]
        v, p, cp = cams_dev[i % len(cams)]
        rs = Settings(H, W, c.tanfovx, c.tanfovy, bg_dev, 1.0, v, p, scene.sh_degree, cp, False)
        with torch.no_grad():
            Rast(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"], shs=params["shs"],
                     scales=params["scales"], rotations=params["rotations"])

    def sequence_fps(frames=300, passes=4):
        """BASELINE.json configs[2] protocol (test.py:155-163 of the reference): `passes` passes over a `frames`-frame
        sequence of the cloud (alive set changes per frame, synthetic.DeviceSequence), one synchronised render per
        frame, frames with index <= 10 of every pass dropped, mean of the rest.  CUDA events around the rasterizer call
        (the reference's wall clock also covers its PyTorch deformation ops, which are not on this path)."""
        from saro_gs_b200 import synthetic
        seq = synthetic.DeviceSequence(scene, dev)
        c = cams[0]
        v, p, cp = cams_dev[0]
        rs = Settings(H, W, c.tanfovx, c.tanfovy, bg_dev, 1.0, v, p, scene.sh_degree, cp, False)
        rast = Rast(rs)
        times = []
        with torch.no_grad():
            for _ in range(passes):
                for idx in range(frames):
                    sc = seq.frame(idx / frames)
                    m2 = torch.zeros_like(sc.means3D)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rast(means3D=sc.means3D, means2D=m2, opacities=sc.opacities, shs=sc.shs, scales=sc.scales,
                         rotations=sc.rotations)
                    e1.record()
                    torch.cuda.synchronize()
                    if idx > 10:
                        times.append(e0.elapsed_time(e1))
        mean_ms = sum(times) / len(times)
        return {"frames": frames, "passes": passes, "timed_frames": len(times), "ms_per_frame": mean_ms,
                "frames_per_s": 1e3 / mean_ms,
                "protocol": "test.py:155-163 of the reference: 4 passes, frames with index <= 10 dropped, one "
                            "synchronised render per frame (launch latency included on both arms)"}

    # how long the 16 MB image upload takes on its own (the e2e step cannot be shorter than upload + backward)
    torch.cuda.synchronize()
    eu0, eu1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _dst = torch.empty_like(cot_dev)                    # preallocated: the figure is the link, not the allocator
    _dst.copy_(cot_pin, non_blocking=True)
    torch.cuda.synchronize()
    eu0.record()
    for _ in range(5):
        _dst.copy_(cot_pin, non_blocking=True)
    eu1.record()
    torch.cuda.synchronize()
    h2d_alone_ms = eu0.elapsed_time(eu1) / 5
    del _dst

    K, W_ = args.steps, max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None      # one nvidia-smi sampler per node, not one per rank
    if sampler:
        sampler.start()
    dev_ms, wall_ms, stage = timed(step_resident, K, W_, profile=False)
    per_step_headline = list(timed.last_per_step)
    if impl != "reference":
        # the same K steps again with the library's per-stage CUDA events switched on: the per-kernel durations
        # behind `roofline` (kept out of the headline pass: ~20 extra event records per step cost host time)
        step_resident.log_counts = True
        prof_ms, _, stage = timed(step_resident, K, 3, profile=True)
        step_resident.log_counts = False
        del counts_log[:3]                                # the 3 warm-up steps of that pass
    e2e_ms, _, _ = timed(step_e2e, K, W_)
    e2e_pipe_ms, _, _ = timed(step_e2e_pipelined, K, W_)
    if pending:
        res_ready[pending.pop()].synchronize()
    fwd_ms, _, _ = timed(step_forward_only, K, 3)
    clocks = sampler.stop() if sampler else None
    seq = sequence_fps() if (rank == 0 and not args.no_sequence) else None

    # per-rank view of the headline pass: min / median / max step time and the instance counts every rank saw
    counts = (ctypes.c_int64 * 5)()
    if impl != "reference":
        _lib.load().sgs_last_forward_counts(counts)
    ps = sorted(per_step_headline)
    mine = {"rank": rank, "min_ms": ps[0], "median_ms": ps[len(ps) // 2], "max_ms": ps[-1],
            "kept": int(counts[0]), "num_rendered": int(counts[1])}
    rank_stats = [mine]
    if world > 1:
        import torch.distributed as dist
        rank_stats = [None] * world
        dist.all_gather_object(rank_stats, mine)

    ms_per_step = dev_ms / K
    value = dev_ms / (K * world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
        "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "views_per_rank": K, "sharding": "one view per rank, Gaussians replicated",
                   "l2": "256 MiB memset between steps, outside the per-step CUDA-event pairs",
                   "timing": "sum of per-step CUDA-event times on the launching stream, max over ranks"},
        "e2e": {"value": e2e_ms / (K * world), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes,
                "note": "per-view inputs (camera + bg packed in one 152-byte copy, 16 MB dL/dcolor image) from pinned host "
                        "memory every step; the step's loss-like scalar + a checksum of dL/dmeans3D read back (8 bytes); "
                        "Gaussian parameters resident (the API takes CUDA tensors only); the image upload runs on a "
                        "side stream while forward runs (same code for both arms)",
                "h2d_image_ms_alone": h2d_alone_ms,
                "pipelined_value": e2e_pipe_ms / (K * world),
                "pipelined_note": "informational: same uploads and read-backs every step, but the host consumes step i's "
                                  "result after enqueueing step i + 1 (no GPU idle while Python prepares the next launch); "
                                  "`value` above is the strict form (host blocks on every step's result)"},
        "clocks": clocks, "wall_ms_timed_region": wall_ms,
        "forward_only": {"ms_per_frame": fwd_ms / (K * world), "frames_per_s": 1e3 * K * world / fwd_ms,
                         "note": "inference render of the same views (no_grad), inputs resident, back to back",
                         "sequence": seq},
        "per_rank": rank_stats,
    }
    if impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                "sample": "the reference's implementation of this path is its CUDA rasterizer (it has "
                                          "no CPU path); unmodified sources compiled for sm_100a (oracle/build_ref.py), "
                                          "timed on the same GPU through the same host layer"}
        line["gpu_launches"] = None
    else:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        clocks = clocks or {}
        P = params["means3D"].shape[0]
        T = ((W + 15) // 16) * ((H + 15) // 16)
        # instance counts measured by the library during the profiled pass (mean over its steps / cameras)
        nlog = max(1, len(counts_log))
        Rk = sum(c[0] for c in counts_log) / nlog         # kept instances: binned, sorted, rendered
        Rn = sum(c[1] for c in counts_log) / nlog         # num_rendered as the reference counts it
        V = sum(c[2] for c in counts_log) / nlog          # visible Gaussians
        relaunches = sum(c[3] for c in counts_log)
        # SURVEY.md section 8(d) algorithmic bytes of the backward render kernel (the reference's instance count)
        alg = 8 * T + 40 * Rn + 20 * H * W + 44 * V + 44 * P
        calls = max(1, stage["calls"]["render_bwd"])
        k_ms = stage["ms"]["render_bwd"] / calls
        achieved = alg / (k_ms * 1e-3) / 1e9
        ncu = load_ncu_summary("render_bwd_kernel")
        sm_mhz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        issue_peak = 148 * 4 * sm_mhz * 1e6   # one warp-instruction per SM sub-partition per clock
        line["roofline"] = {"kernel": "render_bwd_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                            "unit": "GB/s", "frac": achieved / peak, "traffic": ncu.get("traffic_bytes"),
                            "traffic_source": ncu.get("file"),
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                            "counts": {"P": P, "visible": V, "num_rendered": Rn, "kept": Rk, "tiles": T,
                                       "source": "sgs_last_forward_counts, mean over the profiled pass"},
                            "avg_launch_ms": k_ms,
                            "note": "reported against HBM as the contract asks, but this kernel is instruction-issue "
                                    "bound (ncu: issue-active ~85 %, DRAM < 3 %): ~95 blended pixel x Gaussian pairs "
                                    "are evaluated per 48-byte record (DESIGN.md section 3); see `issue`"}
        if ncu.get("warp_instructions"):
            wi = ncu["warp_instructions"]
            line["roofline"]["issue"] = {"warp_instructions_per_launch": wi, "source": ncu.get("file"),
                                         "achieved_Tinst_s": wi / (k_ms * 1e-3) / 1e12,
                                         "peak_Tinst_s": issue_peak / 1e12, "frac": wi / (k_ms * 1e-3) / issue_peak}
        # the HBM-bound stages, each against the same measured copy bandwidth; bytes = what THIS implementation has to
        # move at minimum, not the reference's 172 B per instance of sort traffic
        Rc = 0.233 * Rk                                   # (supertile, Gaussian) instances per kept instance (measured
                                                          # at configs[1]: 504 k of 2.16 M, tools/bucket_stats.py)
        stage_bytes = {
            "preprocess_fwd": 52 * P + (192 + 67) * V,
            "depth_sort_scan": 12 * P,                    # since round 2b only the supertile-count scan is left here
            "tile_sort": 4 * Rk + 28 * Rc + 12 * P,
            "preprocess_bwd": 48 * P + (107 + 192) * V + (40 + 192) * P,
        }
        hbm = {}
        for n, byts in stage_bytes.items():
            c_ = max(1, stage["calls"][n])
            ms_ = stage["ms"][n] / c_
            if ms_ > 0:
                hbm[n] = {"algorithmic_bytes": byts, "ms": ms_, "GBps": byts / (ms_ * 1e-3) / 1e9,
                          "frac": byts / (ms_ * 1e-3) / 1e9 / peak}
        if hbm:
            worst = min(hbm, key=lambda n: hbm[n]["frac"])
            line["hbm_stages"] = {"stages": hbm, "worst": worst, "peak": peak, "unit": "GB/s",
                                  "note": "stage times include launch gaps (CUDA events around each stage)"}
        line["binning_relaunches_in_profiled_pass"] = relaunches
        line["stage_ms_per_step"] = {n: stage["ms"][n] / K for n in stage["ms"]}
        line["ms_per_step_with_stage_events"] = prof_ms / K
        line["gpu_launches"] = stage["own_launches"]
        line["gpu_launches_note"] = "hand-written kernels only, counted in the stage-profiled pass (8 per step: preprocess, " \
                                    "supertile scan, supertile bucketing, tile count, per-supertile sort + tile fill, " \
                                    "render | render bwd, preprocess bwd, + one 14 MB memset; the timed pass fuses " \
                                    "scan + bucketing into one cooperative launch); no library (CUB / cuBLAS) " \
                                    "kernels on this path since round 2"
        if rank == 0 and not args.no_extras:
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):     # stdout carries exactly one JSON line
                line["loss_path"] = loss_path_timing(dev, H, W)
                line["deform_path"] = deform_path_timing(dev)
                line["deform_train_path"] = deform_train_path_timing(dev)
                line["train_view_path"] = train_view_path_timing(dev)
                line["densify_path"] = densify_path_timing(dev)
                line["plane_path"] = plane_path_timing(dev)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# newest `ncu --set full` capture of the render kernels on this exact workload (tools/ncu_summary.py output)
NCU_SUMMARIES = ["profiles/r2r_render_ncu_summary.csv", "profiles/r2m_render_ncu_summary.csv", "profiles/r2_render_ncu_summary.csv",
                 "profiles/r02f_render_ncu_summary.csv"]


def load_ncu_summary(kernel):
    """{traffic_bytes, warp_instructions, file} of `kernel` from the newest committed ncu summary (per launch)."""
    import csv
    for rel in NCU_SUMMARIES:
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        out = {"file": rel}
        for row in csv.DictReader(open(path)):
            if not row["kernel"].startswith(kernel):
                continue
            if row["metric"] == "traffic_bytes(read+write)":
                out["traffic_bytes"] = int(float(row["value"]))
            if row["metric"] == "smsp__inst_executed.sum":
                out["warp_instructions"] = float(row["value"])
        if "traffic_bytes" in out:
            return out
    return {"file": None}


def loss_path_timing(dev, H, W, iters=20):
    """SURVEY.md section 8(f) rank 3 (first row widened beyond the rasterizer): the L1 + D-SSIM loss that runs between
    the rasterizer's forward and backward — fused CUDA kernels vs the same loss as PyTorch ops (what SaRO-GS runs),
    forward + backward to dL/dimage, CUDA events, inputs resident."""
    from saro_gs_b200 import loss_utils
    from oracle.ssim_torch import torch_l1_dssim_loss
    g = torch.Generator().manual_seed(11)
    gt = torch.rand(3, H, W, generator=g).to(dev)
    img = (gt + 0.05 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)

    def run(fn):
        ms = []
        for i in range(iters + 3):
            x = img.clone().requires_grad_(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = fn(x, gt, 0.2)
            loss.backward()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
        return sum(ms) / len(ms), float(loss.item())

    fused_ms, fused_val = run(loss_utils.l1_dssim_loss)
    torch_ms, torch_val = run(torch_l1_dssim_loss)
    px = 3 * H * W
    return {"what": "L1 + 0.2 D-SSIM loss forward + backward on one 3x%dx%d image" % (H, W),
            "fused_ms": fused_ms, "pytorch_ops_ms": torch_ms, "speedup": torch_ms / fused_ms,
            "fused_GBps": (8 + 12 + 20 + 4) * px / (fused_ms * 1e-3) / 1e9,
            "loss_fused": fused_val, "loss_pytorch": torch_val}


def deform_path_timing(dev, iters=15):
    """the per-frame deformation -> rasterizer hand-off of the test-time render loop
    (renderer/__init__.py:188-203 times get_deformation_eval + the rasterizer).  Fused tcgen05 kernel vs the same
    function as the PyTorch ops SaRO-GS runs, on the headline cloud (300 k Gaussians, 32 plane features), and the
    whole test-time frame (hand-off + forward render) on both sides.  CUDA events, inputs resident, no_grad."""
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic, deformation
    from oracle.deform_torch import torch_get_deformation_eval
    from oracle import ref_loader
    scene, cam = synthetic.config2_scene()
    pc = synthetic.model_to(synthetic.dynamic_model(scene), dev)
    rs_args = (cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0, cam.viewmatrix.to(dev),
               cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    native_rast = sgs.GaussianRasterizer(sgs.GaussianRasterizationSettings(*rs_args))
    ref_rast = None
    if ref_loader.available():
        ref_rast = ref_loader.ref_api()[1](sgs.GaussianRasterizationSettings(*rs_args))
    stamps = [0.1 + 0.8 * i / (iters + 2) for i in range(iters + 3)]

    def run(deform, rast):
        ms, sel = [], 0
        with torch.no_grad():
            for i, t in enumerate(stamps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                m3, rot, sc, op, shs = deform(pc, t)
                if rast is not None:
                    rast(means3D=m3, means2D=torch.zeros_like(m3), opacities=op, shs=shs, scales=sc, rotations=rot)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ms.append(e0.elapsed_time(e1))
                    sel += m3.shape[0]
        return sum(ms) / len(ms), sel / len(ms)

    fused_ms, sel = run(deformation.get_deformation_eval, None)
    torch_ms, _ = run(torch_get_deformation_eval, None)
    frame_native, _ = run(deformation.get_deformation_eval, native_rast)
    out = {"what": "get_deformation_eval on 300k Gaussians (32 plane features, 3 MLPs 41-128-128-{3,7,48}), "
                   "timestamps swept over [0.1, 0.9]",
           "selected_mean": sel, "fused_ms": fused_ms, "pytorch_ops_ms": torch_ms, "speedup": torch_ms / fused_ms,
           "fused_TFLOPs_fp32_equivalent": sel * 2 * (3 * 41 * 128 + 3 * 128 * 128 + 128 * 58) / (fused_ms * 1e-3) / 1e12,
           "test_time_frame_ms": {"native": frame_native}}
    if ref_rast is not None:
        frame_ref, _ = run(torch_get_deformation_eval, ref_rast)
        out["test_time_frame_ms"]["reference_ops_and_rasterizer"] = frame_ref
        out["test_time_frame_ms"]["speedup"] = frame_ref / frame_native
    return out


def deform_train_path_timing(dev, N=300_000, iters=8):
    """SURVEY.md section 8(f) rank 1, training half: GaussianModel.get_deformation (scene/saro_gaussian.py:779-847) forward
    + backward on the headline cloud with the N3D configuration's flags (scale_reg on: six MLP evaluations per view,
    five of them differentiated), native tcgen05 job kernels vs the same function as the float32 PyTorch ops SaRO-GS
    runs.  The plane field is a resident feature table on both sides (the sampler has its own leg, plane_path).
    CUDA events, inputs resident."""
    from saro_gs_b200 import deformation
    from oracle import deform_torch
    g = torch.Generator().manual_seed(5)
    rn = lambda *s: torch.randn(*s, generator=g)
    t = dict(xyz=rn(N, 3) * 2, rotation=rn(N, 4), scaling=rn(N, 3) * 0.5 - 3.5, opacity=rn(N, 1) * 2, features_dc=rn(N, 1, 3) * 0.5,
             features_rest=rn(N, 15, 3) * 0.1, temporal_pos=torch.rand(N, 1, generator=g), hexplane_feature=rn(N, 32) * 0.5)
    leaves = {k: v.to(dev).requires_grad_(True) for k, v in t.items()}
    mlps = deform_torch.make_train_mlps(32, device=dev, seed=6)
    pc = deform_torch.TrainModelStandIn(leaves, mlps, (1, 0, 0), 6.0, 300.0)
    w = [rn(N, 3).to(dev), rn(N, 4).to(dev), (rn(N, 3) * 20).to(dev), rn(N, 1).to(dev), rn(N, 16, 3).to(dev)]
    params = [p for m in mlps.values() for p in m.parameters()] + list(leaves.values())

    def run(fn, steps, marks=None):
        f_ms, b_ms, phases = [], [], {}
        for i in range(steps + 2):
            for p in params:
                p.grad = None
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            deformation._phase_marks = [] if marks else None
            e0.record()
            outs = fn(pc, 0.35 + 0.05 * (i - steps - 1))      # both arms end on the same timestamp
            loss = deform_torch.train_objective(pc, outs, w, (8e-6, 0.0, 0.0))
            e1.record()
            loss.backward()
            e2.record()
            torch.cuda.synchronize()
            if i >= 2:
                f_ms.append(e0.elapsed_time(e1))
                b_ms.append(e1.elapsed_time(e2))
                ms = deformation._phase_marks or []
                for (_, a), (name, b) in zip(ms, ms[1:]):
                    phases[name] = phases.get(name, 0.0) + a.elapsed_time(b) / steps
        deformation._phase_marks = None
        grads = {i: p.grad.detach().clone() for i, p in enumerate(params) if p.grad is not None}
        return sum(f_ms) / len(f_ms), sum(b_ms) / len(b_ms), phases, grads

    nf, nb, phases, g_n = run(deformation.get_deformation, iters, marks=True)
    tf, tb, _, g_t = run(deform_torch.torch_get_deformation, 3)
    agree = max(float((g_n[i] - g_t[i]).abs().max() / g_t[i].abs().max().clamp_min(1e-30)) for i in g_t)
    flops = N * 2 * (5 * (41 * 128 + 128 * 128) + 128 * (1 + 3 + 7 + 48 + 7)) + N * 2 * 128 * 3
    return {"what": "get_deformation forward + backward on %d Gaussians (scale_reg on: lifespan, motion, rot, shs at t, rot and "
                    "motion at the base feature), objective = weighted outputs + scale regulariser" % N,
            "native_forward_ms": nf, "native_backward_ms": nb, "pytorch_ops_forward_ms": tf, "pytorch_ops_backward_ms": tb,
            "speedup_forward": tf / nf, "speedup_backward": tb / nb, "speedup_total": (tf + tb) / (nf + nb),
            "native_backward_phases_ms": phases, "forward_mlp_GFLOP": flops / 1e9,
            "max_rel_diff_gradients_vs_pytorch_float32": agree,
            "note": "forward = one tcgen05 launch for the six evaluations + the reference's elementwise epilogues; backward = "
                    "autograd of those epilogues, one tcgen05 launch for the five data-gradient chains, weight gradients; "
                    "gradient agreement on random rows is limited by ReLU kinks (tests/test_deform_train_gpu.py), not by "
                    "arithmetic"}


def train_view_path_timing(dev, iters=6):
    """One training view of the dynamic model end to end — everything between the parameters and their gradients
    (renderer/__init__.py:92-140, train.py:199-226, helper_train.py:50-70): scale-aware plane sampler ->
    get_deformation (six MLP evaluations) -> rasterizer -> L1 + D-SSIM + scale regulariser -> backward to the planes,
    the four MLPs and the Gaussians.  Native arm: this repo's kernels for all four stages.  Reference-ops arm: the same
    functions as the PyTorch ops SaRO-GS runs (oracle/plane_torch.py — grid_sample / avg_pool2d stand-in for the
    un-vendored nvdiffrast op —, oracle/deform_torch.py, oracle/ssim_torch.py) around the compiled reference rasterizer
    when oracle/_ref is present (else the native one; `reference_rasterizer` says which).  configs[1] cloud and camera,
    planes [512, 512, 512, 256] x 32 features (configs/neural_3D/*.json).  CUDA events, inputs resident."""
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic, deformation, loss_utils
    from saro_gs_b200.hexplane import ScaleAwareResField
    from oracle import deform_torch, plane_torch, ref_loader
    from oracle.ssim_torch import torch_l1_dssim_loss
    scene, cam = synthetic.config2_scene()
    P = scene.means3D.shape[0]
    g = torch.Generator().manual_seed(21)
    op = scene.opacities.clamp(1e-4, 1 - 1e-4)
    # lifespans are learned: bias the lifespan MLP so that most Gaussians are alive at the rendered timestamps
    t = dict(xyz=scene.means3D.clone(), rotation=scene.rotations.clone(), scaling=torch.log(scene.scales),
             opacity=torch.log(op / (1 - op)).reshape(P, 1), features_dc=scene.shs[:, :1, :].contiguous(),
             features_rest=scene.shs[:, 1:, :].contiguous(), temporal_pos=torch.rand(P, 1, generator=g))
    leaves = {k: v.to(dev).requires_grad_(True) for k, v in t.items()}
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32, "resolution": [512, 512, 512, 256]}
    field = ScaleAwareResField(cfg, [1]).to(dev)
    with torch.no_grad():
        for p in field.grids[0]:
            p.copy_((torch.randn(p.shape, generator=g) * 0.2).to(dev))
    lo, hi = scene.means3D.min(0).values - 0.5, scene.means3D.max(0).values + 0.5
    field.set_aabb(hi.tolist(), lo.tolist(), 300)
    mlps = deform_torch.make_train_mlps(32, device=dev, seed=22)
    with torch.no_grad():
        for name in ("motion", "rot", "shs"):
            mlps[name][4].weight.mul_(0.02)
            mlps[name][4].bias.mul_(0.02)
        mlps["opacity"][4].bias.fill_(-4.0)               # lifespan = 1 - sigmoid(.) close to 1: everything is alive
    pc = deform_torch.TrainModelStandIn(leaves, mlps, (1, 0, 0), 6.0, 300.0, hexplane=field)
    torch_field = lambda pts, ts, sc: plane_torch.field_forward(field, pts, ts, sc)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    native_rast = sgs.GaussianRasterizer(rs)
    ref_rast = ref_loader.ref_api()[1](rs) if ref_loader.available() else None
    gt = torch.rand(3, cam.height, cam.width, generator=g).to(dev)
    params = list(leaves.values()) + [p for m in mlps.values() for p in m.parameters()] + list(field.parameters())

    def run(native, steps):
        pc.hexplane = field if native else torch_field
        deform = deformation.get_deformation if native else deform_torch.torch_get_deformation
        rast = native_rast if native or ref_rast is None else ref_rast
        loss_fn = loss_utils.l1_dssim_loss if native else torch_l1_dssim_loss
        ms, parts = [], [0.0] * 4
        for i in range(steps + 2):
            for p in params:
                p.grad = None
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            m, rot, sc, opa, shs = deform(pc, 0.35 + 0.01 * i)
            ev[1].record()
            image = rast(means3D=m, means2D=torch.zeros_like(m), opacities=opa, shs=shs, scales=sc, rotations=rot)[0]
            ev[2].record()
            loss = loss_fn(image, gt, 0.2) + 8e-6 * torch.linalg.vector_norm(pc.scale_residual, ord=2)
            ev[3].record()
            loss.backward()
            ev[4].record()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(ev[0].elapsed_time(ev[4]))
                for k in range(4):
                    parts[k] += ev[k].elapsed_time(ev[k + 1]) / steps
        assert all(p.grad is not None for p in params)
        return sum(ms) / len(ms), parts, float(loss.detach())

    n_ms, n_parts, n_loss = run(True, iters)
    r_ms, r_parts, r_loss = run(False, 3)
    names = ("plane_sampler_and_deformation_forward", "rasterizer_forward", "loss_forward", "backward_of_everything")
    return {"what": "one training view of the dynamic model: plane sampler -> get_deformation -> rasterizer -> loss -> backward "
                    "(%d Gaussians @%dx%d)" % (P, cam.width, cam.height),
            "native_ms": n_ms, "reference_ops_ms": r_ms, "speedup": r_ms / n_ms,
            "native_parts_ms": dict(zip(names, n_parts)), "reference_ops_parts_ms": dict(zip(names, r_parts)),
            "reference_rasterizer": "compiled reference (oracle/_ref)" if ref_rast is not None else "native (oracle/_ref absent)",
            "loss_native": n_loss, "loss_reference_ops": r_loss,
            "note": "the reference-ops plane sampler is a grid_sample / avg_pool2d composition (nvdiffrast is not in this image; "
                    "its own kernel would be faster than this stand-in): plane_path separates that stage"}


def plane_path_timing(dev, N=300_000, iters=10):
    """the scale-aware plane sampler (ScaleAwareResField.forward + backward to the
    planes) at the N3D configuration of the reference (configs/neural_3D/*.json: resolution [512, 512, 512, 256], 32
    features, multires [1]) on 300 k points: sm_100a kernels vs the same field as PyTorch ops (grid_sample / avg_pool2d
    stand-in for the un-vendored nvdiffrast op).  CUDA events, inputs resident.  Algorithmic bytes: 36 taps of 128 B per
    point (3 space planes x 2 mip levels x 4 texels + 3 time planes x 4 texels) + 128 B out, forward; the same again as
    read-modify-write traffic backward."""
    from saro_gs_b200.hexplane import ScaleAwareResField
    from oracle import plane_torch
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32, "resolution": [512, 512, 512, 256]}
    field = ScaleAwareResField(cfg, [1]).to(dev)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in field.grids[0]:
            p.copy_((torch.randn(p.shape, generator=g) * 0.2).to(dev))
    xyz_max, xyz_min, duration = [12.0, 9.0, 40.0], [-12.0, -9.0, 4.0], 300
    field.set_aabb(xyz_max, xyz_min, duration)
    ext = torch.tensor(xyz_max) - torch.tensor(xyz_min)
    pts = (torch.tensor(xyz_min) + torch.rand(N, 3, generator=g) * ext).to(dev)
    ts = (torch.rand(N, 1, generator=g) * (duration - 1) / duration).to(dev)
    scales = torch.exp(torch.randn(N, 3, generator=g) * 0.9 - 3.0).to(dev)
    dout = torch.randn(N, 32, generator=g).to(dev)

    def run(fn, steps):
        ms_f, ms_b, last = [], [], None
        for i in range(steps + 2):
            for p in field.grids[0]:
                p.grad = None
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            out = fn()
            e1.record()
            out.backward(dout)
            e2.record()
            torch.cuda.synchronize()
            if i >= 2:
                ms_f.append(e0.elapsed_time(e1))
                ms_b.append(e1.elapsed_time(e2))
            last = out.detach()
        return sum(ms_f) / len(ms_f), sum(ms_b) / len(ms_b), last, [p.grad.clone() for p in field.grids[0]]

    f_ms, b_ms, out_n, grads_n = run(lambda: field(pts, ts, scales), iters)
    tf_ms, tb_ms, out_t, grads_t = run(lambda: plane_torch.field_forward(field, pts, ts, scales), 3)
    agree = float((out_n - out_t).abs().max() / out_t.abs().max())
    gagree = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(grads_n, grads_t))
    fwd_bytes = N * (36 * 128 + 128 + 28)
    return {"what": "ScaleAwareResField forward + backward, %d points, planes [512,512,512,256] x 32 features" % N,
            "fused_forward_ms": f_ms, "fused_backward_ms": b_ms, "pytorch_ops_forward_ms": tf_ms,
            "pytorch_ops_backward_ms": tb_ms, "speedup_forward": tf_ms / f_ms, "speedup_backward": tb_ms / b_ms,
            "forward_GBps": fwd_bytes / (f_ms * 1e-3) / 1e9, "forward_algorithmic_bytes": fwd_bytes,
            "max_rel_diff_forward_vs_pytorch": agree, "max_rel_diff_plane_grads_vs_pytorch": gagree,
            "note": "forward includes nothing but the sampling kernel once the channels-last pyramid is built (it is "
                    "rebuilt only when a plane changes); backward = memset + scatter + 6 fold kernels"}


def densify_path_timing(dev, P=300_000, V=4, iters=20):
    """SURVEY.md section 8(f) rank 4 (the part that consumes the rasterizer's outputs): densification statistics of a
    V-view batch, fused kernels vs the reference's list / stack / boolean-index PyTorch ops (train.py:211-215,
    :281-291).  CUDA events, inputs resident."""
    import types
    from saro_gs_b200.densify import BatchDensifyStats
    g = torch.Generator().manual_seed(0)
    grads = [(torch.randn(P, 3, generator=g) * 1e-4).to(dev) for _ in range(V)]
    radii = [torch.where(torch.rand(P, generator=g) < 0.3, 0, torch.randint(1, 60, (P,), generator=g)).to(torch.int32).to(dev)
             for _ in range(V)]

    def model():
        return types.SimpleNamespace(max_radii2D=torch.zeros(P, device=dev), xyz_gradient_accum=torch.zeros(P, 1, device=dev),
                                     denom=torch.zeros(P, 1, device=dev))

    def fused(m, stats):
        stats.reset()
        for a, b in zip(grads, radii):
            stats.add_view(a, b)
        stats.commit(m)

    def torch_ops(m, _):
        batch_point_grad, batch_radii, batch_vis = [], [], []
        for a, b in zip(grads, radii):
            batch_point_grad.append(torch.norm(a[:, :2], dim=-1))
            batch_radii.append(b)
            batch_vis.append(b > 0)
        visibility_count = torch.stack(batch_vis, 1).sum(1)
        visibility_filter = visibility_count > 0
        r = torch.stack(batch_radii, 1).max(1)[0]
        gr = torch.stack(batch_point_grad, 1).sum(1)
        gr[visibility_filter] = gr[visibility_filter] / visibility_count[visibility_filter]
        gr = gr.unsqueeze(1)
        m.max_radii2D[visibility_filter] = torch.max(m.max_radii2D[visibility_filter], r[visibility_filter])
        m.xyz_gradient_accum[visibility_filter] += gr[visibility_filter]
        m.denom[visibility_filter] += 1

    def run(fn):
        m, stats, ms = model(), BatchDensifyStats(P, dev), []
        for i in range(iters + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(m, stats)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
        return sum(ms) / len(ms), m

    a, ma = run(fused)
    b, mb = run(torch_ops)
    agree = torch.equal(ma.max_radii2D, mb.max_radii2D) and torch.equal(ma.denom, mb.denom) and \
        torch.allclose(ma.xyz_gradient_accum, mb.xyz_gradient_accum, rtol=1e-5)
    return {"what": "densification statistics, %d views x %d Gaussians (6 launches: launch-latency bound)" % (V, P),
            "fused_ms": a, "pytorch_ops_ms": b, "speedup": b / a, "agree": bool(agree),
            "in_backward": densify_in_backward_timing(dev)}


def densify_in_backward_timing(dev, iters=20):
    """The per-view statistics update as an epilogue of the rasterizer's backward (BatchDensifyStats.attach_next_backward
    -> sgs_densify_attach): rasterizer forward + backward of the configs[1] view with the sink armed, without it, and
    without it but followed by the stand-alone add_view kernel."""
    import saro_gs_b200 as sgs
    from saro_gs_b200.densify import BatchDensifyStats
    scene, cams, params, cot = make_inputs(dev, 0)
    c = cams[0]
    rs = sgs.GaussianRasterizationSettings(c.height, c.width, c.tanfovx, c.tanfovy, torch.zeros(3, device=dev), 1.0,
                                           c.viewmatrix.to(dev), c.projmatrix.to(dev), scene.sh_degree, c.campos.to(dev), False)
    cot = cot.to(dev)
    P = params["means3D"].shape[0]
    stats = BatchDensifyStats(P, dev)

    def step(mode):
        m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii, _ = sgs.GaussianRasterizer(rs)(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"],
                                                     shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        if mode == "armed":
            stats.attach_next_backward()
        color.backward(cot)
        if mode == "separate":
            stats.add_view(m2d.grad, radii)
        for p in params.values():
            p.grad = None

    out = {}
    for mode in ("plain", "armed", "separate"):
        for _ in range(5):
            step(mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step(mode)
        e1.record()
        torch.cuda.synchronize()
        out[mode + "_ms_per_view"] = e0.elapsed_time(e1) / iters
    out["what"] = "rasterizer forward + backward of the configs[1] view: no statistics / statistics in the backward's " \
                  "epilogue / stand-alone add_view kernel after backward"
    return out


def cpu_baseline(precision="f32"):
    """CPU oracle port on the host cores: one forward + one backward of the full config-2 frame."""
    from oracle import oracle
    from saro_gs_b200 import synthetic
    oracle.build()
    scene, cam = synthetic.config2_scene()
    cot = synthetic.cotangent(cam.height, cam.width)
    t0 = time.perf_counter()
    r = oracle.forward_scene(scene, cam, torch.zeros(3), precision=precision)
    r.backward(cot)
    ms = (time.perf_counter() - t0) * 1e3
    # BASELINE.json configs[0]: 10 k random Gaussians, one pinhole camera @400x400, forward RGB only, CPU splat
    s0, c0 = synthetic.config1_scene()
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        oracle.forward_scene(s0, c0, torch.zeros(3), precision=precision)
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return {"value": ms, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "1 full frame (forward + backward) of the same workload, C oracle float32 + OpenMP",
            "config0_cpu_splat_forward_ms": best,
            "config0": "BASELINE.json configs[0]: 10k Gaussians @400x400, forward only, same CPU oracle, best of 3"}


def run_cpu_reference(args, rank, world, why):
    """Fallback reference arm when oracle/_ref is absent: the CPU oracle port."""
    if rank != 0:
        return
    K = max(1, min(args.steps, 3))
    vals = [cpu_baseline()["value"] for _ in range(K)]
    v = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": 0, "ms_per_step": v, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{K} full frame(s); compiled reference unavailable ({why})"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sequence", action="store_true", help="skip the 300-frame configs[2] FPS protocol")
    ap.add_argument("--no-extras", action="store_true",
                    help="developer A/B runs: skip the widened-row legs (loss / deformation / densify / plane sampler)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path (by design)")
    run_native_or_ref(args, args.impl)


if __name__ == "__main__":
    main()
