"""Developer tool (GPU box): where the strict e2e step's host time goes (CPU timestamps of one step's phases)."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
import saro_gs_b200 as sgs

dev = torch.device("cuda:0")
scene, cams, params, cot_cpu = bench.make_inputs(dev, 0)
H, W = cams[0].height, cams[0].width
means2D = torch.zeros_like(params["means3D"], requires_grad=True)
pin = lambda t: t.contiguous().pin_memory()
cot_pin = pin(cot_cpu)
bg_cpu = torch.zeros(3)
view_pin = [pin(torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1), bg_cpu])) for c in cams]
res_host = torch.empty((2,), dtype=torch.float32).pin_memory()
s_in = torch.cuda.Stream(dev)
ev_in = torch.cuda.Event()
Rast, Settings = sgs.GaussianRasterizer, sgs.GaussianRasterizationSettings
acc = [0.0] * 7
N = 60
for i in range(N + 10):
    torch.cuda.synchronize()
    main = torch.cuda.current_stream(dev)
    c = cams[i % len(cams)]
    t = [time.perf_counter()]
    with torch.cuda.stream(s_in):
        cot = cot_pin.to(dev, non_blocking=True)
        ev_in.record(s_in)
    pk = view_pin[i % len(cams)].to(dev, non_blocking=True)
    v, p, cp, bg = pk[0:16].view(4, 4), pk[16:32].view(4, 4), pk[32:35], pk[35:38]
    rs = Settings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, v, p, scene.sh_degree, cp, False)
    t.append(time.perf_counter())
    color, radii, depth = Rast(rs)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                   shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
    t.append(time.perf_counter())
    main.wait_event(ev_in)
    cot.record_stream(main)
    color.backward(cot)
    t.append(time.perf_counter())
    res = torch.stack([(color.detach() * cot).sum(), params["means3D"].grad.abs().sum()])
    res_host.copy_(res, non_blocking=True)
    t.append(time.perf_counter())
    main.synchronize()
    t.append(time.perf_counter())
    for q in list(params.values()) + [means2D]:
        q.grad = None
    t.append(time.perf_counter())
    if i >= 10:
        for k in range(6):
            acc[k] += (t[k + 1] - t[k]) * 1e6 / N
        acc[6] += (t[-1] - t[0]) * 1e6 / N
names = ["uploads + settings", "forward call", "backward call", "result ops + D2H enqueue", "synchronize (wait for GPU)", "zero grads", "total"]
for n, a in zip(names, acc):
    print(f"{n:32s} {a:8.1f} us")
