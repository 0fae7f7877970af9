"""Developer tool (GPU box): the plane_path leg of bench.py alone (forward + backward of the scale-aware plane sampler)."""
import json
import sys
sys.path.insert(0, ".")
import torch
import bench
d = bench.plane_path_timing(torch.device("cuda:0"), iters=4)
print(json.dumps({k: v for k, v in d.items() if k not in ("note", "what")}))
