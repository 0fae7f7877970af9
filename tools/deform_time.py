#!/usr/bin/env python
"""Developer tool (GPU box): time the deformation hand-off leg of bench.py alone."""
import json
import sys
import torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.deform_path_timing(torch.device('cuda:0'), iters=int(sys.argv[1]) if len(sys.argv) > 1 else 15)))
