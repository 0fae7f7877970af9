#!/usr/bin/env python
"""Developer tool (GPU box): time the densification-statistics leg of bench.py alone."""
import json
import sys
import torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.densify_path_timing(torch.device('cuda:0'))))
