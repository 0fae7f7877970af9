"""Developer tool (GPU box): the train_view_path leg of bench.py alone."""
import json
import sys
sys.path.insert(0, ".")
import torch
import bench
print(json.dumps(bench.train_view_path_timing(torch.device("cuda:0")), indent=1))
