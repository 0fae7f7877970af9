"""Developer tool (GPU box): the deform_train_path leg of bench.py alone."""
import json
import sys
sys.path.insert(0, ".")
import torch
import bench
print(json.dumps(bench.deform_train_path_timing(torch.device("cuda:0")), indent=1))
