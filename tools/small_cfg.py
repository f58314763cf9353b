import sys, json
sys.path.insert(0, ".")
import torch
import bench
print(json.dumps(bench.small_config_run(torch.device("cuda:0"))))
