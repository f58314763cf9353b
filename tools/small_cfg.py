"""Runs bench.small_config_run (NeuMF f=8 L=3, ML-1M shape, batch 256, CUDA-graph windows) alone and prints its
JSON: the command to put under ncu for the per-kernel times of the small configuration."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
print(json.dumps(bench.small_config_run(torch.device("cuda:0"), steps)))
