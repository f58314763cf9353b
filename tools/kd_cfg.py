"""Runs bench.kd_config_run (config 3: response distillation at batch 256) alone and prints its JSON: the command
to put under ncu for the per-kernel times of the distillation step."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
print(json.dumps(bench.kd_config_run(torch.device("cuda:0"), steps)))
