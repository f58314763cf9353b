"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU).
Checks that N data-parallel ranks on replicated tables (a) stay bit-identical and (b) reproduce the
single-process result at the global batch."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from ncf_b200.dist import ReplicatedDataParallel, partition  # noqa: E402
from ncf_b200.models import NCF  # noqa: E402
from ncf_b200.trainer import FusedTrainStep  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    U, I, f, L, B, T = 300, 200, 16, 2, 64 * world, 6
    rng = np.random.default_rng(0)
    users = torch.from_numpy(rng.integers(0, U, (T, B))).to(dev)
    items = torch.from_numpy(rng.integers(0, I, (T, B))).to(dev)
    labels = torch.from_numpy((rng.random((T, B)) < 0.3).astype(np.float32)).to(dev)

    torch.manual_seed(0)
    model = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
    ts = FusedTrainStep(model, "adam", 1e-3, max_batch=B // world)
    dp = ReplicatedDataParallel(ts)
    lo, hi = partition(B, world, rank)
    for t in range(T):
        dp.step(users[t, lo:hi].contiguous(), items[t, lo:hi].contiguous(), labels[t, lo:hi].contiguous())
    divergence = dp.replica_divergence()

    # single-process reference at the global batch (every rank computes it; identical inputs)
    torch.manual_seed(0)
    ref = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
    rts = FusedTrainStep(ref, "adam", 1e-3, max_batch=B)
    for t in range(T):
        rts.step(users[t], items[t], labels[t])
    rts.flush()
    worst = 0.0
    for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):
        scale = max(b.abs().max().item(), 1e-30)
        worst = max(worst, (a - b).abs().max().item() / scale)
    if rank == 0:
        print(json.dumps({"divergence": divergence, "vs_single_process": worst, "world": world}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
