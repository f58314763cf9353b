"""Worker for tests/test_gpu_multi.py: row-sharded training on N ranks (tables and Adam state split
by row, item rows and their gradients exchanged over peer memory or by NCCL all-to-all:
NCF_SHARD_P2P=1/0) must reproduce single-process training at the global batch - with uneven shards, hot
rows, rows that skip steps and a step in which one rank has no samples at all."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from ncf_b200.dist import RowShardedTrainer, shard_rows, shard_state_dict  # noqa: E402
from ncf_b200.models import NCF  # noqa: E402
from ncf_b200.trainer import FusedTrainStep  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    U, I, f, L, Bg, T = 301, 203, 16, 2, 256, 6   # odd sizes: uneven shards
    rng = np.random.default_rng(0)
    users = rng.integers(0, U, (T, Bg))
    items = rng.integers(0, I, (T, Bg))
    users[:, :32] = rng.integers(0, 6, (T, 32))      # hot rows, duplicates inside a batch
    items[:, :32] = rng.integers(0, 5, (T, 32))
    for t in (2, 3):                                  # rows that skip steps -> lazy-Adam catch-up across ranks
        users[t] = np.where(users[t] < 40, users[t] + 40, users[t])
        items[t] = np.where(items[t] < 30, items[t] + 30, items[t])
    users[4] = users[4] // world * world              # step 4: every sample belongs to rank 0's users
    labels = (rng.random((T, Bg)) < 0.3).astype(np.float32)

    torch.manual_seed(0)
    full = NCF(U, I, f, L, 0.0, "NeuMF-end")
    full_sd = {k: v.clone() for k, v in full.state_dict().items()}

    shard = NCF(shard_rows(U, world, rank), shard_rows(I, world, rank), f, L, 0.0, "NeuMF-end")
    shard.load_state_dict(shard_state_dict(full_sd, world, rank))
    shard = shard.to(dev)
    tr = RowShardedTrainer(shard, U, I, lr=1e-3, max_batch=Bg)
    for t in range(T):
        mine = users[t] % world == rank               # sharding follows the data: my users' samples
        tr.step(torch.from_numpy(users[t][mine]).to(dev), torch.from_numpy(items[t][mine]).to(dev),
                torch.from_numpy(labels[t][mine]).to(dev), global_batch=Bg)
    got = tr.gather_full_state()
    loss = tr.loss_accum.clone()
    dist.all_reduce(loss)

    ref = NCF(U, I, f, L, 0.0, "NeuMF-end")
    ref.load_state_dict(full_sd)
    ref = ref.to(dev)
    rts = FusedTrainStep(ref, "adam", 1e-3, max_batch=Bg)
    for t in range(T):
        rts.step(torch.from_numpy(users[t]).to(dev), torch.from_numpy(items[t]).to(dev),
                 torch.from_numpy(labels[t]).to(dev))
    rts.flush()
    worst = 0.0
    for k, b in ref.state_dict().items():
        scale = max(b.abs().max().item(), 1e-30)
        worst = max(worst, (got[k] - b).abs().max().item() / scale)
    if rank == 0:
        print(json.dumps({"vs_single_process": worst, "loss_sharded": loss.item(),
                          "loss_single": rts.loss_accum.item(), "world": world, "p2p": tr.p2p}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
