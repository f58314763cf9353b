"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU).
Checks that N data-parallel ranks (a) keep their replicated parameters bit-identical and (b) reproduce
the single-process result at the global batch.  Layout from the environment: user-partitioned by
default (every rank trains on the samples of its own user range), NCF_DP_PARTITION=0 = fully
replicated (contiguous slices of the batch); DP_TEACHER=1 adds a response-KD teacher."""
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
    if os.environ.get("DP_BIG") == "1":     # large enough for the tcgen05 path and the all-rows optimiser
        U, I, f, L, B, T = 3000, 2000, 32, 3, 10000 * world, 4
    rng = np.random.default_rng(0)
    users = torch.from_numpy(rng.integers(0, U, (T, B))).to(dev)
    items = torch.from_numpy(rng.integers(0, I, (T, B))).to(dev)
    labels = torch.from_numpy((rng.random((T, B)) < 0.3).astype(np.float32)).to(dev)
    teacher = None
    if os.environ.get("DP_TEACHER") == "1":
        torch.manual_seed(7)
        teacher = NCF(U, I, 2 * f, L, 0.0, "NeuMF-end").to(dev).eval()

    torch.manual_seed(0)
    model = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
    ts = FusedTrainStep(model, "adam", 1e-3, max_batch=B, teacher=teacher, alpha=0.5)
    dp = ReplicatedDataParallel(ts, global_batch=B)
    lo, hi = partition(B, world, rank)
    for t in range(T):
        if dp.partition_users:   # the samples of the global batch whose user this rank owns
            mine = (users[t] >= dp.user_lo) & (users[t] < dp.user_hi)
            dp.step(users[t][mine].contiguous(), items[t][mine].contiguous(), labels[t][mine].contiguous(),
                    global_batch=B)
        else:
            dp.step(users[t, lo:hi].contiguous(), items[t, lo:hi].contiguous(), labels[t, lo:hi].contiguous())
    loss_dp = dp.global_loss()
    divergence = dp.replica_divergence()

    # single-process reference at the global batch (every rank computes it; identical inputs)
    torch.manual_seed(0)
    ref = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
    rts = FusedTrainStep(ref, "adam", 1e-3, max_batch=B, teacher=teacher, alpha=0.5)
    for t in range(T):
        rts.step(users[t], items[t], labels[t])
    rts.flush()
    loss_ref = rts.pop_loss()
    # `worst`: largest deviation relative to the tensor's largest entry.  `share_over`: share of elements
    # beyond 2e-4.  On the tcgen05 path the three MMA-issuing warps accumulate in a run-dependent order,
    # so a ReLU pre-activation within fp32 rounding of zero (a few samples per 10^4) may take the other
    # branch in the other run: that sample's rows then differ by O(10 %) of one Adam step while every
    # other element agrees - the large-batch test bounds the share instead of the maximum.
    worst, over, total = 0.0, 0, 0
    for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):
        scale = max(b.abs().max().item(), 1e-30)
        d = (a - b).abs() / scale
        worst = max(worst, d.max().item())
        over += int((d > 2e-4).sum().item())
        total += d.numel()
    if rank == 0:
        print(json.dumps({"divergence": divergence, "vs_single_process": worst, "share_over": over / total, "world": world,
                          "partitioned": dp.partition_users, "p2p_tail": dp.tail is not None, "loss_dp": loss_dp,
                          "loss_single": loss_ref}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
