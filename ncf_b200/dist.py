"""Multi-GPU training (one process per GPU, torch.distributed over NCCL / NVLink).

The reference is single-process; both modes below are new design that must
reproduce the single-process semantics at the GLOBAL batch.

ReplicatedDataParallel — BASELINE config 4 (MovieLens-sized tables, replicated).  Every rank holds
the full tables and Adam state and processes its own slice of the global batch.  Per step:
  1. all_gather of the batch indices (16 B/sample): every rank learns which rows the global batch
     touches, registers them and replays their pending zero-gradient Adam steps (ncf_adam_prepare);
  2. fused forward+loss+backward on the local slice into the local gradient buffer;
  3. ONE all-reduce (average) of the flat gradient buffer — embedding-row gradients and tower
     gradients together.  north_star names the tower all-reduce; the row gradients have to travel
     too or the replicas diverge.  NCCL's all-reduce leaves bit-identical
     results on every rank, so the replicas apply identical updates and stay bit-identical;
  4. sparse-row Adam over the union of touched rows + dense Adam on the tower.
The gradient of the global-batch mean loss is the rank-average of the local-mean gradients, hence
ReduceOp.AVG.  The dense all-reduce moves the whole buffer (zeros included); a sparse
(index, row) exchange is the planned refinement and does not change results.

`partition` / `union_rows` are pure functions so that the plan is testable on CPU (gloo).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def partition(n: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of n samples owned by `rank` (remainder to the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_indices(user: torch.Tensor, item: torch.Tensor, world: int):
    """all_gather of the per-rank (user, item) index slices -> global [world * B] tensors."""
    B = user.numel()
    both = torch.stack([user, item])                      # [2, B]
    out = torch.empty(world, 2, B, dtype=user.dtype, device=user.device)
    dist.all_gather_into_tensor(out.view(-1), both.view(-1))
    return out[:, 0, :].reshape(-1).contiguous(), out[:, 1, :].reshape(-1).contiguous()


def average_(t: torch.Tensor, world: int):
    """In-place rank average (NCCL has ReduceOp.AVG; gloo does not)."""
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(world)
    return t


class ReplicatedDataParallel:
    """Wraps a FusedTrainStep whose model is replicated on every rank."""

    def __init__(self, ts, check_replicas: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.ts = ts
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if ts.grads.flat is None:
            raise RuntimeError("gradient buffers must be one flat allocation")
        # the global batch may touch up to world * B distinct rows
        dev = ts.device
        cap = ts.max_batch * self.world
        ts.grads.user_list = torch.zeros(min(cap, ts.model.user_num), dtype=torch.int64, device=dev)
        ts.grads.item_list = torch.zeros(min(cap, ts.model.item_num), dtype=torch.int64, device=dev)
        ts._refresh()
        if check_replicas:
            for p in ts.model.parameters():  # start from rank 0's weights
                dist.broadcast(p.data, src=0)

    def step(self, user, item, label):
        ts = self.ts
        gu, gi = gather_indices(user, item, self.world)
        if ts.optimizer == "adam":
            ops.adam_prepare(ts._m, ts._g, ts._s, gu, gi, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        else:
            raise NotImplementedError("replicated DP is implemented for Adam")
        ops.train_step_grads(ts._m, ts._g, user, item, label, None, 1.0, ts.loss_accum, ts.workspace)
        average_(ts.grads.flat, self.world)
        ops.adam_step(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        ts._dirty = True
        ts.num_steps += 1

    def replica_divergence(self) -> float:
        """max over parameters and ranks of |w_rank - w_0| (0.0 when the replicas are identical)."""
        self.ts.flush()
        worst = torch.zeros(1, device=self.ts.device)
        for p in self.ts.model.parameters():
            ref = p.data.clone()
            dist.broadcast(ref, src=0)
            worst = torch.maximum(worst, (p.data - ref).abs().max().reshape(1))
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        return float(worst.item())
